"""Helper (not a test): step one scene a few times so a profiler can capture its kernels.
    python tests/prof_scene.py <scene> <steps>"""
import sys, os
sys.path.insert(0, os.getcwd())
from rapier_b200 import scenes
from rapier_b200.world import PhysicsWorld
name = sys.argv[1]; n = int(sys.argv[2])
MAKE = {"80x20": scenes.many_pyramids_label, "14x14x10": scenes.many_pyramids, "keva5": lambda: scenes.keva(5),
        "jg100": lambda: scenes.joint_grid(100), "p3_50": lambda: scenes.pyramid3(50), "p3_20": lambda: scenes.pyramid3(20),
        "lw300": lambda: scenes.large_world(grid=300, spheres=100), "convex": lambda: scenes.convex_polyhedra(25)}
w = PhysicsWorld(MAKE[name]())
w.step(5); w.step(5)
if name == "convex":
    w.step(60)   # let the hulls land: the capture should see a tumbling heap
w.step(n)
