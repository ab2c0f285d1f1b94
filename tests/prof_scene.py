import sys, os
sys.path.insert(0, os.getcwd())
from rapier_b200 import scenes
from rapier_b200.world import PhysicsWorld
name = sys.argv[1]; n = int(sys.argv[2])
scene = scenes.many_pyramids_label() if name == "80x20" else scenes.many_pyramids()
w = PhysicsWorld(scene)
w.step(5); w.step(5); w.step(n)
