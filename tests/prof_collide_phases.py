"""Helper (not a test): section timeline of k_collide (RB_DEBUG_FLAGS=4, the -DRB_DEBUG build): python tests/prof_collide_phases.py <scene>"""
import sys, os
os.environ["RAPIER_B200_DEBUG_LIB"] = "1"
os.environ["RB_DEBUG_FLAGS"] = "4"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from rapier_b200 import scenes
from rapier_b200.world import PhysicsWorld
MAKE = {"convex": lambda: scenes.convex_polyhedra(25), "keva5": lambda: scenes.keva(5), "p3_50": lambda: scenes.pyramid3(50), "80x20": scenes.many_pyramids_label}
NAMES = ["refresh", "broad", "narrow", "kin/wake/colour", "components+sleep", "schedule"]
SCHED = ["S1 reset", "S4 counts", "S5 stage order", "S6 item scan", "S7 item counts", "S8 scans", "S9 scatter", "S10 per-item sort", "S11 order"]
name = sys.argv[1]
w = PhysicsWorld(MAKE[name]())
w.step(int(sys.argv[2]) if len(sys.argv) > 2 else 70)
acc = np.zeros(len(NAMES))
sacc = np.zeros(len(SCHED))
n = 20
for _ in range(n):
    w.step(1)
    t = w.debug_read("dbg_times", np.int64)[16:16 + len(NAMES) + 1]
    acc += (t[1:] - t[:-1]) / 1965.0
    u = w.debug_read("dbg_times", np.int64)[16:32]
    ts = np.concatenate([[u[5]], u[7:15], [u[6]]])          # schedule entry (= stamp 5), S4 .. S11 starts
    sacc += (ts[1:] - ts[:-1]) / 1965.0
c = w.counters()
print(name, "pairs", c["num_pairs"], "manifolds", c["num_active_manifolds"], " ".join(f"{k}={v / n:.1f}us" for k, v in zip(NAMES, acc)), f"total={acc.sum() / n:.1f}us")
print("   schedule:", " ".join(f"{k}={v / n:.1f}us" for k, v in zip(SCHED, sacc)))
