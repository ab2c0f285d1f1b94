"""Helper (not a test): phase timeline of one shared-memory item (RB_DEBUG_FLAGS=2) and solve time with the sweeps skipped (=1)."""
import sys, os
os.environ["RAPIER_B200_DEBUG_LIB"] = "1"   # the -DRB_DEBUG build (python -c "import __graft_entry__ as g; g.build_cuda(debug=True)")
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from rapier_b200 import scenes
from rapier_b200.world import PhysicsWorld
NAMES = ["start", "body_init", "generate", "W.begin", "W.end", "B.end", "integrate", "R.end", "sweeps.end", "cons_writeback", "body_writeback"]
for name, mk in (("80x20", scenes.many_pyramids_label), ("14x14x10", scenes.many_pyramids)):
    w = PhysicsWorld(mk())
    for _ in range(4):
        w.step(5)
    w.physics_pipeline.enable_profiling(True)
    w.step(50)
    c = w.counters()
    print(name, "flags", os.environ.get("RB_DEBUG_FLAGS"), "collide_us", round(c["collision_detection_ms"] * 1000, 1), "solve_us", round(c["solver_ms"] * 1000, 1), flush=True)
    if int(os.environ.get("RB_DEBUG_FLAGS", "0")) & 2:
        t = w.debug_read("dbg_times", np.int64)
        d = (t[1:len(NAMES)] - t[:len(NAMES) - 1]) / 1965.0
        print("   ", ", ".join(f"{n}={x:.1f}us" for n, x in zip(NAMES[1:], d)), f"total={(t[len(NAMES)-1]-t[0])/1965.0:.1f}us")
