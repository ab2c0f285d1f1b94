"""Diagnostic (not a test): step GPU and oracle side by side, report the first diverging table."""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import oracle_lib
from parity_util import compare_worlds, is_exact, TABLES
from rapier_b200 import scenes
from rapier_b200.world import PhysicsWorld

name = sys.argv[1] if len(sys.argv) > 1 else "pile"
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 60
scene = {"pile": lambda: scenes.box_pile(4, 4, 5), "pyr": lambda: scenes.pyramids(2, 2, 10), "p3": lambda: scenes.pyramid3(10),
         "jg": lambda: scenes.joint_grid(20), "keva": lambda: scenes.keva(1), "keva5": lambda: scenes.keva(5), "jg100": lambda: scenes.joint_grid(100),
         "keva2": lambda: scenes.keva(2), "jg40": lambda: scenes.joint_grid(40)}[name]()
w = PhysicsWorld(scene)
o = oracle_lib.OracleWorld(scene)
for i in range(steps):
    print('step', i, flush=True)
    w.step(); o.step()
    d = compare_worlds(w, o)
    if not is_exact(d):
        print("first mismatch at step", i, d)
        kg, ko = w.debug_read("pair_keys", np.uint64), o.debug_read("pair_keys", np.uint64)
        sg, so = set(kg.tolist()), set(ko.tolist())
        print("pairs gpu", len(kg), "oracle", len(ko), "only gpu", [(k >> 32, k & 0xffffffff) for k in sorted(sg - so)][:10],
              "only oracle", [(k >> 32, k & 0xffffffff) for k in sorted(so - sg)][:10])
        print("sorted gpu?", bool((np.diff(kg.astype(np.int64)) > 0).all()))
        if len(kg) == len(ko) and (kg == ko).all():
            for t, dt in TABLES[1:]:
                a, b = w.debug_read(t, dt), o.debug_read(t, dt)
                bad = np.nonzero(a != b)[0]
                if len(bad):
                    per = {"pair_nsc": 1, "pair_npts": 1, "pair_color": 1, "pair_normal": 3, "pair_points": 36, "pair_data": 48, "pair_sc": 28}[t]
                    prs = np.unique(bad // per)
                    print(t, "mismatching pairs", prs[:10], "of", len(kg))
                    p0 = prs[0]
                    print("  pair", p0, "colliders", kg[p0] >> 32, kg[p0] & 0xffffffff)
                    print("  gpu   ", a[p0 * per:(p0 + 1) * per])
                    print("  oracle", b[p0 * per:(p0 + 1) * per])
        pg, vg = w.body_states(); po, vo = o.body_states()
        bad = np.nonzero((pg != po).any(axis=1) | (vg != vo).any(axis=1))[0]
        print("bodies differing:", bad[:20], "of", len(pg))
        if len(bad):
            b0 = bad[0]
            print(" gpu", pg[b0], vg[b0]); print(" orc", po[b0], vo[b0])
        print("gpu state", w.debug_read("state", np.int32)[:20])
        fa, fb = w.debug_read("collider_fat", np.float32), o.debug_read("collider_fat", np.float32)
        print("fat aabb mismatches", int((fa != fb).sum()))
        break
else:
    print("no mismatch in", steps, "steps")
