"""Helper (not a test): the measurement protocol of SURVEY.md 8(d) for one scene on the GPU --
1 000 steps from construction, mean steps/s over steps 100-999 and over all 1 000, pose drift against the CPU
oracle after 1 000 steps (max over bodies of |x_gpu - x_cpu| / scene height and max quaternion angle), and the
per-step max |delta| of the first 10 steps.

    python tests/protocol_1000.py b3d_many_pyramids_80x20 [oracle_threads]
"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    import oracle_lib
    from rapier_b200 import scenes
    from rapier_b200.world import PhysicsWorld
    name = sys.argv[1] if len(sys.argv) > 1 else "b3d_many_pyramids_80x20"
    threads = int(sys.argv[2]) if len(sys.argv) > 2 else 16
    make = scenes.REGISTRY[name]
    scene = make()
    w = PhysicsWorld(scene)
    o = oracle_lib.OracleWorld(scene, threads=threads)
    first10 = []
    for _ in range(10):
        w.step(); o.step()
        pg, _ = w.body_states(); pc, _ = o.body_states()
        first10.append(float(np.abs(pg - pc).max()))
    t0 = time.perf_counter(); w.step(90); t_100 = time.perf_counter() - t0
    t0 = time.perf_counter(); w.step(900); t_900 = time.perf_counter() - t0
    o.step(990)
    pg, _ = w.body_states(); pc, _ = o.body_states()
    height = float(pc[:, 1].max() - pc[:, 1].min()) or 1.0
    drift = float(np.linalg.norm(pg[:, :3] - pc[:, :3], axis=1).max() / height)
    dots = np.clip(np.abs((pg[:, 3:] * pc[:, 3:]).sum(axis=1)), 0.0, 1.0)
    print(json.dumps({"scene": name, "bodies": int(pg.shape[0]), "steps_per_s_100_999": 900 / t_900,
                      "steps_per_s_10_999": 990 / (t_100 + t_900), "first10_max_abs_pose_delta": first10,
                      "drift_pos_over_height_after_1000": drift, "drift_max_quat_angle_rad": float(2 * np.arccos(dots).max()),
                      "bit_exact_after_1000": bool((pg.view(np.uint32) == pc.view(np.uint32)).all())}))


if __name__ == "__main__":
    main()
