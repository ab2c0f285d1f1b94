"""Regenerates the golden fixtures of tests/golden/ from the CPU oracle (run from the repo root):

    python tests/golden/make_golden.py

The reference has no reusable golden vectors for this path (its only pinned values are whole-binary
hashes), and it cannot be run here, so these fixtures pin the ORACLE -- body states after a
fixed number of steps of deterministic scenes -- against accidental changes between rounds.  They do not
pin parity with the reference ("parity unpinned" stays)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

CASES = {
    "pyramids_2x2x6_60": ("pyramids", (2, 2, 6), 60),
    "box_pile_3x3x4_150": ("box_pile", (3, 3, 4), 150),
    "single_pyramid_20_40": ("single_pyramid", (20,), 40),
    "joint_grid_8_120": ("joint_grid", (8,), 120),
    "keva_1_40": ("keva", (1,), 40),
    "ball_drop_100": ("box_on_ground", ("ball", 2.0), 100),
}


def build(name):
    from rapier_b200 import scenes
    fn, args, steps = CASES[name]
    return getattr(scenes, fn)(*args), steps


def main():
    import oracle_lib
    out = os.path.dirname(os.path.abspath(__file__))
    for name in CASES:
        scene, steps = build(name)
        o = oracle_lib.OracleWorld(scene)
        o.step(steps)
        pose, vel = o.body_states()
        np.savez_compressed(os.path.join(out, name + ".npz"), pose=pose.astype(np.float32), vel=vel.astype(np.float32), steps=np.int32(steps))
        print(name, pose.shape)


if __name__ == "__main__":
    main()
