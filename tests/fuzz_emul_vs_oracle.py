"""Differential fuzzing on the CPU (helper, also driven by tests/test_fuzz.py with a few seeds): random small
scenes -- cuboids and balls of random sizes, poses, velocities, materials, collision groups, locked axes,
additional mass, spherical / fixed / revolute joints -- stepped through the host emulation of the kernels and
through the oracle; every pose, velocity and persistent contact table must agree bit for bit.

    python tests/fuzz_emul_vs_oracle.py [first_seed] [count]
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def random_scene(seed):
    from rapier_b200 import _abi as A
    from rapier_b200 import scenes
    from rapier_b200.sets import (ColliderBuilder, FixedJointBuilder, RevoluteJointBuilder, RigidBodyBuilder,
                                  SphericalJointBuilder)
    r = np.random.default_rng(seed)
    s = scenes.Scene(f"fuzz_{seed}", gravity=(0.0, float(r.choice([-9.81, -10.0, -3.0])), 0.0))
    s.insert(RigidBodyBuilder.fixed().translation((0.0, -0.5, 0.0)),
             ColliderBuilder.cuboid(12.0, 0.5, 12.0).friction(float(r.uniform(0.0, 1.0))).restitution(float(r.choice([0.0, 0.0, 0.5]))))
    if r.random() < 0.3:   # a wall of parentless colliders
        for i in range(int(r.integers(1, 5))):
            s.colliders.insert(ColliderBuilder.cuboid(0.3, 1.5, 3.0).translation((4.0 + i * 0.7, 1.5, 0.0)))
    n = int(r.integers(2, 40))
    handles = []
    dense = r.random() < 0.5    # dense: bodies start close together (many contacts); sparse: mostly free fall
    span = 1.5 if dense else 5.0
    for i in range(n):
        pos = (float(r.uniform(-span, span)), float(r.uniform(0.4, 3.0 if dense else 8.0)), float(r.uniform(-span, span)))
        b = RigidBodyBuilder.dynamic().translation(pos)
        if r.random() < 0.7:
            b = b.rotation(tuple(float(x) for x in r.uniform(-1.0, 1.0, 3)))
        if r.random() < 0.5:
            b = b.linvel(tuple(float(x) for x in r.uniform(-3.0, 3.0, 3))).angvel(tuple(float(x) for x in r.uniform(-4.0, 4.0, 3)))
        if r.random() < 0.15:
            b = b.linear_damping(float(r.uniform(0.0, 1.0))).angular_damping(float(r.uniform(0.0, 1.0)))
        if r.random() < 0.1:
            b = b.locked_axes(int(r.choice([A.RB_BODY_LOCK_RX | A.RB_BODY_LOCK_RZ, A.RB_BODY_LOCK_TX, A.RB_BODY_LOCK_RY])))
        if r.random() < 0.1:
            b = b.additional_mass(float(r.uniform(0.5, 20.0)))
        if r.random() < 0.1:
            b = b.gyroscopic_forces_enabled(False)
        if r.random() < 0.6:
            c = ColliderBuilder.cuboid(*[float(x) for x in r.uniform(0.15, 0.7, 3)])
        else:
            c = ColliderBuilder.ball(float(r.uniform(0.15, 0.6)))
        c = c.density(float(r.choice([0.0, 0.5, 1.0, 10.0, 100.0]) if r.random() < 0.9 else 1.0))
        c = c.friction(float(r.uniform(0.0, 1.2))).restitution(float(r.choice([0.0, 0.0, 0.0, 0.3, 0.9, 1.0])))
        if r.random() < 0.1:
            c = c.collision_groups(int(r.choice([1, 2, 3])), int(r.choice([1, 2, 3])))
        if r.random() < 0.1:
            c = c.translation(tuple(float(x) for x in r.uniform(-0.3, 0.3, 3)))
        handles.append(s.insert(b, c))
    for _ in range(int(r.integers(0, max(1, n // 4)))):
        a, b = (int(x) for x in r.choice(handles, 2, replace=False)) if n >= 2 else (handles[0], handles[0])
        if a == b:
            continue
        kind = r.integers(0, 3)
        j = SphericalJointBuilder() if kind == 0 else (FixedJointBuilder() if kind == 1 else RevoluteJointBuilder(tuple(float(x) for x in r.choice([(1, 0, 0), (0, 0, 1), (0.6, 0.0, 0.8)]))))
        j = j.local_anchor1(tuple(float(x) for x in r.uniform(-0.5, 0.5, 3))).local_anchor2(tuple(float(x) for x in r.uniform(-0.5, 0.5, 3)))
        if r.random() < 0.3:
            j = j.contacts_enabled(False)
        s.joints.insert(a, b, j)
    params = A.RbIntegrationParameters.default()
    if r.random() < 0.4:
        params.num_solver_iterations = int(r.choice([1, 2, 6]))
        params.num_internal_pgs_iterations = int(r.choice([1, 2]))
        params.num_internal_stabilization_iterations = int(r.choice([0, 1, 2]))
        params.warmstart_coefficient = float(r.choice([0.0, 0.5, 1.0]))
        params.friction_in_bias_pass = int(r.choice([0, 1]))
        params.contact_recycling = int(r.choice([0, 1]))
    return s, params


def run(seed, steps=90, smem_floats=None):
    import emul_lib
    import oracle_lib
    from parity_util import compare_worlds, is_exact
    from rapier_b200.world import PhysicsWorld
    if smem_floats is not None:
        os.environ["RB_EMU_COOP_SMEM_FLOATS"] = str(smem_floats)
    else:
        os.environ.pop("RB_EMU_COOP_SMEM_FLOATS", None)
    scene, params = random_scene(seed)
    w = PhysicsWorld(scene, integration_parameters=params, _lib=emul_lib.lib())
    o = oracle_lib.OracleWorld(scene, params=params)
    for i in range(steps):
        w.step()
        o.step()
        if i % 15 == 14 or i < 2:
            d = compare_worlds(w, o)
            if not is_exact(d):
                return False, f"seed {seed} step {i}: {d}"
    pose, vel = w.body_states()
    if not (np.isfinite(pose).all() and np.isfinite(vel).all()):
        return False, f"seed {seed}: non-finite state"
    return True, ""


if __name__ == "__main__":
    first = int(sys.argv[1]) if len(sys.argv) > 1 else 0
    count = int(sys.argv[2]) if len(sys.argv) > 2 else 50
    bad = 0
    for seed in range(first, first + count):
        ok, msg = run(seed, smem_floats=(9000 if seed % 3 == 0 else None))
        if not ok:
            bad += 1
            print("MISMATCH", msg, flush=True)
    print(f"{count - bad}/{count} seeds bit-exact")
