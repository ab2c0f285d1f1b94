"""Differential fuzzing on the CPU (helper, also driven by tests/test_fuzz.py with a few seeds): random small
scenes -- cuboids, balls and capsules of random sizes, poses, velocities, materials, collision groups, locked axes,
additional mass, spherical / fixed / revolute / prismatic joints with random limits and motors, kinematic bodies
(velocity- and position-based), fast bodies (CCD), either friction model, collision / contact-force events --
stepped through the host emulation of the kernels and through the oracle; every pose, velocity, persistent
contact table, joint impulse and event must agree bit for bit.

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
    from rapier_b200.sets import (ColliderBuilder, FixedJointBuilder, PrismaticJointBuilder, RevoluteJointBuilder,
                                  RigidBodyBuilder, SphericalJointBuilder)
    r = np.random.default_rng(seed)
    r2 = np.random.default_rng(seed + 1000003)   # later options draw from a second stream: the scenes of old seeds keep their shape
    dominance = r2.random() < 0.35
    convex = r2.random() < 0.4
    s = scenes.Scene(f"fuzz_{seed}", gravity=(0.0, float(r.choice([-9.81, -10.0, -3.0])), 0.0))
    s.insert(RigidBodyBuilder.fixed().translation((0.0, -0.5, 0.0)),
             ColliderBuilder.cuboid(12.0, 0.5, 12.0).friction(float(r.uniform(0.0, 1.0))).restitution(float(r.choice([0.0, 0.0, 0.5]))))
    if r.random() < 0.3:   # a wall of parentless colliders
        for i in range(int(r.integers(1, 5))):
            s.colliders.insert(ColliderBuilder.cuboid(0.3, 1.5, 3.0).translation((4.0 + i * 0.7, 1.5, 0.0)))
    n = int(r.integers(2, 40))
    handles = []
    kinematic = []
    capsules = r.random() < 0.5
    dense = r.random() < 0.5    # dense: bodies start close together (many contacts); sparse: mostly free fall
    span = 1.5 if dense else 5.0
    for i in range(n):
        pos = (float(r.uniform(-span, span)), float(r.uniform(0.4, 3.0 if dense else 8.0)), float(r.uniform(-span, span)))
        kin = r.random() < 0.08
        b = (RigidBodyBuilder.kinematic_velocity_based() if r.random() < 0.5 else RigidBodyBuilder.kinematic_position_based()) if kin else RigidBodyBuilder.dynamic()
        b = b.translation(pos)
        if r.random() < 0.06:   # a fast body: CCD motion clamping against the ground / the walls
            b = b.linvel((float(r.uniform(-150.0, 150.0)), float(r.uniform(-200.0, 20.0)), float(r.uniform(-150.0, 150.0)))).ccd_enabled(bool(r.random() < 0.5))
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
        if dominance and r2.random() < 0.4:
            b = b.dominance_group(int(r2.choice([-128, -2, -1, 1, 3, 127])))
        sh = r.random()
        if sh < 0.5:
            c = ColliderBuilder.cuboid(*[float(x) for x in r.uniform(0.15, 0.7, 3)])
        elif sh < 0.8 or not capsules:
            c = ColliderBuilder.ball(float(r.uniform(0.15, 0.6)))
        else:
            c = [ColliderBuilder.capsule_x, ColliderBuilder.capsule_y, ColliderBuilder.capsule_z][int(r.integers(0, 3))](float(r.uniform(0.1, 0.7)), float(r.uniform(0.12, 0.4)))
        hull = convex and r2.random() < 0.4
        if hull:   # a random convex polyhedron (sharp or round) instead
            c = ColliderBuilder.round_convex_hull(r2.uniform(-0.6, 0.6, (int(r2.integers(4, 14)), 3)), float(r2.choice([0.0, 0.0, 0.05, 0.15])))
        c = c.density(float(r.choice([0.0, 0.5, 1.0, 10.0, 100.0]) if r.random() < 0.9 else 1.0))
        if hull and c._density == 0.0:
            c = c.density(1.0)
        c = c.friction(float(r.uniform(0.0, 1.2))).restitution(float(r.choice([0.0, 0.0, 0.0, 0.3, 0.9, 1.0])))
        if r.random() < 0.1:
            c = c.collision_groups(int(r.choice([1, 2, 3])), int(r.choice([1, 2, 3])))
        if r.random() < 0.1:
            c = c.translation(tuple(float(x) for x in r.uniform(-0.3, 0.3, 3)))
        if r.random() < 0.3:
            c = c.active_events(int(r.integers(1, 4))).contact_force_event_threshold(float(r.choice([0.0, 5.0, 50.0])))
        handles.append(s.insert(b, c))
        if kin:
            kinematic.append(handles[-1])
    generic = r.random() < 0.4
    for _ in range(int(r.integers(0, max(1, n // 4)))):
        a, b = (int(x) for x in r.choice(handles, 2, replace=False)) if n >= 2 else (handles[0], handles[0])
        if a == b:
            continue
        kind = r.integers(0, 4)
        axis = tuple(float(x) for x in r.choice([(1, 0, 0), (0, 0, 1), (0.6, 0.0, 0.8)]))
        j = SphericalJointBuilder() if kind == 0 else (FixedJointBuilder() if kind == 1 else (RevoluteJointBuilder(axis) if kind == 2 else PrismaticJointBuilder(axis)))
        if generic and kind != 1:   # limits / motors on the free axes
            for ax in ((3, 4, 5) if kind == 0 else ((3,) if kind == 2 else (0,))):
                if r.random() < 0.6:
                    lo = float(r.uniform(-1.5, 0.2))
                    j = j.limits(ax, lo, lo + float(r.uniform(0.1, 2.5)))
                m = r.random()
                if m < 0.3:
                    j = j.motor_velocity(ax, float(r.uniform(-3.0, 3.0)), float(r.uniform(0.5, 30.0)))
                elif m < 0.5:
                    j = j.motor_position(ax, float(r.uniform(-1.0, 1.0)), float(r.uniform(5.0, 200.0)), float(r.uniform(0.5, 20.0)))
                if m < 0.5 and r.random() < 0.4:
                    j = j.motor_max_force(ax, float(r.uniform(0.5, 50.0)))
                if m < 0.5 and r.random() < 0.3:
                    j = j.motor_model(ax, 1)
        j = j.local_anchor1(tuple(float(x) for x in r.uniform(-0.5, 0.5, 3))).local_anchor2(tuple(float(x) for x in r.uniform(-0.5, 0.5, 3)))
        if r.random() < 0.3:
            j = j.contacts_enabled(False)
        s.joints.insert(a, b, j)
    r3 = np.random.default_rng(seed + 2000003)   # third stream: coupled joint axes (spring / rope joints, cone limits)
    if r3.random() < 0.35 and n >= 2:
        from rapier_b200.sets import GenericJointBuilder, RopeJointBuilder, SpringJointBuilder
        for _ in range(int(r3.integers(1, 4))):
            a, b = (int(x) for x in r3.choice(handles, 2, replace=False))
            kind = int(r3.integers(0, 4))
            if kind == 0:
                j = SpringJointBuilder(float(r3.uniform(0.0, 2.0)), float(r3.uniform(5.0, 500.0)), float(r3.uniform(0.0, 20.0)))
                if r3.random() < 0.3:
                    j = j.motor_model(0, 0)
                if r3.random() < 0.3:
                    j = j.limits(0, 0.0, float(r3.uniform(0.5, 3.0)))
            elif kind == 1:
                j = RopeJointBuilder(float(r3.uniform(0.3, 3.0)))
            elif kind == 2:
                mask = int(r3.choice([0b101000, 0b011000, 0b110000]))
                first = 3 if mask & 0b001000 else 4
                lo = float(r3.uniform(-0.2, 0.2))
                j = GenericJointBuilder(0b000111).coupled_axes(mask).limits(first, lo, lo + float(r3.uniform(0.1, 1.5)))
            else:   # two coupled linear axes with a motor, the third one limited or locked
                j = GenericJointBuilder(int(r3.choice([0, 0b000010]))).coupled_axes(0b000101).motor_position(0, float(r3.uniform(0.0, 1.5)), float(r3.uniform(10.0, 200.0)), float(r3.uniform(0.5, 10.0)))
                if r3.random() < 0.5:
                    j = j.limits(1, -0.5, 0.5)
            j = j.local_anchor1(tuple(float(x) for x in r3.uniform(-0.5, 0.5, 3))).local_anchor2(tuple(float(x) for x in r3.uniform(-0.5, 0.5, 3)))
            s.joints.insert(a, b, j)
    if r3.random() < 0.3:   # sensors: fixed zones and auras on bodies (intersection events only)
        for _ in range(int(r3.integers(1, 4))):
            c = (ColliderBuilder.ball(float(r3.uniform(0.3, 1.5))) if r3.random() < 0.5 else
                 ColliderBuilder.cuboid(*(float(x) for x in r3.uniform(0.3, 2.0, 3)))).sensor(True).active_events(int(r3.integers(0, 2)))
            if r3.random() < 0.5:
                s.colliders.insert(c.translation(tuple(float(x) for x in r3.uniform(-3.0, 3.0, 3))))
            else:
                s.colliders.insert_with_parent(c.density(0.0), int(r3.choice(handles)))   # (massless: the body keeps its mass properties)
    if r3.random() < 0.3:   # compound bodies: extra massive parts at arbitrary offsets and orientations (general composite inertia)
        for _ in range(int(r3.integers(1, 4))):
            h = int(r3.choice(handles))
            if any(d.parent == h and d.shape == A.RB_SHAPE_CONVEX for d in s.colliders.descs):   # (no multi-collider bodies with hulls)
                continue
            c = (ColliderBuilder.ball(float(r3.uniform(0.15, 0.4))) if r3.random() < 0.3 else
                 ColliderBuilder.cuboid(*(float(x) for x in r3.uniform(0.1, 0.6, 3))))
            c = c.translation(tuple(float(x) for x in r3.uniform(-0.7, 0.7, 3))).density(float(r3.uniform(0.5, 3.0)))
            if r3.random() < 0.6:
                c = c.rotation(tuple(float(x) for x in r3.uniform(-1.0, 1.0, 3)))
            s.colliders.insert_with_parent(c, h)
    if r3.random() < 0.3:   # substep solve-groups: extra substeps for the islands of a few bodies
        for h in r3.choice(handles, min(len(handles), int(r3.integers(1, 4))), replace=False):
            s.bodies.descs[int(h)].flags |= int(r3.choice([1, 2, 5])) << A.RB_BODY_EXTRA_ITERS_SHIFT
    params = A.RbIntegrationParameters.default()
    if r.random() < 0.4:
        params.num_solver_iterations = int(r.choice([1, 2, 6]))
        params.num_internal_pgs_iterations = int(r.choice([1, 2]))
        params.num_internal_stabilization_iterations = int(r.choice([0, 1, 2]))
        params.warmstart_coefficient = float(r.choice([0.0, 0.5, 1.0]))
        params.friction_in_bias_pass = int(r.choice([0, 1]))
        params.contact_recycling = int(r.choice([0, 1]))
    if r.random() < 0.3:
        params.friction_model = 1
    if r.random() < 0.15:
        params.max_ccd_substeps = 0
    if r2.random() < 0.3:
        params.warmstart_joints = 1
    s.kinematic_position_based = [h for h in kinematic if s.bodies.descs[h].body_type == A.RB_BODY_KINEMATIC_POSITION_BASED]
    return s, params


def run(seed, steps=90, smem_floats=None):
    import emul_lib
    import oracle_lib
    from parity_util import compare_worlds, is_exact
    from rapier_b200.world import PhysicsWorld, RapierError
    if smem_floats is not None:
        os.environ["RB_EMU_COOP_SMEM_FLOATS"] = str(smem_floats)
    else:
        os.environ.pop("RB_EMU_COOP_SMEM_FLOATS", None)
    scene, params = random_scene(seed)
    w = PhysicsWorld(scene, integration_parameters=params, _lib=emul_lib.lib())
    try:
        o = oracle_lib.OracleWorld(scene, params=params)
    except (AssertionError, RuntimeError):   # a scene the oracle refuses (e.g. a degenerate random hull) must be refused by the library too
        try:
            w.step()
        except RapierError:
            return True, ""
        return False, f"seed {seed}: only the oracle refused the scene"
    import math
    # joint-set edits at run time (a fourth stream): insertions, removals, in-place updates of motors / limits
    from rapier_b200 import _abi as A
    r4 = np.random.default_rng(seed + 3000003)
    edits = {}
    dyn_handles = [h for h, d in enumerate(scene.bodies.descs) if d.body_type != A.RB_BODY_FIXED]
    if r4.random() < 0.35 and len(dyn_handles) >= 2:
        from rapier_b200.sets import RopeJointBuilder, SphericalJointBuilder, SpringJointBuilder, RevoluteJointBuilder
        w.reserve_joints(len(scene.joints) + 8, generic=True)   # (total slots, like rb_world_reserve)
        for _ in range(int(r4.integers(1, 5))):
            edits.setdefault(int(r4.integers(1, steps - 5)), []).append(int(r4.integers(0, 3)))
    live = list(range(len(scene.joints)))
    njoints = len(scene.joints)
    descs = {k: scene.joints.descs[k] for k in live}
    for i in range(steps):
        for kind in edits.get(i, ()):
            if kind == 0 and njoints - len(scene.joints) < 8:
                a, b = (int(x) for x in r4.choice(dyn_handles, 2, replace=False))
                jk = int(r4.integers(0, 4))
                jb = (SphericalJointBuilder() if jk == 0 else RopeJointBuilder(float(r4.uniform(0.5, 3.0))) if jk == 1 else
                      SpringJointBuilder(float(r4.uniform(0.2, 1.5)), float(r4.uniform(20.0, 300.0)), float(r4.uniform(1.0, 10.0))) if jk == 2 else
                      RevoluteJointBuilder((0.0, 0.0, 1.0)).limits(3, -0.5, 0.8).motor_velocity(3, float(r4.uniform(-2.0, 2.0)), 5.0))
                d = jb.local_anchor1(tuple(float(x) for x in r4.uniform(-0.4, 0.4, 3))).contacts_enabled(bool(r4.random() < 0.7)).build_desc(a, b)
                w.physics_pipeline.insert_joints([d]); o.insert_joints([d])
                live.append(njoints); descs[njoints] = d; njoints += 1
            elif kind == 1 and live:
                k = live.pop(int(r4.integers(0, len(live))))
                w.physics_pipeline.remove_joints([k]); o.remove_joints([k])
            elif kind == 2 and live:
                k = live[int(r4.integers(0, len(live)))]
                d = descs[k]
                for ax in range(6):
                    if d.motor_axes & (1 << ax):
                        d.motors[ax].target_vel = float(r4.uniform(-2.0, 2.0))
                    if d.limit_axes & (1 << ax):
                        d.limits[ax][1] = d.limits[ax][1] + 0.1
                wake = bool(r4.random() < 0.5)
                w.physics_pipeline.update_joints([k], [d], wake_up=wake); o.update_joints([k], [d], wake_up=wake)
        if scene.kinematic_position_based and i % 2 == 0:   # drive the position-based kinematic bodies along a curve
            t = 0.02 * (i + 1)
            poses = [(scene.bodies.descs[h].translation[0] + math.sin(t + h), scene.bodies.descs[h].translation[1] + 0.5 * t,
                      scene.bodies.descs[h].translation[2], 0.0, math.sin(0.3 * t), 0.0, math.cos(0.3 * t)) for h in scene.kinematic_position_based]
            w.set_next_kinematic_positions(scene.kinematic_position_based, poses)
            o.set_next_kinematic_positions(scene.kinematic_position_based, poses)
        try:
            w.step()
        except RapierError as e:   # a blow-up (random motors can be that violent) must quarantine the same bodies on both sides
            if "-5" not in str(e):
                raise
            o.step()
            qw = sorted(w.quarantine().tolist())
            if qw != sorted(o.quarantine().tolist()):
                return False, f"seed {seed} step {i}: quarantine lists differ"
            # (a NaN impulse times the zero inverse mass of a kinematic body is NaN too: quarantined bodies are disabled, stop driving them)
            scene.kinematic_position_based = [h for h in scene.kinematic_position_based if h not in qw]
            continue
        o.step()
        if len(o.quarantine()):
            return False, f"seed {seed} step {i}: only the oracle quarantined"
        if i % 15 == 14 or i < 2:
            d = compare_worlds(w, o)
            if not is_exact(d):
                return False, f"seed {seed} step {i}: {d}"
            if w.collision_events() != o.collision_events() or w.contact_force_events() != o.contact_force_events():
                return False, f"seed {seed} step {i}: event lists differ"
            ji_w, ji_o = w.debug_read("joint_impulses", np.float32), o.debug_read("joint_impulses", np.float32)
            if not ((ji_w.view(np.uint32) == ji_o.view(np.uint32)) | (np.isnan(ji_w) & np.isnan(ji_o))).all():   # (NaN payloads of a blown-up joint carry no meaning)
                return False, f"seed {seed} step {i}: joint impulses differ"
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
