"""Pins the oracle BEHAVIOURALLY against the reference's physical known-answer tests.
The reference holds no numeric fixture for this path that a re-implementation can reproduce (its
golden values are whole-binary bit hashes), so these are the pins available."""
import numpy as np
import pytest

import oracle_lib
from rapier_b200 import _abi as A
from rapier_b200 import scenes
from rapier_b200.sets import ColliderBuilder, RigidBodyBuilder


@pytest.mark.parametrize("shape", ["cuboid", "ball"])
@pytest.mark.parametrize("warmstart", [1.0, 0.5, 0.0])
def test_total_contact_impulse(shape, warmstart):
    """crates/rapier3d/tests/total_contact_impulse.rs:58-75: unit-mass cube / ball resting on a slab,
    sum of contact impulses = m g dt within 1 % after 300 steps, for several warm-start coefficients."""
    p = A.RbIntegrationParameters.default()
    p.warmstart_coefficient = warmstart
    w = oracle_lib.OracleWorld(scenes.box_on_ground(shape), params=p)
    w.step(300)
    total = float(w.contact_pairs()["impulses"].sum())
    expected = 9.81 / 60.0
    assert abs(total - expected) / expected < 0.01


def test_ball_rests_on_slab():
    """src/geometry/broad_phase_bvh/mod.rs:281-330: ball r=0.5 dropped from y=4 on a slab whose top is
    at y=0.5 rests at y = 1.0 +- 0.02 after 200 steps."""
    s = scenes.Scene("drop")
    s.insert(RigidBodyBuilder.fixed(), ColliderBuilder.cuboid(10.0, 0.5, 10.0))
    s.insert(RigidBodyBuilder.dynamic().translation((0.0, 4.0, 0.0)), ColliderBuilder.ball(0.5))
    w = oracle_lib.OracleWorld(s)
    w.step(200)
    pose, vel = w.body_states()
    assert abs(pose[1, 1] - 1.0) < 0.02
    assert np.abs(vel[1]).max() < 0.05


@pytest.mark.parametrize("e", [0.0, 0.3, 0.5, 0.8, 0.95])
def test_restitution(e):
    """crates/rapier3d/tests/issue_974_restitution.rs:11-90 (same scene): a ball dropped from 2 m
    recovers the e^2 fraction of the drop height +- 0.05; e = 0 does not bounce (< 0.02)."""
    s = scenes.Scene("bounce")
    s.insert(RigidBodyBuilder.fixed(), ColliderBuilder.cuboid(30.0, 0.1, 30.0).restitution(e))
    s.insert(RigidBodyBuilder.dynamic().translation((0.0, 2.3, 0.0)), ColliderBuilder.ball(0.2).restitution(e))
    w = oracle_lib.OracleWorld(s)
    touched, apex = False, 0.0
    for _ in range(400):
        w.step()
        y = float(w.body_states()[0][1, 1])
        if not touched and y < 0.35:
            touched = True
        if touched and y > apex:
            apex = y
    measured = (apex - 0.3) / 2.0
    if e == 0.0:
        assert measured < 0.02
    else:
        assert abs(measured - e * e) < 0.05


def test_speed_cap():
    """crates/rapier3d/tests/speed_cap.rs: linear speed capped at 400 m/s, angular at (pi/4)/dt."""
    s = scenes.Scene("caps", gravity=(0.0, 0.0, 0.0))
    s.insert(RigidBodyBuilder.dynamic().linvel((1000.0, 0.0, 0.0)).angvel((0.0, 500.0, 0.0)), ColliderBuilder.ball(0.5))
    w = oracle_lib.OracleWorld(s)
    w.step(1)
    _, vel = w.body_states()
    assert abs(np.linalg.norm(vel[0, :3]) - 400.0) < 1e-2
    assert abs(np.linalg.norm(vel[0, 3:]) - (np.pi / 4) * 60.0) < 1e-2


def test_gyroscopic_momentum():
    """crates/rapier3d/tests/gyroscopic.rs: the explicit gyroscopic term preserves |L| of a free body."""
    s = scenes.Scene("gyro", gravity=(0.0, 0.0, 0.0))
    s.insert(RigidBodyBuilder.dynamic().angvel((3.0, 0.5, 0.2)), ColliderBuilder.cuboid(0.1, 0.5, 1.0))
    w = oracle_lib.OracleWorld(s)
    mp = w.debug_read("body_mprops", np.float32).reshape(-1, 16)

    def momentum():
        pose, vel = w.body_states()
        q = pose[0, 3:]
        x, y, z, ww = q
        R = np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - z * ww), 2 * (x * z + y * ww)],
                      [2 * (x * y + z * ww), 1 - 2 * (x * x + z * z), 2 * (y * z - x * ww)],
                      [2 * (x * z - y * ww), 2 * (y * z + x * ww), 1 - 2 * (x * x + y * y)]])
        I = 1.0 / mp[0, 4:7]
        return np.linalg.norm(R @ (I * (R.T @ vel[0, 3:])))

    l0 = momentum()
    w.step(200)
    assert abs(momentum() - l0) / l0 < 1e-3


def test_joint_chain_bounded():
    """crates/rapier3d/tests/joint_stability.rs:19-30: joint nets stay bounded."""
    w = oracle_lib.OracleWorld(scenes.joint_grid(8))
    w.step(300)
    pose, vel = w.body_states()
    assert np.isfinite(pose).all()
    assert np.abs(pose[:, :3]).max() < 20.0
    # anchors of neighbouring balls stay within a few millimetres
    grid = pose[:, :3].reshape(8, 8, 3)
    d = np.linalg.norm(grid[:, 1:] - grid[:, :-1], axis=-1)
    assert np.abs(d - 1.0).max() < 0.05


def test_pyramid_is_stable():
    """examples3d/b3d_many_pyramids.rs scene semantics: a 10-base pyramid stays standing."""
    w = oracle_lib.OracleWorld(scenes.pyramids(1, 2, 10))
    p0, _ = w.body_states()
    w.step(300)
    p1, v1 = w.body_states()
    assert np.abs(p1[:, :3] - p0[:, :3]).max() < 0.05
    assert np.abs(v1).max() < 0.01
    c = w.counters()
    assert c["num_pairs"] == 2 * 145 and c["num_active_manifolds"] == 2 * 145


def test_manifold_face_face():
    """Cuboid-cuboid manifold contract: stacked unit cubes give 4 points, dist = gap,
    normal +y; separated beyond the prediction distance gives none."""
    pts, n1, n2 = oracle_lib.contact_manifold(A.RB_SHAPE_CUBOID, (0.5, 0.5, 0.5), A.RB_SHAPE_CUBOID, (0.5, 0.5, 0.5),
                                              (0.25, 1.01, 0.0), (0, 0, 0, 1))
    assert len(pts) == 4
    assert np.allclose(pts[:, 6], 0.01, atol=1e-6)
    assert np.allclose(n1, (0, 1, 0)) and np.allclose(n2, (0, -1, 0))
    pts, _, _ = oracle_lib.contact_manifold(A.RB_SHAPE_CUBOID, (0.5, 0.5, 0.5), A.RB_SHAPE_CUBOID, (0.5, 0.5, 0.5),
                                            (0.0, 1.5, 0.0), (0, 0, 0, 1))
    assert len(pts) == 0


def test_manifold_rotated_box_eight_points():
    """A box rotated 45 degrees about y on top of another clips to an octagon (8 raw points)."""
    s = np.sin(np.pi / 8)
    c = np.cos(np.pi / 8)
    pts, n1, _ = oracle_lib.contact_manifold(A.RB_SHAPE_CUBOID, (0.5, 0.5, 0.5), A.RB_SHAPE_CUBOID, (0.5, 0.5, 0.5),
                                             (0.0, 1.0, 0.0), (0, s, 0, c))
    assert len(pts) == 8
    assert np.allclose(np.abs(pts[:, 6]), 0.0, atol=1e-5)


def _rotate(q, v):
    x, y, z, w = q
    u = np.array([x, y, z])
    return v + 2.0 * np.cross(u, np.cross(u, v) + w * v)


@pytest.mark.parametrize("n", [8, 128])
def test_joint_anchor_with_offset_centre_of_mass(n):
    """crates/rapier3d/tests/issue_952_simd_joint_offset_com.rs:48-83: a joint anchored at the ORIGIN of a body whose
    centre of mass is offset must pin the origin (error < 1e-2 over 120 steps), for few and for many joints."""
    scene = scenes.offset_com_pendulums(n)
    w = oracle_lib.OracleWorld(scene)
    max_err = 0.0
    for _ in range(120):
        w.step()
        pose, _ = w.body_states()
        for body, base in scene.pendulums:
            max_err = max(max_err, float(np.linalg.norm(pose[body, :3] - np.array(base))))
    assert max_err < 1.0e-2, max_err
    # and it really is a pendulum: the body swung down about the joint
    pose, _ = w.body_states()
    assert pose[scene.pendulums[0][0], 3:].tolist() != [0.0, 0.0, 0.0, 1.0]


def _peak_stretch(world, scene, steps):
    peak = 0.0
    for _ in range(steps):
        world.step()
        pose, _ = world.body_states()
        for b1, b2, a1, a2 in scene.chain_joints:
            p1 = pose[b1, :3] + _rotate(pose[b1, 3:], np.array(a1))
            p2 = pose[b2, :3] + _rotate(pose[b2, 3:], np.array(a2))
            peak = max(peak, float(np.linalg.norm(p1 - p2)))
    return peak


def test_heavy_end_chain_stiffens_with_more_substeps():
    """crates/rapier3d/tests/substep_chain_high_mass_ratio.rs:147-161: 1000:1 chain over 300 steps; 16 additional
    solver iterations (= 20 substeps for the chain's group, the whole scene here) must cut the peak joint stretch
    at least 4x compared to the default 4."""
    scene = scenes.heavy_end_chain()
    baseline = _peak_stretch(oracle_lib.OracleWorld(scene), scene, 300)
    from rapier_b200._abi import RbIntegrationParameters
    params = RbIntegrationParameters.default()
    params.num_solver_iterations = 20
    elevated = _peak_stretch(oracle_lib.OracleWorld(scene, params=params), scene, 300)
    assert np.isfinite(baseline) and np.isfinite(elevated)
    assert elevated < baseline / 4.0, (baseline, elevated)


def _mass_twin_world(additional, tall):
    """Scenes of issue_78_additional_mass_rest.rs:9-40 (tilted cube dropped on a slab) and
    issue_666_additional_mass_inertia.rs:12-39 (fast tall cuboid on rough ground): the body's mass comes either from
    its collider's density or from RigidBody::additional_mass with a massless collider."""
    from rapier_b200.sets import ColliderBuilder, RigidBodyBuilder
    s = scenes.Scene("mass_twin")
    if tall:
        s.insert(RigidBodyBuilder.fixed().translation((0.0, -0.1, 0.0)), ColliderBuilder.cuboid(100.1, 0.1, 100.1).friction(0.5))
        body = RigidBodyBuilder.dynamic().translation((-10.0, 6.0, 0.0)).linvel((20.0, 0.0, 0.0))
        he, mass = (0.2, 5.0, 1.5), 0.5
    else:
        s.insert(RigidBodyBuilder.fixed().translation((0.0, -0.5, 0.0)), ColliderBuilder.cuboid(5.0, 0.5, 5.0).friction(0.5))
        body = RigidBodyBuilder.dynamic().translation((0.0, 1.0, 0.0)).rotation((0.0, 0.0, np.pi / 4 * 0.9))
        he, mass = (0.5, 0.5, 0.5), 100.0
    volume = 8.0 * he[0] * he[1] * he[2]
    col = ColliderBuilder.cuboid(*he).friction(0.5).restitution(0.0)
    if additional:
        body = body.additional_mass(mass)
        col = col.density(0.0)
    else:
        col = col.density(mass / volume)    # ColliderBuilder::mass(m)
    h = s.insert(body, col)
    return s, h


def test_additional_mass_body_rests_like_density_twin():
    """crates/rapier3d/tests/issue_78_additional_mass_rest.rs:58-78 (sleeping is out of scope: "at rest" = slow)."""
    rest_y, rest_step = [], []
    for additional in (False, True):
        scene, h = _mass_twin_world(additional, tall=False)
        w = oracle_lib.OracleWorld(scene)
        at = None
        for step in range(600):
            w.step()
            pose, vel = w.body_states()
            assert pose[h, 1] > -0.5, f"tunnelled at step {step}"
            if at is None and step > 10 and np.abs(vel[h]).max() < 1e-2:
                at = step
        assert at is not None, "never came to rest"
        rest_y.append(float(pose[h, 1])); rest_step.append(at)
    assert abs(rest_y[0] - rest_y[1]) < 0.1, rest_y
    assert rest_step[1] < max(rest_step[0], 1) * 4, rest_step


def test_additional_mass_body_topples_like_density_twin():
    """crates/rapier3d/tests/issue_666_additional_mass_inertia.rs:53-83: additional mass on a massless collider must
    get an angular inertia from the shape, so the fast tall cuboid topples (rotation > 0.5 rad within 200 steps)."""
    angles = []
    for additional in (False, True):
        scene, h = _mass_twin_world(additional, tall=True)
        w = oracle_lib.OracleWorld(scene)
        w.step()
        if additional:
            mp = w.debug_read("body_mprops", np.float32).reshape(-1, 16)
            assert np.abs(mp[h, 4:7]).max() > 0.0, "additional_mass must produce a non-zero angular inertia"
        mx = 0.0
        for _ in range(200):
            w.step()
            pose, _ = w.body_states()
            mx = max(mx, 2.0 * float(np.arccos(min(1.0, abs(pose[h, 6])))))
        angles.append(mx)
    assert angles[0] > 0.5 and angles[1] > 0.5, angles


def test_no_fixed_fixed_pairs():
    """crates/rapier3d/tests/broad_phase_pair_filter.rs:9-29: two overlapping parentless (fixed) colliders never form
    a pair; the same colliders form one as soon as one of them hangs on a dynamic body."""
    from rapier_b200.sets import ColliderBuilder, RigidBodyBuilder
    import emul_lib
    from rapier_b200.world import PhysicsWorld
    s = scenes.Scene("fixed_fixed")
    s.colliders.insert(ColliderBuilder.cuboid(1.0, 1.0, 1.0))
    s.colliders.insert(ColliderBuilder.cuboid(1.0, 1.0, 1.0).translation((0.5, 0.5, 0.0)))
    w = oracle_lib.OracleWorld(s)
    w.step(3)
    assert w.counters()["num_pairs"] == 0
    k = PhysicsWorld(s, _lib=emul_lib.lib())   # the kernels' logic (host emulation) applies the same filter
    k.step(3)
    assert k.counters()["num_pairs"] == 0
    s2 = scenes.Scene("fixed_dynamic")
    s2.colliders.insert(ColliderBuilder.cuboid(1.0, 1.0, 1.0))
    s2.insert(RigidBodyBuilder.dynamic().translation((0.5, 0.5, 0.0)), ColliderBuilder.cuboid(1.0, 1.0, 1.0))
    w2 = oracle_lib.OracleWorld(s2)
    w2.step(3)
    assert w2.counters()["num_pairs"] == 1


def _offset_com_locked_body():
    from rapier_b200.sets import ColliderBuilder, RigidBodyBuilder
    s = scenes.Scene("locked_rotations_offset_com", gravity=(0.0, 0.0, 0.0))
    b = s.bodies.insert(RigidBodyBuilder.dynamic().lock_rotations())
    s.colliders.insert_with_parent(ColliderBuilder.cuboid(0.5, 0.5, 0.5).translation((1.0, 0.0, 0.0)), b)
    return s, b


def _world_com(pose, local_com):
    return pose[:3] + _rotate(pose[3:], np.array(local_com))


@pytest.mark.parametrize("which", ["oracle", "emulated_kernels"])
def test_locked_rotations_pivot_about_centre_of_mass(which):
    """crates/rapier3d/tests/issue_309_locked_rotations_offset_com.rs:8-38: a rotation-locked body whose collider is
    offset from its origin, spun by a manually set angular velocity every step, must rotate about its centre of mass
    (the world-space centre of mass moves < 1e-3 over 100 steps)."""
    scene, b = _offset_com_locked_body()
    if which == "oracle":
        w = oracle_lib.OracleWorld(scene)
    else:
        import emul_lib
        from rapier_b200.world import PhysicsWorld
        w = PhysicsWorld(scene, _lib=emul_lib.lib())
    w.step()
    pose, _ = w.body_states()
    com0 = _world_com(pose[b], (1.0, 0.0, 0.0))
    for i in range(100):
        vel = np.array([[0.0, 0.0, 0.0, 0.0, 0.0, 3.0]], np.float32)
        if which == "oracle":
            w.set_body_states([b], vel6=vel)
        else:
            w.physics_pipeline.set_body_states([b], vel6=vel)
        w.step()
        pose, _ = w.body_states()
        assert np.linalg.norm(_world_com(pose[b], (1.0, 0.0, 0.0)) - com0) < 1.0e-3, i
    assert abs(pose[b, 6]) < 0.999, "the body must actually have rotated"


def test_set_rotation_does_not_inject_motion():
    """issue_309_locked_rotations_offset_com.rs:40-72: teleporting the orientation about the body origin every step
    must not make the body drift or acquire velocity."""
    scene, b = _offset_com_locked_body()
    w = oracle_lib.OracleWorld(scene)
    for i in range(100):
        a = 0.05 * i
        pose = np.array([[0.0, 0.0, 0.0, 0.0, 0.0, np.sin(a / 2), np.cos(a / 2)]], np.float32)
        p_now, _ = w.body_states()
        pose[0, :3] = p_now[b, :3]
        w.set_body_states([b], pose7=pose)
        w.step()
        p, v = w.body_states()
        assert np.linalg.norm(p[b, :3]) < 1.0e-4 and np.linalg.norm(v[b, :3]) < 1.0e-4, (i, p[b], v[b])


def test_thread_count_determinism():
    """crates/rapier3d/tests/thread_count_determinism.rs:94-150 (scene: jittered 12x3x12 pile + a swinging joint
    chain; the golden hash itself needs the reference binary): every thread count must give the same bits."""
    ref = None
    for threads in (1, 3, 8):
        w = oracle_lib.OracleWorld(scenes.box_pile(12, 3, 12), threads=threads)
        w.step(60)
        pose, vel = w.body_states()
        bits = (pose.view(np.uint32).copy(), vel.view(np.uint32).copy())
        if ref is None:
            ref = bits
        else:
            assert (bits[0] == ref[0]).all() and (bits[1] == ref[1]).all(), threads


def test_cpu_baseline_build_tracks_the_bit_exact_oracle():
    """bench.py's CPU arm uses the oracle sources rebuilt -O3 -march=native with FMA contraction (oracle/Makefile
    `fast`): a BASELINE, never a checker.  It must stay within 1e-4 of the bit-exact oracle after 10 steps."""
    scene = scenes.box_pile(4, 4, 5)
    a = oracle_lib.OracleWorld(scene)
    b = oracle_lib.OracleWorld(scene, fast=True, threads=2)
    a.step(10)
    b.step(10)
    pa, va = a.body_states()
    pb, vb = b.body_states()
    assert np.abs(pa - pb).max() <= 1e-4 and np.abs(va - vb).max() <= 1e-3


# ---- sleeping: crates/rapier3d/tests/whole_island_sleep.rs ----
def _stack_world(n, restless=False):
    s = scenes.Scene("stack")
    s.insert(RigidBodyBuilder.fixed().translation((0.0, -0.5, 0.0)), ColliderBuilder.cuboid(100.0, 0.5, 100.0))
    stack = [s.insert(RigidBodyBuilder.dynamic().translation((0.0, 0.5 + i, 0.0)), ColliderBuilder.cuboid(0.5, 0.5, 0.5)) for i in range(n)]
    r = None
    if restless:
        r = s.insert(RigidBodyBuilder.dynamic().translation((0.0, n + 0.5, 0.0)).can_sleep(False), ColliderBuilder.cuboid(0.5, 0.5, 0.5))
    return s, stack, r


def whole_island_blocks_partial_sleep(make_world):
    """whole_island_sleep.rs:40-83: a can_sleep(false) body atop a stack keeps the WHOLE stack awake; once removed the
    stack sleeps; waking one body wakes the island as a unit."""
    s, stack, restless = _stack_world(6, restless=True)
    w = make_world(s)
    w.step(240)
    assert w.sleeping().sum() == 0
    w.remove_bodies([restless]) if hasattr(w, "remove_bodies") else w.remove(restless)
    w.step(240)
    sl = w.sleeping()
    assert all(sl[b] == 1 for b in stack)
    w.wake_up([stack[0]])
    assert w.sleeping().sum() == 0


def sleeping_stack_is_woken_by_an_impact(make_world):
    """A resting stack falls asleep (no constraint is solved any more); a ball dropped on it wakes every body."""
    s, stack, _ = _stack_world(4)
    ball = s.insert(RigidBodyBuilder.dynamic().translation((0.2, 40.0, 0.0)).can_sleep(False), ColliderBuilder.ball(0.4))
    w = make_world(s)
    w.step(100)
    sl = w.sleeping()
    assert all(sl[b] == 1 for b in stack) and sl[ball] == 0
    assert w.counters()["num_active_manifolds"] == 0
    for _ in range(120):
        w.step(1)
        if w.sleeping()[stack[0]] == 0:
            break
    sl = w.sleeping()
    assert all(sl[b] == 0 for b in stack), "the impact must wake the whole island"


def test_whole_island_sleep_oracle():
    whole_island_blocks_partial_sleep(lambda s: oracle_lib.OracleWorld(s))
    sleeping_stack_is_woken_by_an_impact(lambda s: oracle_lib.OracleWorld(s))


# ---- quarantine: src/pipeline/physics_pipeline/quarantine.rs:295-352 ----
def nan_force_is_quarantined(make_world, expect_error):
    """`nan_force_is_quarantined_at_end_of_step` + `nan_spread_through_contacts_is_contained`: a NaN force becomes a NaN
    velocity mid-step; the end-of-step chokepoint rolls the body back to its last valid pose, zeroes its velocity,
    disables it and reports it; whatever it infected through contacts is quarantined too, nothing else is corrupted."""
    s = scenes.Scene("nan")
    s.insert(RigidBodyBuilder.fixed().translation((0.0, -0.5, 0.0)), ColliderBuilder.cuboid(50.0, 0.5, 50.0))
    poisoned = s.insert(RigidBodyBuilder.dynamic().translation((0.0, 3.0, 0.0)), ColliderBuilder.ball(0.5))
    s.bodies.descs[poisoned].user_force[:] = (float("nan"), 0.0, 0.0)
    bottom = s.insert(RigidBodyBuilder.dynamic().translation((5.0, 0.5, 0.0)), ColliderBuilder.ball(0.5))
    s.bodies.descs[bottom].user_torque[:] = (0.0, float("nan"), 0.0)
    top = s.insert(RigidBodyBuilder.dynamic().translation((5.0, 1.5, 0.0)), ColliderBuilder.ball(0.5))
    healthy = s.insert(RigidBodyBuilder.dynamic().translation((-6.0, 0.5, 0.0)), ColliderBuilder.cuboid(0.5, 0.5, 0.5))
    w = make_world(s)
    if expect_error:
        with pytest.raises(Exception, match="-5"):
            w.step()
    else:
        w.step()
    q = w.quarantine().tolist()
    assert poisoned in q and bottom in q and healthy not in q
    pose, vel = w.body_states()
    assert np.allclose(pose[poisoned], [0.0, 3.0, 0.0, 0.0, 0.0, 0.0, 1.0]) and (vel[poisoned] == 0).all()
    for _ in range(10):
        w.step()
        assert len(w.quarantine()) == 0 or set(w.quarantine().tolist()) <= {top}
        pose, vel = w.body_states()
        alive = [b for b in (top, healthy) if b not in q]
        assert np.isfinite(pose[alive]).all() and np.isfinite(vel[alive]).all()


def test_quarantine_oracle():
    nan_force_is_quarantined(lambda s: oracle_lib.OracleWorld(s), expect_error=False)


# ---- FrictionModel::Coulomb (contact_with_coulomb_friction.rs; integration_parameters.rs:16-30) --------------------
def _coulomb_params():
    p = A.RbIntegrationParameters.default()
    p.friction_model = 1
    return p


def coulomb_friction_cone(make_world):
    """Coulomb's law per contact point (contact_with_coulomb_friction.rs:659-671: tangent impulse capped at mu * lambda_k):
    a unit cube (mu = 0.5, m = 1) pushed sideways on a slab stays put for F < mu m g and accelerates at (F - mu m g) / m
    above it.  The total contact impulse is m g dt as under the twist model (total_contact_impulse.rs:58-75)."""
    for force, slides in ((3.0, False), (8.0, True)):
        s = scenes.box_on_ground("cuboid")
        s.bodies.descs[1].user_force[:] = (force, 0.0, 0.0)
        w = make_world(s)
        w.step(30)
        pose0, vel0 = w.body_states()
        w.step(30)
        pose1, vel1 = w.body_states()
        if not slides:
            assert abs(vel1[1, 0]) < 1e-3 and abs(pose1[1, 0] - pose0[1, 0]) < 1e-3
        else:
            acc = (vel1[1, 0] - vel0[1, 0]) / 0.5
            assert abs(acc - (force - 0.5 * 9.81)) < 0.5, acc   # friction force within 10 % of mu m g
        if hasattr(w, "contact_pairs") and not slides:
            total = float(w.contact_pairs()["impulses"].sum())
            assert abs(total - 9.81 / 60.0) / (9.81 / 60.0) < 0.01


def test_coulomb_friction_cone_oracle():
    coulomb_friction_cone(lambda s: oracle_lib.OracleWorld(s, params=_coulomb_params()))


def test_coulomb_differs_from_twist_only_in_friction():
    """Both models share the normal rows: a resting pyramid stays put under either; a spinning box on the ground is
    braked by per-point friction under Coulomb and by the twist row under the simplified model (different trajectories)."""
    s = scenes.Scene("spin")
    s.insert(RigidBodyBuilder.fixed().translation((0.0, -0.5, 0.0)), ColliderBuilder.cuboid(10.0, 0.5, 10.0))
    s.insert(RigidBodyBuilder.dynamic().translation((0.0, 0.5, 0.0)).angvel((0.0, 6.0, 0.0)), ColliderBuilder.cuboid(0.5, 0.5, 0.5))
    wc = oracle_lib.OracleWorld(s, params=_coulomb_params())
    wt = oracle_lib.OracleWorld(s)
    wc.step(8); wt.step(8)
    (_, vc), (_, vt) = wc.body_states(), wt.body_states()
    assert 0.0 < vc[1, 4] < 6.0 and 0.0 < vt[1, 4] < 6.0 and vc[1, 4] != vt[1, 4]   # both brake the spin, differently
    wc.step(200); wt.step(200)
    assert abs(wc.body_states()[1][1, 4]) < 1e-2 and abs(wt.body_states()[1][1, 4]) < 1e-2
    p = scenes.pyramids(1, 1, 8)
    w = oracle_lib.OracleWorld(p, params=_coulomb_params())
    start, _ = w.body_states()
    w.step(200)
    end, _ = w.body_states()
    assert np.abs(end[:, :3] - start[:, :3]).max() < 0.02


# ---- CCD motion clamping (src/dynamics/ccd; crates/rapier3d/tests/ccd_default_vs_fixed.rs, issue_217_ccd_large_dt_hitch.rs) ----
def _thin_wall_scene(with_second_body=False):
    s = scenes.Scene("ccd_wall", gravity=(0.0, 0.0, 0.0))
    if not with_second_body:
        s.insert(RigidBodyBuilder.fixed(), ColliderBuilder.cuboid(0.05, 5.0, 5.0))                      # insert_thin_fixed_wall
    s.insert(RigidBodyBuilder.dynamic().translation((-3.0, 0.0, 0.0)).linvel((200.0, 0.0, 0.0)), ColliderBuilder.cuboid(0.1, 0.1, 0.1))
    if with_second_body:
        s.insert(RigidBodyBuilder.dynamic().translation((3.0, 0.0, 0.0)).linvel((-200.0, 0.0, 0.0)), ColliderBuilder.cuboid(0.1, 0.1, 0.1))
    return s


def ccd_default_tier(make_world):
    """ccd_default_vs_fixed.rs: default_ccd_vs_fixed_no_tunnel (a 0.2 m cube at 200 m/s is stopped on the near side of a
    0.1 m fixed wall, :98-113), global_ccd_off_tunnels (max_ccd_substeps = 0: it tunnels, :190-205),
    default_tier_ignores_dynamic (two fast dynamic bodies pass through each other, :118-152)."""
    w = make_world(_thin_wall_scene(), None)
    w.step(120)
    assert w.body_states()[0][1, 0] < 0.0
    p = A.RbIntegrationParameters.default()
    p.max_ccd_substeps = 0
    w = make_world(_thin_wall_scene(), p)
    w.step(60)
    assert w.body_states()[0][1, 0] > 1.0
    w = make_world(_thin_wall_scene(with_second_body=True), None)
    w.step(5)
    pose, _ = w.body_states()
    assert pose[0, 0] > 0.0 and pose[1, 0] < 0.0


def ccd_bullet_still_hits_dynamic(make_world):
    """ccd_default_vs_fixed.rs:157-186: a ccd_enabled body (a bullet) sweeps against a dynamic target: after 60 steps the
    target has been pushed along +X and the bullet is still behind it."""
    s = scenes.Scene("ccd_bullet", gravity=(0.0, 0.0, 0.0))
    bullet = s.insert(RigidBodyBuilder.dynamic().translation((-3.0, 0.0, 0.0)).linvel((200.0, 0.0, 0.0)).ccd_enabled(True), ColliderBuilder.cuboid(0.1, 0.1, 0.1))
    target = s.insert(RigidBodyBuilder.dynamic().translation((0.0, 0.0, 0.0)), ColliderBuilder.cuboid(0.2, 0.2, 0.2))
    w = make_world(s, None)
    w.step(60)
    pose, _ = w.body_states()
    assert pose[target, 0] > 0.05 and pose[bullet, 0] < pose[target, 0], pose[:, 0]


def ccd_large_dt_no_mid_air_hitch(make_world):
    """issue_217_ccd_large_dt_hitch.rs:62-121: ball (r = 0.5, 20 m/s, dt = 0.25: 5 m per step) against a thin wall whose
    near face is at x = 12.2: full-speed advance while far, then adjacent to the wall (never through, never frozen short
    of it), at rest from the fifth step on."""
    p = A.RbIntegrationParameters.default()
    p.dt = 0.25
    s = scenes.Scene("ccd_hitch", gravity=(0.0, 0.0, 0.0))
    s.insert(RigidBodyBuilder.fixed(), ColliderBuilder.cuboid(0.05, 5.0, 5.0).translation((12.25, 0.0, 0.0)))
    s.insert(RigidBodyBuilder.dynamic().linvel((20.0, 0.0, 0.0)), ColliderBuilder.ball(0.5))
    w = make_world(s, p)
    contact_x, travel, prev_x = 12.2 - 0.5, 5.0, 0.0
    for i in range(10):
        w.step()
        pose, vel = w.body_states()
        x, vx = float(pose[1, 0]), float(vel[1, 0])
        if prev_x + travel < contact_x - 0.5:
            assert abs(x - (prev_x + travel)) < 1.0e-3, (i, x)
        else:
            assert contact_x - 0.35 < x < contact_x + 0.01, (i, x)
            if i >= 4:
                assert abs(vx) < 0.1, (i, vx)
        prev_x = x


def test_ccd_oracle():
    mk = lambda s, p: oracle_lib.OracleWorld(s, params=p)
    ccd_default_tier(mk)
    ccd_large_dt_no_mid_air_hitch(mk)
    ccd_bullet_still_hits_dynamic(mk)


# ---- events (EventHandler; crates/rapier3d/tests/contact_force_event_first_tick.rs) -------------------------------------
def contact_force_event_started_marks_threshold_crossings(make_world):
    """contact_force_event_first_tick.rs:53-170: a 1 kg ball resting on a slab (threshold 30 N) emits one
    CollisionEvent::Started and no force event; pressed with 100 N it emits force events whose FIRST has started = true
    and the following ones false; released, the events stop; pressed again, a new episode starts with started = true."""
    s = scenes.Scene("force_events", gravity=(0.0, -9.81, 0.0))
    s.colliders.insert(ColliderBuilder.cuboid(10.0, 0.5, 10.0).translation((0.0, -0.5, 0.0)))
    ball = s.insert(RigidBodyBuilder.dynamic().translation((0.0, 0.5, 0.0)).additional_mass(1.0).can_sleep(False),
                    ColliderBuilder.ball(0.5).density(0.0).active_events(A.RB_EVENT_COLLISION | A.RB_EVENT_CONTACT_FORCE)
                    .contact_force_event_threshold(30.0))
    w = make_world(s)
    started_steps, force_events = [], []

    def run(rng, press):
        for i in rng:
            w.set_body_forces([ball], force3=[(0.0, -100.0 if press else 0.0, 0.0)])
            w.step()
            started_steps.extend(i for (_, _, st, _) in w.collision_events() if st)
            force_events.extend((i, bool(e["started"])) for e in w.contact_force_events())

    run(range(0, 100), False)
    assert len(started_steps) == 1 and not force_events
    run(range(100, 160), True)
    assert force_events and force_events[0][1] and force_events[0][0] >= 100 and force_events[0][0] > started_steps[0] + 50
    assert all(not first for _, first in force_events[1:]) and len(force_events) > 10
    run(range(160, 165), False)
    assert all(not first for _, first in force_events[1:])
    n_after = len(force_events)
    run(range(165, 220), False)
    assert len(force_events) == n_after
    run(range(220, 280), True)
    episode = force_events[n_after:]
    assert episode and episode[0][1] and all(not first for _, first in episode[1:])


def test_contact_force_events_oracle():
    contact_force_event_started_marks_threshold_crossings(lambda s: oracle_lib.OracleWorld(s))


# ---- kinematic bodies (RigidBodyType::Kinematic*; rigid_body_components.rs:20-46) ----------------------------------------
def kinematic_bodies(make_world):
    """A velocity-based kinematic platform carries a box riding on it at its own speed (its velocity enters the contact rows: a
    platform read as static would brake the box to rest) and is not slowed by it; a
    position-based one lifts a box to the targets it is given and ends exactly on them (worker.rs:836-842); kinematic
    bodies collide with neither fixed nor other kinematic bodies (ActiveCollisionTypes::default()); a kinematic wall is
    not a CCD target of the default tier (ccd_default_vs_fixed.rs: kinematic_not_a_default_target)."""
    s = scenes.Scene("kinematic")
    plat = s.insert(RigidBodyBuilder.kinematic_velocity_based().translation((0.0, 0.0, 0.0)).linvel((1.5, 0.0, 0.0)), ColliderBuilder.cuboid(4.0, 0.25, 2.0).friction(1.0))
    box = s.insert(RigidBodyBuilder.dynamic().translation((0.0, 0.76, 0.0)).linvel((1.5, 0.0, 0.0)), ColliderBuilder.cuboid(0.5, 0.5, 0.5).friction(1.0))
    lift = s.insert(RigidBodyBuilder.kinematic_position_based().translation((10.0, 0.0, 0.0)), ColliderBuilder.cuboid(1.0, 0.25, 1.0))
    cargo = s.insert(RigidBodyBuilder.dynamic().translation((10.0, 0.76, 0.0)), ColliderBuilder.cuboid(0.5, 0.5, 0.5))
    s.insert(RigidBodyBuilder.fixed().translation((10.0, 0.3, 0.0)), ColliderBuilder.cuboid(0.3, 0.3, 0.3))        # inside the lift: no pair
    s.insert(RigidBodyBuilder.kinematic_velocity_based().translation((0.5, -0.1, 0.0)).linvel((1.5, 0.0, 0.0)), ColliderBuilder.ball(0.3))   # inside the platform: no pair
    w = make_world(s)
    for i in range(120):
        w.set_next_kinematic_positions([lift], [(10.0, 0.01 * (i + 1), 0.0, 0.0, 0.0, 0.0, 1.0)])
        w.step()
    pose, vel = w.body_states()
    assert abs(pose[plat, 0] - 1.5 * 2.0) < 1e-3 and abs(vel[plat, 0] - 1.5) < 1e-6 and abs(pose[plat, 1]) < 1e-6   # unperturbed
    assert abs(vel[box, 0] - 1.5) < 0.05 and abs(pose[box, 0] - pose[plat, 0]) < 0.2 and abs(pose[box, 1] - 0.75) < 0.02   # carried along
    assert pose[lift, 1] == np.float32(0.01 * 120) and abs(vel[lift, 1] - 0.6) < 1e-3                              # exactly on its target
    assert abs(pose[cargo, 1] - (1.2 + 0.75)) < 0.03 and abs(vel[cargo, 1] - 0.6) < 0.05                            # lifted
    pairs = w.contact_pairs()["colliders"]
    got = set(map(tuple, pairs.tolist()))
    assert (0, 1) in got and (2, 3) in got and (0, 5) not in got and (2, 4) not in got, pairs   # no kinematic-kinematic / kinematic-fixed pair
    w.step(30)                                           # no new target: the lift stops (next_position == position)
    pose2, vel2 = w.body_states()
    assert pose2[lift, 1] == pose[lift, 1] and vel2[lift, 1] == 0.0


def test_kinematic_bodies_oracle():
    kinematic_bodies(lambda s: oracle_lib.OracleWorld(s))


def moving_kinematic_wakes_jointed_dynamic(make_world):
    """issue_287_kinematic_wakes_jointed_dynamic.rs: a ball hanging from a position-based kinematic body by a revolute
    joint falls asleep; driving the kinematic body sideways wakes it and drags it along."""
    from rapier_b200.sets import RevoluteJointBuilder
    s = scenes.Scene("issue_287")
    kin = s.bodies.insert(RigidBodyBuilder.kinematic_position_based())
    dyn = s.insert(RigidBodyBuilder.dynamic().translation((0.0, -2.0, 0.0)), ColliderBuilder.ball(0.5))
    s.joints.insert(kin, dyn, RevoluteJointBuilder((1.0, 0.0, 0.0)).local_anchor1((0.0, 0.0, 0.0)).local_anchor2((0.0, 2.0, 0.0)))
    w = make_world(s)
    steps = 0
    while not w.sleeping()[dyn]:
        w.step()
        steps += 1
        assert steps < 2000, "dynamic body never fell asleep"
    woke, x = False, 0.0
    for _ in range(200):
        x += 0.05
        w.set_next_kinematic_positions([kin], [(x, 0.0, 0.0, 0.0, 0.0, 0.0, 1.0)])
        w.step()
        woke = woke or not w.sleeping()[dyn]
    assert woke
    assert abs(float(w.body_states()[0][dyn, 0]) - x) < 2.0


def test_kinematic_wakes_jointed_dynamic_oracle():
    moving_kinematic_wakes_jointed_dynamic(lambda s: oracle_lib.OracleWorld(s))
    # ccd_default_vs_fixed.rs: kinematic_not_a_default_target -- a default-tier fast body tunnels through a kinematic wall
    s = scenes.Scene("ccd_kinematic_wall", gravity=(0.0, 0.0, 0.0))
    s.insert(RigidBodyBuilder.kinematic_position_based(), ColliderBuilder.cuboid(0.05, 5.0, 5.0))
    s.insert(RigidBodyBuilder.dynamic().translation((-3.0, 0.0, 0.0)).linvel((200.0, 0.0, 0.0)), ColliderBuilder.cuboid(0.1, 0.1, 0.1))
    w = oracle_lib.OracleWorld(s)
    w.step(60)
    assert w.body_states()[0][1, 0] > 1.0


# ---- joint limits and motors (joint_constraint_helper.rs:166-626; crates/rapier3d/tests/issue_499_angular_limits.rs) -----------
def _settled_angle(make_world, drive, limits_deg, direction):
    """issue_499_angular_limits.rs:32-100: a bar on a revolute joint about Z with angular limits, driven into a limit
    by a velocity motor (5 rad/s, factor 20) or by a constant torque; the unwrapped angle after 600 steps, degrees."""
    import math
    from rapier_b200.sets import RevoluteJointBuilder
    s = scenes.Scene("issue_499", gravity=(0.0, 0.0, 0.0))
    b1 = s.bodies.insert(RigidBodyBuilder.fixed())
    b2 = s.insert(RigidBodyBuilder.dynamic().translation((1.0, 0.0, 0.0)).angular_damping(3.0).can_sleep(False), ColliderBuilder.cuboid(0.5, 0.1, 0.1))
    j = RevoluteJointBuilder((0.0, 0.0, 1.0)).local_anchor1((0.0, 0.0, 0.0)).local_anchor2((-1.0, 0.0, 0.0))
    j = j.limits(3, math.radians(limits_deg[0]), math.radians(limits_deg[1]))
    if drive == "motor":
        j = j.motor_velocity(3, direction * 5.0, 20.0)
    s.joints.insert(b1, b2, j)
    w = make_world(s)
    if drive == "torque":
        w.set_body_forces([b2], torque3=[(0.0, 0.0, direction * 0.1)])
    unwrapped = prev = 0.0
    for _ in range(600):
        w.step()
        q = w.body_states()[0][b2, 3:7]
        ang = 2.0 * math.atan2(float(q[2]), float(q[3]))
        delta = ang - prev
        if delta > math.pi:
            delta -= 2.0 * math.pi
        elif delta < -math.pi:
            delta += 2.0 * math.pi
        unwrapped += delta
        prev = ang
    return math.degrees(unwrapped)


def angular_limits_are_reached(make_world, cases=((-45.0, 45.0), (0.0, 90.0), (0.0, 270.0), (45.0, 315.0), (-170.0, -10.0))):
    for lim in cases:
        for drive in ("motor", "torque"):
            hi = _settled_angle(make_world, drive, lim, 1.0)
            lo = _settled_angle(make_world, drive, lim, -1.0)
            assert abs(hi - lim[1]) < 2.0 and abs(lo - lim[0]) < 2.0, (lim, drive, lo, hi)
    assert _settled_angle(make_world, "motor", (-200.0, 200.0), 1.0) > 360.0   # wider than a turn: free


def prismatic_limits_and_position_motor(make_world):
    """A body on a prismatic joint along X under gravity along +X stops at its upper limit (limit_linear, :166-207); a
    position motor (spring, motor_linear :285-330) holds another one near its target against gravity; a velocity motor
    with a small max force cannot lift a third one (impulse bounds)."""
    from rapier_b200.sets import PrismaticJointBuilder
    s = scenes.Scene("prismatic", gravity=(3.0, 0.0, 0.0))
    base = s.bodies.insert(RigidBodyBuilder.fixed())
    a = s.insert(RigidBodyBuilder.dynamic().translation((0.0, 0.0, 0.0)).can_sleep(False), ColliderBuilder.cuboid(0.2, 0.2, 0.2))
    s.joints.insert(base, a, PrismaticJointBuilder((1.0, 0.0, 0.0)).limits(0, -0.5, 1.5))
    b = s.insert(RigidBodyBuilder.dynamic().translation((0.0, 2.0, 0.0)).can_sleep(False), ColliderBuilder.cuboid(0.2, 0.2, 0.2))
    s.joints.insert(base, b, PrismaticJointBuilder((1.0, 0.0, 0.0)).local_anchor1((0.0, 2.0, 0.0)).motor_position(0, -0.7, 400.0, 40.0))
    c = s.insert(RigidBodyBuilder.dynamic().translation((0.0, 4.0, 0.0)).can_sleep(False), ColliderBuilder.cuboid(0.2, 0.2, 0.2))
    s.joints.insert(base, c, PrismaticJointBuilder((1.0, 0.0, 0.0)).local_anchor1((0.0, 4.0, 0.0)).motor_velocity(0, -2.0, 50.0).motor_max_force(0, 0.05).limits(0, -3.0, 0.8))
    w = make_world(s)
    w.step(400)
    pose, vel = w.body_states()
    assert abs(pose[a, 0] - 1.5) < 0.02 and abs(vel[a, 0]) < 0.02 and abs(pose[a, 1]) < 1e-3
    assert abs(pose[b, 0] - (-0.7)) < 0.05 and abs(vel[b, 0]) < 0.02     # acceleration-based spring: offset g / k = 3 / 400
    assert abs(pose[c, 0] - 0.8) < 0.02                                     # the weak motor loses against gravity: upper limit


def test_joint_limits_and_motors_oracle():
    mk = lambda s: oracle_lib.OracleWorld(s)
    angular_limits_are_reached(mk)
    prismatic_limits_and_position_motor(mk)


# ---- coupled joint axes: SpringJoint / RopeJoint / coupled angular limit (joint_constraint_helper.rs:210-283, :333-409, :725-798) ----
def spring_and_rope_joints(make_world):
    """A ball hanging from a SpringJoint rests where k * extension = m g (force-based motor on the distance between the anchors,
    spring_joint.rs:32-38); a ball on a RopeJoint released sideways never gets farther than max_dist from its anchor and ends
    up hanging at max_dist (rope_joint.rs); a rope shorter than needed leaves a resting ball alone (limit row inactive)."""
    import math
    from rapier_b200.sets import SpringJointBuilder, RopeJointBuilder
    s = scenes.Scene("spring_rope", gravity=(0.0, -9.81, 0.0))
    base = s.bodies.insert(RigidBodyBuilder.fixed())
    m = 4.0 / 3.0 * math.pi * 0.5 ** 3
    a = s.insert(RigidBodyBuilder.dynamic().translation((0.0, -2.0, 0.0)).can_sleep(False), ColliderBuilder.ball(0.5))
    s.joints.insert(base, a, SpringJointBuilder(1.5, 50.0, 5.0))
    b = s.insert(RigidBodyBuilder.dynamic().translation((11.0, 0.0, 0.0)).linear_damping(0.8).can_sleep(False), ColliderBuilder.ball(0.5))
    s.joints.insert(base, b, RopeJointBuilder(2.0).local_anchor1((10.0, 0.0, 0.0)))
    s.insert(RigidBodyBuilder.fixed().translation((20.0, -0.5, 0.0)), ColliderBuilder.cuboid(2.0, 0.5, 2.0))
    c = s.insert(RigidBodyBuilder.dynamic().translation((20.0, 0.5, 0.0)).can_sleep(False), ColliderBuilder.ball(0.5))
    s.joints.insert(base, c, RopeJointBuilder(5.0).local_anchor1((20.0, 3.0, 0.0)))
    w = make_world(s)
    far = 0.0
    for _ in range(600):
        w.step()
        pb = w.body_states()[0][b, :3]
        far = max(far, math.dist(pb, (10.0, 0.0, 0.0)))
    pose, vel = w.body_states()
    assert abs(pose[a, 1] - (-(1.5 + m * 9.81 / 50.0))) < 0.02 and abs(pose[a, 0]) < 1e-3 and abs(vel[a, 1]) < 0.02, pose[a]
    assert far < 2.06, far
    assert abs(math.dist(pose[b, :3], (10.0, 0.0, 0.0)) - 2.0) < 0.03 and abs(pose[b, 0] - 10.0) < 0.2, pose[b]
    assert abs(pose[c, 1] - 0.5) < 0.01 and abs(pose[c, 0] - 20.0) < 1e-3


def heavy_cubes_rest_on_spring_jointed_balls(make_world, num=30):
    """crates/rapier3d/tests/joint_contact_solve_order.rs: heavy cubes (200 x the ball mass) dropped onto light balls hanging
    from spring joints of increasing damping; joints are solved before contacts in every pass, so no cube tunnels through."""
    import math
    from rapier_b200.sets import SpringJointBuilder
    s = scenes.Scene("spring_balls", gravity=(0.0, -9.81, 0.0))
    ground = s.bodies.insert(RigidBodyBuilder.fixed())
    radius, stiffness = 0.5, 1.0e3
    mass = 4.0 / 3.0 * math.pi * radius ** 3
    critical = 2.0 * math.sqrt(stiffness * mass)
    pairs = []
    for i in range(num + 1):
        bp = (-6.0 + 1.5 * i, 4.5, 0.0)
        ball = s.insert(RigidBodyBuilder.dynamic().translation(bp).can_sleep(False), ColliderBuilder.ball(radius))
        damping = (i / (num / 2.0)) * critical
        s.joints.insert(ground, ball, SpringJointBuilder(0.0, stiffness, damping).local_anchor1((bp[0], bp[1] - 3.0, bp[2])))
        cube = s.insert(RigidBodyBuilder.dynamic().translation((bp[0], bp[1] + 5.0, bp[2])), ColliderBuilder.cuboid(radius, radius, radius).density(100.0))
        pairs.append((ball, cube))
    w = make_world(s)
    w.step(300)
    pose, _ = w.body_states()
    assert np.isfinite(pose).all()
    for i, (ball, cube) in enumerate(pairs):
        assert pose[cube, 1] > pose[ball, 1], (i, pose[cube, 1], pose[ball, 1])


def coupled_angular_spring_joint_stays_finite(make_world):
    """crates/rapier3d/tests/issue_792_coupled_angular_spring.rs: coupled ANG_X | ANG_Z axes with spring-like motors and
    limits [0, 0.5], first body kinematic; the coupled angular motor is a solver no-op, the coupled limit keeps the angle
    between the two frames' Y axes inside the cone."""
    import math
    from rapier_b200.sets import GenericJointBuilder
    s = scenes.Scene("issue_792", gravity=(0.0, -9.81, 0.0))
    kin = s.insert(RigidBodyBuilder.kinematic_position_based(), ColliderBuilder.ball(1.0))
    dyn = s.insert(RigidBodyBuilder.dynamic().translation((0.0, -5.0, 0.0)).linvel((0.1, 0.0, 0.1)).can_sleep(False), ColliderBuilder.ball(1.0))
    j = (GenericJointBuilder(0b000111).local_anchor2((0.0, 5.0, 0.0)).motor(3, 0.0, 0.0, 0.0, 0.5).motor(5, 0.0, 0.0, 0.0, 0.5)
         .limits(3, 0.0, 0.5).limits(5, 0.0, 0.5).coupled_axes(0b101000))
    s.joints.insert(kin, dyn, j)
    w = make_world(s)
    m = 4.0 / 3.0 * math.pi
    worst = 0.0
    for i in range(200):
        if i == 50:   # apply_impulse((5, 0, 0))
            pose, vel = w.body_states()
            v = vel[dyn].copy()
            v[0] += 5.0 / m
            w.set_body_states([dyn], vel6=[v])
        w.step()
        pose, _ = w.body_states()
        q = pose[dyn, 3:7]
        ycol_y = 1.0 - 2.0 * (float(q[0]) ** 2 + float(q[2]) ** 2)          # (R e_y) . e_y
        worst = max(worst, math.acos(max(-1.0, min(1.0, ycol_y))))
    pose, vel = w.body_states()
    assert np.isfinite(pose).all() and np.isfinite(vel).all()
    assert abs(math.dist(pose[dyn, :3], (0.0, 0.0, 0.0)) - 5.0) < 0.05      # the spherical part holds
    assert worst < 0.5 + 0.08, worst                                         # the cone limit holds (soft: a few degrees of overshoot)
    # the cone itself: two bodies pinned at their centres, no gravity; one tumbles about X and is stopped when its Y axis has
    # tilted by the limit, the other spins about Y (the axis that is not coupled) and is not limited at all
    s = scenes.Scene("cone_limit", gravity=(0.0, 0.0, 0.0))
    base = s.bodies.insert(RigidBodyBuilder.fixed())
    tilt = s.insert(RigidBodyBuilder.dynamic().angvel((1.5, 0.0, 0.0)).can_sleep(False), ColliderBuilder.cuboid(0.5, 0.5, 0.5))
    spin = s.insert(RigidBodyBuilder.dynamic().translation((5.0, 0.0, 0.0)).angvel((0.0, 2.0, 0.0)).can_sleep(False), ColliderBuilder.cuboid(0.5, 0.5, 0.5))
    s.joints.insert(base, tilt, GenericJointBuilder(0b000111).limits(3, 0.0, 0.5).limits(5, 0.0, 0.5).coupled_axes(0b101000))
    s.joints.insert(base, spin, GenericJointBuilder(0b000111).local_anchor1((5.0, 0.0, 0.0)).limits(3, 0.0, 0.5).limits(5, 0.0, 0.5).coupled_axes(0b101000))
    w = make_world(s)
    worst = 0.0
    for i in range(240):
        w.step()
        q = w.body_states()[0][tilt, 3:7]
        worst = max(worst, math.acos(max(-1.0, min(1.0, 1.0 - 2.0 * (float(q[0]) ** 2 + float(q[2]) ** 2)))))
    pose, vel = w.body_states()
    assert 0.45 < worst < 0.56, worst
    assert abs(vel[spin, 4] - 2.0) < 1e-3 and abs(vel[tilt, 3]) < 1.5        # the spin is untouched, the tumble was stopped / reversed


def test_coupled_joint_axes_oracle():
    mk = lambda s: oracle_lib.OracleWorld(s)
    spring_and_rope_joints(mk)
    heavy_cubes_rest_on_spring_jointed_balls(mk)
    coupled_angular_spring_joint_stays_finite(mk)


# ---- substep solve-groups: RigidBody::additional_solver_iterations (island_manager/substep_groups.rs; crates/rapier3d/tests/additional_solver_iterations.rs) ----
def _heavy_stack(extra, drop=False):
    s = scenes.Scene("heavy_stack", gravity=(0.0, -9.81, 0.0))
    s.insert(RigidBodyBuilder.fixed(), ColliderBuilder.cuboid(10.0, 0.5, 10.0).translation((0.0, -0.5, 0.0)))
    light = s.insert(RigidBodyBuilder.dynamic().translation((0.0, 0.5, 0.0)), ColliderBuilder.cuboid(0.5, 0.5, 0.5).density(1.0))
    heavy = s.insert(RigidBodyBuilder.dynamic().translation((0.1, 3.0, 0.0) if drop else (0.0, 1.5, 0.0)).additional_solver_iterations(extra),
                     ColliderBuilder.cuboid(0.5, 0.5, 0.5).density(200.0))
    return s, light, heavy


def additional_solver_iterations(make_world):
    """additional_solver_iterations.rs: a heavy cube (200 x) on a light one with 16 extra substeps on the heavy body stays put;
    a rope of spherical joints with a 100 x heavier end and 16 extra substeps keeps its length; the extra substeps change the
    trajectory of a stack that is still settling and are deterministic; a second, unrelated stack in the same world keeps the
    base cadence: it moves exactly as it does alone (the groups are per connected component)."""
    s, light, heavy = _heavy_stack(16)
    w = make_world(s)
    w.step(300)
    pose, vel = w.body_states()
    assert 0.3 < pose[light, 1] < 0.7 and 1.2 < pose[heavy, 1] < 1.8, (pose[light], pose[heavy])
    assert np.linalg.norm(vel[heavy, :3]) < 0.1

    from rapier_b200.sets import SphericalJointBuilder
    s = scenes.Scene("heavy_chain", gravity=(0.0, -9.81, 0.0))
    prev = s.bodies.insert(RigidBodyBuilder.fixed())
    for i in range(6):
        last = i == 5
        b = RigidBodyBuilder.dynamic().translation((0.0, -(i + 1.0), 0.0))
        if last:
            b = b.additional_solver_iterations(16)
        link = s.insert(b, ColliderBuilder.ball(0.4).density(100.0 if last else 1.0))
        s.joints.insert(prev, link, SphericalJointBuilder().local_anchor1((0.0, -0.5, 0.0)).local_anchor2((0.0, 0.5, 0.0)))
        prev = link
    w = make_world(s)
    w.step(300)
    end = w.body_states()[0][prev, :3]
    assert np.isfinite(end).all() and np.linalg.norm(end) < 20.0 and -7.5 < end[1] < -4.5, end
    assert -6.1 < end[1] < -5.9, end   # (with 20 substeps the joints barely stretch)

    def run(extra):
        s, light, heavy = _heavy_stack(extra, drop=True)
        w = make_world(s)
        w.step(60)
        pose, _ = w.body_states()
        return pose[[light, heavy]].copy()
    plain, extra1, extra2 = run(0), run(8), run(8)
    assert (extra1.view(np.uint32) == extra2.view(np.uint32)).all()
    assert not (plain.view(np.uint32) == extra1.view(np.uint32)).all()

    # per-component cadence: a base-cadence stack next to an elevated one moves as it does on its own, bit for bit
    def twin(with_elevated):
        s = scenes.Scene("twin", gravity=(0.0, -9.81, 0.0))
        s.insert(RigidBodyBuilder.fixed(), ColliderBuilder.cuboid(30.0, 0.5, 30.0).translation((0.0, -0.5, 0.0)))
        a = s.insert(RigidBodyBuilder.dynamic().translation((0.0, 0.6, 0.0)).rotation((0.0, 0.3, 0.1)), ColliderBuilder.cuboid(0.5, 0.5, 0.5))
        b = s.insert(RigidBodyBuilder.dynamic().translation((0.2, 1.9, 0.1)), ColliderBuilder.cuboid(0.4, 0.4, 0.4).density(3.0))
        if with_elevated:
            s.insert(RigidBodyBuilder.dynamic().translation((8.0, 0.5, 0.0)), ColliderBuilder.cuboid(0.5, 0.5, 0.5).density(1.0))
            s.insert(RigidBodyBuilder.dynamic().translation((8.1, 2.5, 0.0)).additional_solver_iterations(5), ColliderBuilder.cuboid(0.5, 0.5, 0.5).density(200.0))
        w = make_world(s)
        w.step(90)
        return w.body_states()[0][[a, b]].copy()
    alone, beside = twin(False), twin(True)
    assert (alone.view(np.uint32) == beside.view(np.uint32)).all()


def test_additional_solver_iterations_oracle():
    additional_solver_iterations(lambda s: oracle_lib.OracleWorld(s))


# ---- sensors (Collider::is_sensor; narrow_phase/intersections.rs; CollisionEventFlags::SENSOR) ---------------------------------
def sensors(make_world):
    """A ball falls through a fixed sensor zone: Started then Stopped with the SENSOR flag, and its fall is the free fall of the
    same scene without the zone, bit for bit; a massless sensor aura on a dynamic body reports the ground before the body
    touches it and does not hold the body up; a fast body is not stopped by a thin sensor wall (a solid twin of the wall stops it)."""
    def scene(with_zone):
        s = scenes.Scene("sensor_zone", gravity=(0.0, -9.81, 0.0))
        s.insert(RigidBodyBuilder.fixed().translation((0.0, -0.5, 0.0)), ColliderBuilder.cuboid(10.0, 0.5, 10.0))
        ball = s.insert(RigidBodyBuilder.dynamic().translation((0.0, 6.0, 0.0)).can_sleep(False), ColliderBuilder.ball(0.3).active_events(1))
        zone = s.colliders.insert(ColliderBuilder.cuboid(2.0, 0.5, 2.0).translation((0.0, 3.0, 0.0)).sensor(True)) if with_zone else -1
        return s, ball, zone
    s, ball, zone = scene(True)
    w = make_world(s)
    s0, ball0, _ = scene(False)
    w0 = make_world(s0)
    events = []
    for i in range(70):
        w.step(); w0.step()
        events += w.collision_events(with_flags=True)
        if i < 60:
            assert (w.body_states()[0][ball].view(np.uint32) == w0.body_states()[0][ball0].view(np.uint32)).all(), i
    zone_events = [e for e in events if zone in (e[0], e[1])]
    assert [e[2] for e in zone_events] == [1, 0] and all(e[4] == 1 for e in zone_events), zone_events
    assert zone_events[0][3] < zone_events[1][3]
    # the narrow phase of step k sees the pose after k - 1 steps of 4 substeps of semi-implicit Euler; the ball (radius 0.3) and
    # the zone (y in [2.5, 3.5]) intersect while 2.2 <= y <= 3.8
    g, h = 9.81, 1.0 / 240.0
    y_after = lambda n: 6.0 - g * h * h * (4 * n) * (4 * n + 1) / 2.0
    k_in, k_out = zone_events[0][3], zone_events[1][3]
    assert y_after(k_in - 1) <= 3.8 < y_after(k_in - 2), (k_in, y_after(k_in - 1))
    assert y_after(k_out - 1) < 2.2 <= y_after(k_out - 2), (k_out, y_after(k_out - 1))
    ground_events = [e for e in events if zone not in (e[0], e[1])]
    assert ground_events and all(e[4] == 0 for e in ground_events)

    s = scenes.Scene("sensor_aura", gravity=(0.0, -9.81, 0.0))
    s.insert(RigidBodyBuilder.fixed().translation((0.0, -0.5, 0.0)), ColliderBuilder.cuboid(10.0, 0.5, 10.0))
    body = s.insert(RigidBodyBuilder.dynamic().translation((0.0, 3.0, 0.0)).can_sleep(False), ColliderBuilder.cuboid(0.3, 0.3, 0.3))
    aura = s.colliders.insert_with_parent(ColliderBuilder.ball(1.0).density(0.0).sensor(True).active_events(1), body)
    w = make_world(s)
    first = None
    for i in range(120):
        w.step()
        for e in w.collision_events(with_flags=True):
            if aura in (e[0], e[1]) and e[2] == 1 and first is None:
                first = (i, float(w.body_states()[0][body, 1]))
                assert e[4] == 1
    pose, vel = w.body_states()
    assert first is not None and 0.75 < first[1] < 1.0, first         # reported in the step whose narrow phase saw the aura (radius 1) reach the ground
    assert abs(pose[body, 1] - 0.3) < 0.01 and abs(vel[body, 1]) < 0.02   # the box rests on the ground: the aura carries nothing

    def wall(sensor):
        s = scenes.Scene("sensor_wall", gravity=(0.0, 0.0, 0.0))
        c = ColliderBuilder.cuboid(0.02, 2.0, 2.0).translation((3.0, 0.0, 0.0))
        s.colliders.insert(c.sensor(True) if sensor else c)
        b = s.insert(RigidBodyBuilder.dynamic().linvel((300.0, 0.0, 0.0)).can_sleep(False), ColliderBuilder.ball(0.2))
        w = make_world(s)
        w.step(3)
        return float(w.body_states()[0][b, 0])
    assert wall(True) > 10.0 and wall(False) < 3.0


def test_sensors_oracle():
    sensors(lambda s: oracle_lib.OracleWorld(s))


# ---- compound bodies: general composite mass properties, wide bodies asleep (crates/rapier3d/tests/sleep_wide_bodies.rs) ----------
def _wide_compound_scene(n=8):
    """sleep_wide_bodies.rs:10-48: U-shaped compounds (a wide bar plus two uprights, max_extent ~ 4) at varied yaws."""
    s = scenes.Scene("wide_compounds", gravity=(0.0, -9.81, 0.0))
    s.insert(RigidBodyBuilder.fixed().translation((0.0, -0.1, 0.0)), ColliderBuilder.cuboid(50.0, 0.1, 50.0))
    rad = 0.2
    handles = []
    for i in range(n):
        for k in range(n):
            h = s.insert(RigidBodyBuilder.dynamic().translation((i * 6.0, rad + 0.01, k * 6.0)).rotation((0.0, i * 0.11 + k * 0.037, 0.0)),
                         ColliderBuilder.cuboid(rad * 10.0, rad, rad))
            s.colliders.insert_with_parent(ColliderBuilder.cuboid(rad, rad * 10.0, rad).translation((rad * 10.0, rad * 10.0, 0.0)), h)
            s.colliders.insert_with_parent(ColliderBuilder.cuboid(rad, rad * 10.0, rad).translation((-rad * 10.0, rad * 10.0, 0.0)), h)
            handles.append(h)
    return s, handles


def compound_bodies(make_world, n=8):
    """sleep_wide_bodies.rs: wide compounds at rest fall asleep (all of them), a still far-reaching body does not move at all and
    sleeps.  The composite mass properties behind them: a dumbbell of two boxes rotated by 45 degrees about Z spins about its long
    axis at constant rate (a principal axis found by the eigen-decomposition) and its angular momentum is conserved when it tumbles;
    an L-shaped body (products of inertia non-zero in the body frame) dropped flat comes to rest flat."""
    import math
    s, handles = _wide_compound_scene(n)
    w = make_world(s)
    w.step(300)
    asleep = w.sleeping()
    assert all(asleep[h] for h in handles), int(sum(1 for h in handles if not asleep[h]))

    s = scenes.Scene("still_wide", gravity=(0.0, 0.0, 0.0))
    h = s.insert(RigidBodyBuilder.dynamic().rotation((0.3, -0.7, 0.15)), ColliderBuilder.cuboid(0.2, 8.0, 0.2))
    w = make_world(s)
    p0 = w.body_states()[0][h].copy()
    w.step(200)
    assert (w.body_states()[0][h].view(np.uint32) == p0.view(np.uint32)).all() and w.sleeping()[h]

    # dumbbell along the (1, 1, 0) diagonal: two cubes offset by +-(1, 1, 0), each rotated 45 degrees about Z
    q45 = (0.0, 0.0, math.pi / 4.0)
    def dumbbell(angvel):
        s = scenes.Scene("dumbbell", gravity=(0.0, 0.0, 0.0))
        b = s.insert(RigidBodyBuilder.dynamic().angvel(angvel).can_sleep(False), ColliderBuilder.cuboid(0.5, 0.2, 0.2).translation((1.0, 1.0, 0.0)).rotation(q45))
        s.colliders.insert_with_parent(ColliderBuilder.cuboid(0.5, 0.2, 0.2).translation((-1.0, -1.0, 0.0)).rotation(q45), b)
        return s, b
    s, b = dumbbell((2.0 / math.sqrt(2.0), 2.0 / math.sqrt(2.0), 0.0))   # spin about the long (principal) axis
    w = make_world(s)
    w.step(120)
    pose, vel = w.body_states()
    assert np.allclose(vel[b, 3:], (math.sqrt(2.0), math.sqrt(2.0), 0.0), atol=2e-3), vel[b]
    assert np.allclose(pose[b, :3], 0.0, atol=1e-4)
    # tumbling about a non-principal axis: |L| = |I w| is conserved by the gyroscopic term; analytic principal moments
    m = 0.5 * 0.2 * 0.2 * 8.0
    i_long = 2.0 * (m * (0.2 ** 2 + 0.2 ** 2) / 3.0)
    i_cross = 2.0 * (m * (0.5 ** 2 + 0.2 ** 2) / 3.0 + m * 2.0)
    s, b = dumbbell((1.0, 0.0, 0.5))
    w = make_world(s)
    def ang_momentum():
        pose, vel = w.body_states()
        q = pose[b, 3:7].astype(np.float64)
        x, y, z, ww = q
        R = np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - z * ww), 2 * (x * z + y * ww)],
                      [2 * (x * y + z * ww), 1 - 2 * (x * x + z * z), 2 * (y * z - x * ww)],
                      [2 * (x * z - y * ww), 2 * (y * z + x * ww), 1 - 2 * (x * x + y * y)]])
        e = R @ (np.array([1.0, 1.0, 0.0]) / math.sqrt(2.0))     # the long axis in world space
        om = vel[b, 3:].astype(np.float64)
        return i_long * e * (e @ om) + i_cross * (om - e * (e @ om))
    l0 = ang_momentum()
    w.step(240)
    l1 = ang_momentum()
    assert np.linalg.norm(l1 - l0) < 0.01 * np.linalg.norm(l0), (l0, l1)

    s = scenes.Scene("l_shape", gravity=(0.0, -9.81, 0.0))
    s.insert(RigidBodyBuilder.fixed().translation((0.0, -0.5, 0.0)), ColliderBuilder.cuboid(10.0, 0.5, 10.0))
    b = s.insert(RigidBodyBuilder.dynamic().translation((0.0, 0.5, 0.0)).can_sleep(False), ColliderBuilder.cuboid(1.0, 0.2, 0.2))
    s.colliders.insert_with_parent(ColliderBuilder.cuboid(0.2, 0.2, 0.8).translation((0.8, 0.0, 1.0)).density(3.0), b)
    w = make_world(s)
    w.step(200)
    pose, vel = w.body_states()
    assert abs(pose[b, 1] - 0.2) < 0.01 and np.abs(vel[b]).max() < 0.02, (pose[b], vel[b])
    q = pose[b, 3:7]
    assert abs(float(q[0])) < 0.01 and abs(float(q[2])) < 0.01     # still flat


def test_compound_bodies_oracle():
    compound_bodies(lambda s: oracle_lib.OracleWorld(s))


# ---- island structure (crates/rapier3d/tests/persistent_islands.rs; here islands are relabelled, not maintained: DESIGN dev. 5) ----
def island_structure(make_world, island_of):
    """persistent_islands.rs: touching boxes share an island and distant ones do not (merge_on_touch...); a teleported top box is
    split off in the very step its contact stops (separation_splits_immediately) and stays apart (split_on_separation); removing
    the middle box of a touching row splits the sides (body_removal_splits_row); a detached two-box chunk leaves together."""
    def ground():
        s = scenes.Scene("islands", gravity=(0.0, -9.81, 0.0))
        s.insert(RigidBodyBuilder.fixed().translation((0.0, -0.5, 0.0)), ColliderBuilder.cuboid(100.0, 0.5, 100.0))
        return s
    box = lambda s, x, y: s.insert(RigidBodyBuilder.dynamic().translation((x, y, 0.0)), ColliderBuilder.cuboid(0.5, 0.5, 0.5))
    same = lambda w, a, b: island_of(w)[a] >= 0 and island_of(w)[a] == island_of(w)[b]
    s = ground()
    bottom, top, lone = box(s, 0.0, 0.5), box(s, 0.0, 1.5), box(s, 20.0, 0.5)
    w = make_world(s)
    w.step(240)
    assert same(w, bottom, top) and not same(w, bottom, lone)
    w.set_body_states([top], pose7=[(40.0, 0.5, 0.0, 0.0, 0.0, 0.0, 1.0)])   # set_translation(.., wake_up = true)
    w.wake_up([top])
    w.step(1)
    assert not same(w, bottom, top), "the split must happen in the very step the contact stops"
    w.step(240)
    assert not same(w, bottom, top)

    s = ground()
    left, middle, right = box(s, 0.0, 0.5), box(s, 1.0, 0.5), box(s, 2.0, 0.5)
    w = make_world(s)
    w.step(240)
    assert same(w, left, right)
    w.remove(middle) if hasattr(w, "remove") else w.remove_bodies([middle])
    w.step(240)
    assert not same(w, left, right), "removing the bridging body must split the island"

    s = ground()
    a, b = box(s, 0.0, 0.5), box(s, 0.0, 1.5)       # a stack of four; the upper two are teleported away together
    c, d = box(s, 0.0, 2.5), box(s, 0.0, 3.5)
    w = make_world(s)
    w.step(240)
    assert same(w, a, d)
    pose = w.body_states()[0]
    moved = []
    for h in (c, d):
        p = pose[h].copy(); p[0] += 30.0; p[1] -= 2.0
        moved.append(p)
    w.set_body_states([c, d], pose7=moved)
    w.wake_up([c, d])
    w.step(2)
    assert same(w, c, d) and same(w, a, b) and not same(w, a, c)


def test_island_structure_oracle():
    island_structure(lambda s: oracle_lib.OracleWorld(s), lambda w: w.debug_read("island_of", np.int32))


# ---- the sleep metric gates on displacement (rigid_body_components.rs:1563-1590 test_sleep_gates_position_corrections) -----------
def sleep_gates_on_displacement(make_world):
    """update_energy tolerates a per-step displacement of 2 * threshold * dt (threshold 0.05 m/s, halved inside the test): a free
    body creeping at 1.5 x that budget never becomes eligible, one at 0.5 x the budget or at rest falls asleep."""
    s = scenes.Scene("creep", gravity=(0.0, 0.0, 0.0))
    budget_speed = 2.0 * 0.05
    fast = s.insert(RigidBodyBuilder.dynamic().linvel((budget_speed * 1.5, 0.0, 0.0)), ColliderBuilder.ball(0.5))
    slow = s.insert(RigidBodyBuilder.dynamic().translation((0.0, 5.0, 0.0)).linvel((budget_speed * 0.5, 0.0, 0.0)), ColliderBuilder.ball(0.5))
    still = s.insert(RigidBodyBuilder.dynamic().translation((0.0, 10.0, 0.0)), ColliderBuilder.ball(0.5))
    w = make_world(s)
    w.step(300)   # 10 x time_until_sleep
    asleep = w.sleeping()
    assert not asleep[fast] and asleep[slow] and asleep[still], asleep


def test_sleep_gates_on_displacement_oracle():
    sleep_gates_on_displacement(lambda s: oracle_lib.OracleWorld(s))


def multi_collider_slab_sleeps(make_world, ni=50, nj=40):
    """crates/rapier3d/tests/issue_970_multi_collider_body_perf.rs (its behavioural half): one dynamic body made of ni x nj unit boxes
    resting just above the ground settles on it (every one of its manifolds against the ground shares the two bodies: 128 colours, the
    rest in the overflow colour) and falls asleep within 400 steps."""
    s = scenes.Scene("slab", gravity=(0.0, -9.81, 0.0))
    s.colliders.insert(ColliderBuilder.cuboid(100.0, 0.5, 100.0))
    b = s.bodies.insert(RigidBodyBuilder.dynamic().translation((0.0, 1.05, 0.0)))
    for i in range(ni):
        for j in range(nj):
            s.colliders.insert_with_parent(ColliderBuilder.cuboid(0.5, 0.5, 0.5).translation((i - ni / 2.0, 0.0, j - nj / 2.0)), b)
    w = make_world(s)
    w.step(400)
    pose, _ = w.body_states()
    assert w.sleeping()[b] and abs(pose[b, 1] - 1.0) < 0.01, pose[b]


def test_multi_collider_slab_oracle():
    multi_collider_slab_sleeps(lambda s: oracle_lib.OracleWorld(s, threads=8))


# ---- capsules (parry Capsule; ColliderBuilder::capsule_{x,y,z}) ------------------------------------------------------------------
def capsules_rest(make_world):
    """Capsules of unit mass on a slab: lying (two-point manifold, rest height = radius), standing (rest height = half height +
    radius), a capsule lying along another one (two contacts on the shared interval), a ball on a capsule; the total contact
    impulse of each resting body on its support is m g dt within 1 % (total_contact_impulse.rs:58-75 with capsules)."""
    s = scenes.Scene("capsules", gravity=(0.0, -9.81, 0.0))
    s.insert(RigidBodyBuilder.fixed().translation((0.0, -0.5, 0.0)), ColliderBuilder.cuboid(20.0, 0.5, 20.0))
    lying = s.insert(RigidBodyBuilder.dynamic().translation((0.0, 0.3, 0.0)), ColliderBuilder.capsule_x(0.5, 0.3).mass(1.0))
    standing = s.insert(RigidBodyBuilder.dynamic().translation((4.0, 0.8, 0.0)), ColliderBuilder.capsule_y(0.5, 0.3).mass(1.0))
    lower = s.insert(RigidBodyBuilder.dynamic().translation((8.0, 0.5, 0.0)), ColliderBuilder.cuboid(1.0, 0.5, 0.5).mass(5.0))
    upper = s.insert(RigidBodyBuilder.dynamic().translation((8.0, 1.3, 0.0)), ColliderBuilder.capsule_x(0.6, 0.3).mass(1.0))
    w = make_world(s)
    w.step(300)
    pose, vel = w.body_states()
    assert abs(pose[lying, 1] - 0.3) < 0.01 and abs(pose[standing, 1] - 0.8) < 0.01 and abs(pose[upper, 1] - 1.3) < 0.015
    assert np.abs(vel).max() < 0.05
    cp = w.contact_pairs()
    by_pair = {tuple(k): (n, imp.sum()) for k, n, imp in zip(cp["colliders"].tolist(), cp["num_contacts"].tolist(), cp["impulses"])}
    mg_dt = 9.81 / 60.0
    assert by_pair[(0, lying + 0)][0] == 2 and abs(by_pair[(0, 1)][1] - mg_dt) / mg_dt < 0.01          # lying: two points
    assert abs(by_pair[(0, 2)][1] - mg_dt) / mg_dt < 0.01                                               # standing
    assert by_pair[(3, 4)][0] == 2 and abs(by_pair[(3, 4)][1] - mg_dt) / mg_dt < 0.01                   # capsule along a box: two points
    assert abs(by_pair[(0, 3)][1] - 6.0 * mg_dt) / (6.0 * mg_dt) < 0.01                                 # the box carries both


def test_capsules_oracle():
    capsules_rest(lambda s: oracle_lib.OracleWorld(s))


def prismatic_joint_stays_bounded_for_all_axis_rotations(make_world):
    """issue_746_prismatic_axis_frames.rs: two free-floating boxes on a prismatic joint whose second frame is given through
    local_axis2 for eight quarter-turn rotations of the second body: nothing may drift further than 5 units in 60 steps."""
    import math
    from rapier_b200.sets import PrismaticJointBuilder
    for i in range(8):
        angle = math.pi / 2.0 * i
        p = A.RbIntegrationParameters.default()
        p.dt = 0.016
        s = scenes.Scene("issue_746")
        b1 = s.insert(RigidBodyBuilder.dynamic().gravity_scale(0.0), ColliderBuilder.cuboid(1.0, 1.0, 1.0))
        b2 = s.insert(RigidBodyBuilder.dynamic().translation((1.0, 0.0, 0.0)).rotation((0.0, angle, 0.0)).gravity_scale(0.0), ColliderBuilder.cuboid(1.0, 1.0, 1.0))
        local_axis2 = (math.cos(angle), 0.0, math.sin(angle))   # Ry(angle)^-1 * X
        s.joints.insert(b1, b2, PrismaticJointBuilder((1.0, 0.0, 0.0)).local_axis1((1.0, 0.0, 0.0)).local_axis2(local_axis2).contacts_enabled(False))
        w = make_world(s, p)
        w.step(60)
        pose, _ = w.body_states()
        assert np.isfinite(pose).all() and np.linalg.norm(pose[:, :3], axis=1).max() < 5.0, (i, pose[:, :3])


def motor_position_with_rotating_base_stays_finite(make_world):
    """issue_856_motor_position_rotating_base.rs: a stiff position motor (target pi) on a revolute joint whose base is re-oriented
    by the user every step: every body stays finite and nothing is quarantined for 300 steps.  (The base carries a cuboid here:
    cylinders are not supported; its shape plays no role.)"""
    import math
    from rapier_b200.sets import RevoluteJointBuilder
    s = scenes.Scene("issue_856")
    base = s.insert(RigidBodyBuilder.dynamic().translation((0.0, 3.0, 0.0)), ColliderBuilder.cuboid(1.0, 0.2, 1.0))
    hammer = s.insert(RigidBodyBuilder.dynamic().translation((2.0, 3.0, 0.0)), ColliderBuilder.cuboid(0.5, 0.1, 0.1))
    s.joints.insert(base, hammer, RevoluteJointBuilder((0.0, 0.0, 1.0)).local_anchor1((1.0, 0.0, 0.0)).local_anchor2((-1.0, 0.0, 0.0)).motor_position(3, math.pi, 1.0e4, 100.0))
    w = make_world(s, None)
    for i in range(300):
        angle = i * 0.05
        pose, _ = w.body_states()
        w.set_body_states([base], pose7=[(pose[base, 0], pose[base, 1], pose[base, 2], 0.0, math.sin(angle / 2.0), 0.0, math.cos(angle / 2.0))])
        w.step()
        pose, _ = w.body_states()
        assert np.isfinite(pose).all(), i
        assert len(w.quarantine()) == 0, i


def test_more_joint_known_answers_oracle():
    mk = lambda s, p: oracle_lib.OracleWorld(s, params=p)
    prismatic_joint_stays_bounded_for_all_axis_rotations(mk)
    motor_position_with_rotating_base_stays_finite(mk)


# ---- crates/rapier3d/tests/sleep_wake.rs ----------------------------------------------------------------------------------------------
def _sleep_world():
    s = scenes.Scene("sleep_wake", gravity=(0.0, -9.81, 0.0))
    s.insert(RigidBodyBuilder.fixed().translation((0.0, -0.5, 0.0)), ColliderBuilder.cuboid(50.0, 0.5, 50.0))
    return s


def _cube(s, builder):
    return s.insert(builder, ColliderBuilder.cuboid(0.5, 0.5, 0.5))


def sleep_wake_scenarios(make_world):
    """sleep_wake.rs: woken_body_is_supported_by_recycled_contacts (:3-90), non_sleeping_neighbor_keeps_touching_row_awake
    (:150-181), impact_wakes_sleeping_region (:183-218), joint_keeps_both_sides_awake (:221-253), corner_velocity_sleep_metric
    (:255-287).  (sliding_support_wakes_sleeping_rider needs partial-island sleep, which is not modelled: DESIGN.md deviation 5.)"""
    # a sleeping cube woken by an impulse keeps being supported by its (recycled) contacts
    s = _sleep_world()
    cube = _cube(s, RigidBodyBuilder.dynamic().translation((0.0, 0.6, 0.0)))
    w = make_world(s)
    slept = False
    for _ in range(400):
        w.step()
        if w.sleeping()[cube]:
            slept = True
            break
    assert slept
    w.set_body_states([cube], vel6=[(0.5, 0.0, 0.0, 0.0, 0.0, 0.0)])   # apply_impulse(0.5 N s) on the 1 kg cube
    assert not w.sleeping()[cube]
    for _ in range(120):
        w.step()
        assert w.body_states()[0][cube, 1] > 0.4
    # a row of touching cubes with one that cannot sleep stays awake; a lone cube sleeps
    s = _sleep_world()
    row = [_cube(s, RigidBodyBuilder.dynamic().translation((float(i), 0.5, 0.0)).can_sleep(i != 0)) for i in range(8)]
    lone = _cube(s, RigidBodyBuilder.dynamic().translation((30.0, 0.5, 0.0)))
    w = make_world(s)
    w.step(400)
    sl, (pose, _) = w.sleeping(), w.body_states()
    assert not any(sl[h] for h in row) and sl[lone]
    assert all(abs(pose[h, 1] - 0.5) < 0.1 and abs(pose[h, 0] - i) < 0.1 for i, h in enumerate(row))
    # an impact wakes a sleeping row
    s = _sleep_world()
    row = [_cube(s, RigidBodyBuilder.dynamic().translation((float(i), 0.5, 0.0))) for i in range(6)]
    w = make_world(s)
    w.step(400)
    assert all(w.sleeping()[h] for h in row)
    w.insert(RigidBodyBuilder.dynamic().translation((-3.0, 0.5, 0.0)).linvel((20.0, 0.0, 0.0)), ColliderBuilder.cuboid(0.5, 0.5, 0.5))
    w.step(30)
    pose, vel = w.body_states()
    assert not w.sleeping()[row[0]] and (np.linalg.norm(vel[row[0], :3]) > 0.05 or pose[row[0], 0] > 0.05)
    # a joint keeps both sides awake next to a body that cannot sleep
    s = _sleep_world()
    mover = _cube(s, RigidBodyBuilder.dynamic().translation((0.0, 0.5, 0.0)).can_sleep(False))
    b = _cube(s, RigidBodyBuilder.dynamic().translation((1.0, 0.5, 0.0)))
    a = _cube(s, RigidBodyBuilder.dynamic().translation((2.5, 0.5, 0.0)))
    control = _cube(s, RigidBodyBuilder.dynamic().translation((10.0, 0.5, 0.0)))
    from rapier_b200.sets import FixedJointBuilder
    s.joints.insert(b, a, FixedJointBuilder().local_anchor1((1.5, 0.0, 0.0)).local_anchor2((0.0, 0.0, 0.0)))
    w = make_world(s)
    w.step(400)
    sl, (pose, _) = w.sleeping(), w.body_states()
    assert not sl[mover] and not sl[b] and not sl[a] and sl[control]
    assert abs(pose[a, 0] - 2.5) < 0.1 and abs(pose[b, 0] - 1.0) < 0.1
    # the sleep metric looks at the motion of the farthest point: a slowly pivoting long beam stays awake, a small spinner sleeps
    s = scenes.Scene("corner_velocity", gravity=(0.0, -9.81, 0.0))
    beam = s.insert(RigidBodyBuilder.dynamic().translation((0.0, 30.0, 0.0)).angvel((0.0, 0.0, 0.3)).gravity_scale(0.0), ColliderBuilder.cuboid(10.0, 0.1, 0.1))
    pebble = s.insert(RigidBodyBuilder.dynamic().translation((0.0, 30.0, 20.0)).angvel((0.0, 0.0, 0.55)).gravity_scale(0.0), ColliderBuilder.cuboid(0.05, 0.05, 0.05))
    w = make_world(s)
    w.step(200)
    assert not w.sleeping()[beam] and w.sleeping()[pebble]


def test_sleep_wake_oracle():
    from incremental_cases import OracleSets
    sleep_wake_scenarios(lambda s: OracleSets(s))


def dominance_groups(make_world):
    """RigidBodyDominance (rigid_body_components.rs:1255-1276; contact_with_twist_friction.rs:71-84): in a contact between
    two groups the higher one is world-attached.  No test of the reference pins this; the scenarios follow its documented
    behaviour: (1) a light dominating box carries a dominated box 320 times its mass without sinking, while the same pair
    with equal groups sinks measurably deeper into the ground contact; (2) a dominating dynamic box is not pushed by a
    kinematic wall sweeping through it, an ordinary one is; (3) a dominated ball bounces off a dominating free-floating
    box without moving it."""
    def crush(groups):
        s = scenes.Scene("crush", gravity=(0.0, -9.81, 0.0))
        s.insert(RigidBodyBuilder.fixed().translation((0.0, -0.5, 0.0)), ColliderBuilder.cuboid(10.0, 0.5, 10.0))
        light = s.insert(RigidBodyBuilder.dynamic().translation((0.0, 0.25, 0.0)).dominance_group(groups[0]), ColliderBuilder.cuboid(0.25, 0.25, 0.25).density(0.5))
        heavy = s.insert(RigidBodyBuilder.dynamic().translation((0.0, 1.0, 0.0)).dominance_group(groups[1]), ColliderBuilder.cuboid(0.5, 0.5, 0.5).density(20.0))
        w = make_world(s)
        w.step(240)
        pose, vel = w.body_states()
        assert np.abs(vel[[light, heavy]]).max() < 0.05
        return float(pose[light, 1]), float(pose[heavy, 1])
    y_dom, top_dom = crush((4, 0))
    y_eq, _ = crush((0, 0))
    assert abs(y_dom - 0.25) < 0.003 and abs(top_dom - 1.0) < 0.01, (y_dom, top_dom)   # as if nothing lay on it
    assert y_eq < y_dom - 1e-4, (y_eq, y_dom)                                             # the load shows without dominance
    s = scenes.Scene("wall", gravity=(0.0, 0.0, 0.0))
    s.insert(RigidBodyBuilder.kinematic_velocity_based().translation((-2.0, 0.0, 0.0)).linvel((2.0, 0.0, 0.0)), ColliderBuilder.cuboid(0.2, 1.0, 1.0))
    proud = s.insert(RigidBodyBuilder.dynamic().translation((0.0, 0.0, 0.0)).dominance_group(1), ColliderBuilder.cuboid(0.3, 0.3, 0.3))
    s.insert(RigidBodyBuilder.kinematic_velocity_based().translation((-2.0, 0.0, 10.0)).linvel((2.0, 0.0, 0.0)), ColliderBuilder.cuboid(0.2, 1.0, 1.0))
    meek = s.insert(RigidBodyBuilder.dynamic().translation((0.0, 0.0, 10.0)), ColliderBuilder.cuboid(0.3, 0.3, 0.3))
    w = make_world(s)
    w.step(90)
    pose, vel = w.body_states()
    assert abs(pose[proud, 0]) < 1e-6 and np.abs(vel[proud]).max() < 1e-6     # the wall passed through it
    assert pose[meek, 0] > 0.4 and vel[meek, 0] > 1.5                          # the ordinary box is shoved along
    s = scenes.Scene("bounce_off", gravity=(0.0, 0.0, 0.0))
    rock = s.insert(RigidBodyBuilder.dynamic().dominance_group(7), ColliderBuilder.cuboid(0.5, 0.5, 0.5).restitution(1.0))
    ball = s.insert(RigidBodyBuilder.dynamic().translation((-3.0, 0.0, 0.0)).linvel((5.0, 0.0, 0.0)).dominance_group(-7), ColliderBuilder.ball(0.3).restitution(1.0))
    w = make_world(s)
    w.step(90)
    pose, vel = w.body_states()
    assert np.abs(pose[rock, :3]).max() < 1e-6 and np.abs(vel[rock]).max() < 1e-6
    assert vel[ball, 0] < -4.0                                                  # full elastic return: the rock took no momentum


def test_dominance_groups_oracle():
    dominance_groups(lambda s: oracle_lib.OracleWorld(s))


def joint_warmstart(make_world):
    """IntegrationParameters::warmstart_joints (integration_parameters.rs:300; joint_constraint_builder.rs:116-150;
    worker.rs:548).  The carried impulse is applied row by row right before that row's exact solve, so for rigid
    (unbounded, cfm ~ 0) rows the body velocities after the solve are those of a cold solve: the option only shows through
    the cfm term of soft rows.  No reference test pins it; checked here: a hanging chain of rigid joints moves as without
    the option (to rounding), its top joint reports the same impulse (the chain's weight x substep length), a spring
    (position) motor DOES respond differently, and with a warm-start coefficient of 0 the option is a no-op bit for bit."""
    from rapier_b200.sets import RevoluteJointBuilder, SphericalJointBuilder
    def chain(warm, coeff=1.0, spring=False):
        s = scenes.Scene("chain", gravity=(0.0, -9.81, 0.0))
        prev = s.bodies.insert(RigidBodyBuilder.fixed().translation((0.0, 10.0, 0.0)))
        ids = []
        for i in range(12):
            b = s.insert(RigidBodyBuilder.dynamic().translation((0.0, 9.5 - 1.0 * i, 0.0)).can_sleep(False), ColliderBuilder.ball(0.2).density(30.0))
            s.joints.insert(prev, b, SphericalJointBuilder().local_anchor1((0.0, -0.5, 0.0) if i else (0.0, 0.0, 0.0)).local_anchor2((0.0, 0.5, 0.0)))
            ids.append(b)
            prev = b
        if spring:
            arm = s.insert(RigidBodyBuilder.dynamic().translation((5.5, 10.0, 0.0)).can_sleep(False), ColliderBuilder.cuboid(0.5, 0.1, 0.1))
            s.joints.insert(0, arm, RevoluteJointBuilder((0.0, 0.0, 1.0)).local_anchor1((5.0, 0.0, 0.0)).local_anchor2((-0.5, 0.0, 0.0)).motor_position(3, 0.5, 40.0, 2.0))
            ids.append(arm)
        p = A.RbIntegrationParameters.default()
        p.warmstart_joints = 1 if warm else 0
        p.warmstart_coefficient = coeff
        w = make_world(s, p)
        w.step(240)
        pose, vel = w.body_states()
        assert np.isfinite(pose).all()
        return pose[ids], w
    cold, wc = chain(False)
    warm, ww = chain(True)
    assert np.abs(cold - warm).max() < 1e-3 and abs(float(warm[-1, 1]) + 1.5) < 0.05
    mass = 12 * 30.0 * 4.0 / 3.0 * np.pi * 0.2 ** 3
    for w in (wc, ww):
        ji = w.debug_read("joint_impulses", np.float32).reshape(-1, 6)
        assert abs(np.abs(ji[0, :3]).max() - mass * 9.81 / 240.0) < 0.05 * mass * 9.81 / 240.0, (ji[0], mass * 9.81 / 240.0)
    cold_s, _ = chain(False, spring=True)
    warm_s, _ = chain(True, spring=True)
    assert (cold_s[-1].view(np.uint32) != warm_s[-1].view(np.uint32)).any() and np.abs(cold_s[-1] - warm_s[-1]).max() < 0.2
    off, _ = chain(True, coeff=0.0, spring=True)
    ref, _ = chain(False, coeff=0.0, spring=True)
    assert (off.view(np.uint32) == ref.view(np.uint32)).all()


def test_joint_warmstart_oracle():
    joint_warmstart(lambda s, p: oracle_lib.OracleWorld(s, params=p))


def convex_polyhedra(make_world):
    """ColliderBuilder::{convex_hull, round_convex_hull} (collider.rs:1039-1060) against everything else.  parry's manifolds
    are not in the tree, so the answers are physical: (1) a tetrahedron of mass 1 (volume 1/6, density 6) rests on its
    base and is carried by exactly its weight (total_contact_impulse.rs's criterion: sum of impulses = m g dt +- 1 %);
    (2) a cube given as a hull moves like the cuboid twin dropped beside it; (3) hull-on-hull, ball-on-hull, hull-on-
    cuboid and capsule-on-hull stacks rest at the analytic heights; a round hull rests higher by its border radius;
    (4) a spinning free tetrahedron keeps its angular momentum direction (principal axes and centre of mass are right:
    no wobble of the centre of mass); (5) a heap of random hulls settles on the ground without sinking in."""
    tet = [(0.0, 0.0, 0.0), (1.0, 0.0, 0.0), (0.0, 1.0, 0.0), (0.0, 0.0, 1.0)]
    cube = [(x, y, z) for x in (-0.5, 0.5) for y in (-0.5, 0.5) for z in (-0.5, 0.5)]
    s = scenes.Scene("convex", gravity=(0.0, -9.81, 0.0))
    s.insert(RigidBodyBuilder.fixed().translation((0.0, -0.5, 0.0)), ColliderBuilder.cuboid(30.0, 0.5, 30.0))
    t = s.insert(RigidBodyBuilder.dynamic().translation((0.0, 0.3, 0.0)).can_sleep(False), ColliderBuilder.convex_hull(tet).density(6.0))
    hc = s.insert(RigidBodyBuilder.dynamic().translation((4.0, 2.0, 0.0)).rotation((0.3, 0.2, 0.1)).can_sleep(False), ColliderBuilder.convex_hull(cube))
    cc = s.insert(RigidBodyBuilder.dynamic().translation((8.0, 2.0, 0.0)).rotation((0.3, 0.2, 0.1)).can_sleep(False), ColliderBuilder.cuboid(0.5, 0.5, 0.5))
    a0 = s.insert(RigidBodyBuilder.dynamic().translation((12.0, 0.55, 0.0)), ColliderBuilder.convex_hull(cube))
    a1 = s.insert(RigidBodyBuilder.dynamic().translation((12.05, 1.6, 0.05)), ColliderBuilder.convex_hull(cube))
    a2 = s.insert(RigidBodyBuilder.dynamic().translation((12.0, 2.5, 0.0)), ColliderBuilder.ball(0.4))
    b0 = s.insert(RigidBodyBuilder.dynamic().translation((16.0, 0.55, 0.0)), ColliderBuilder.cuboid(0.5, 0.5, 0.5))
    b1 = s.insert(RigidBodyBuilder.dynamic().translation((16.0, 1.6, 0.0)).rotation((0.0, 0.4, 0.0)), ColliderBuilder.convex_hull(cube))
    b2 = s.insert(RigidBodyBuilder.dynamic().translation((16.0, 2.4, 0.0)), ColliderBuilder.capsule_x(0.3, 0.2))
    rd = s.insert(RigidBodyBuilder.dynamic().translation((20.0, 0.7, 0.0)), ColliderBuilder.round_convex_hull(cube, 0.1))
    w = make_world(s)
    w.step(300)
    pose, vel = w.body_states()
    assert np.abs(vel).max() < 0.05
    assert abs(pose[t, 1]) < 0.01                                               # on its base (the body origin is the corner)
    cp = w.contact_pairs()
    tet_impulse = sum(float(imp.sum()) for k, imp in zip(cp["colliders"].tolist(), cp["impulses"]) if 1 in k)
    assert abs(tet_impulse - 9.81 / 60.0) < 0.01 * 9.81 / 60.0, tet_impulse
    assert abs(pose[hc, 1] - 0.5) < 0.01 and abs(pose[cc, 1] - 0.5) < 0.01
    assert np.abs((pose[hc, :3] - pose[cc, :3]) - (-4.0, 0.0, 0.0)).max() < 0.05     # same tumble as the cuboid twin
    assert abs(pose[a0, 1] - 0.5) < 0.01 and abs(pose[a1, 1] - 1.5) < 0.015 and abs(pose[a2, 1] - 2.4) < 0.02
    assert abs(pose[b0, 1] - 0.5) < 0.01 and abs(pose[b1, 1] - 1.5) < 0.015 and abs(pose[b2, 1] - 2.2) < 0.02
    assert abs(pose[rd, 1] - 0.6) < 0.01
    s = scenes.Scene("spin", gravity=(0.0, 0.0, 0.0))
    sp = s.insert(RigidBodyBuilder.dynamic().angvel((0.0, 3.0, 0.0)).linvel((0.5, 0.0, 0.0)), ColliderBuilder.convex_hull(tet))
    w = make_world(s)
    w.step(1)
    p0, _ = w.body_states()
    w.step(239)
    p1, v1 = w.body_states()

    def com(p):   # centre of mass (1/4, 1/4, 1/4) of the tetrahedron carried by the pose
        x, y, z, qw = p[3], p[4], p[5], p[6]
        v = np.array([0.25, 0.25, 0.25])
        q = np.array([x, y, z])
        tt = 2.0 * np.cross(q, v)
        return p[:3] + v + qw * tt + np.cross(q, tt)
    travelled = com(p1[sp]) - com(p0[sp])
    assert np.abs(travelled - (0.5 * 239 / 60.0, 0.0, 0.0)).max() < 2e-3, travelled   # the centre of mass moves straight
    assert abs(np.linalg.norm(v1[sp, 3:]) - 3.0) < 0.2
    s = scenes.Scene("heap", gravity=(0.0, -9.81, 0.0))
    s.insert(RigidBodyBuilder.fixed().translation((0.0, -0.5, 0.0)), ColliderBuilder.cuboid(30.0, 0.5, 30.0))
    r = np.random.default_rng(7)
    ids = []
    for i in range(30):
        pts = r.uniform(-0.5, 0.5, (10, 3))
        ids.append(s.insert(RigidBodyBuilder.dynamic().translation((0.4 * (i % 3), 1.0 + 1.2 * i, 0.3 * (i % 2))).rotation(tuple(r.uniform(-1, 1, 3))),
                            ColliderBuilder.round_convex_hull(pts, 0.03 * (i % 3))))
    w = make_world(s)
    w.step(700)
    pose, vel = w.body_states()
    assert np.isfinite(pose).all() and pose[ids, 1].min() > 0.0 and pose[ids, 1].max() < 6.0, (pose[ids, 1].min(), pose[ids, 1].max())
    assert np.abs(vel[ids]).max() < 1.0


def test_convex_polyhedra_oracle():
    convex_polyhedra(lambda s: oracle_lib.OracleWorld(s))


def test_convex_hull_builder_and_mesh_validation():
    """rb_convex_hull (host routine of the product library): hulls of random point sets are closed convex meshes (Euler's
    formula, every point behind every face), interior points are dropped, a cube's coplanar corners merge into 6 quads;
    rb_world_add_hull's checks are mirrored by the oracle's."""
    from rapier_b200.sets import convex_hull_mesh
    cube = [(x, y, z) for x in (-1.0, 1.0) for y in (-1.0, 1.0) for z in (-1.0, 1.0)]
    v, f = convex_hull_mesh(cube + [(0.0, 0.0, 0.0), (0.3, -0.2, 0.9)])
    assert len(v) == 8 and sorted(len(x) for x in f) == [4] * 6
    r = np.random.default_rng(3)
    for _ in range(20):
        pts = r.uniform(-1.0, 1.0, (int(r.integers(4, 20)), 3)).astype(np.float32)
        v, f = convex_hull_mesh(pts)
        e = sum(len(x) for x in f) // 2
        assert len(v) - e + len(f) == 2
        for loop in f:
            a, b, c = v[loop[0]], v[loop[1]], v[loop[2]]
            n = np.cross(b - a, c - a).astype(np.float64)
            n /= np.linalg.norm(n)
            assert ((pts - a) @ n).max() < 1e-4          # counter-clockwise seen from outside: everything behind
    with pytest.raises(ValueError):
        convex_hull_mesh([(0, 0, 0), (1, 0, 0), (0, 1, 0), (1, 1, 0)])   # coplanar
