"""Parity tests proper: the CUDA path (through the C ABI) against the CPU oracle on the same inputs.
Bit-exact at sizes the oracle finishes in seconds; size-independent properties at BASELINE sizes."""
import numpy as np
import pytest

import oracle_lib
from parity_util import compare_worlds, is_exact
from rapier_b200 import scenes
from rapier_b200.world import PhysicsWorld

pytestmark = pytest.mark.gpu

CASES = [
    ("pyramids_2x2x10", lambda: scenes.pyramids(2, 2, 10), 40, 10),
    ("pile_with_joint_chain", lambda: scenes.box_pile(4, 4, 5), 200, 25),
    ("single_pyramid_20", lambda: scenes.single_pyramid(20), 30, 10),
    ("pyramid3_large_island", lambda: scenes.pyramid3(10), 30, 10),
    ("joint_grid_large_island", lambda: scenes.joint_grid(20), 60, 20),
    ("ball_on_slab", lambda: scenes.box_on_ground("ball", 2.0), 100, 25),
    ("keva_1", lambda: scenes.keva(1), 30, 10),
    ("keva_2_large_island", lambda: scenes.keva(2), 12, 4),
    ("keva_5_full_size", lambda: scenes.keva(5), 4, 2),
    ("joint_grid_100_full_size", lambda: scenes.joint_grid(100), 8, 4),
    ("pyramid3_20_large_island", lambda: scenes.pyramid3(20), 30, 10),
    ("large_world_reduced", lambda: scenes.large_world(grid=60, spheres=16), 150, 30),
    ("large_world_300", lambda: scenes.large_world(grid=300, spheres=100), 40, 20),
]


@pytest.mark.parametrize("name,make,steps,every", CASES, ids=[c[0] for c in CASES])
def test_cuda_matches_oracle_bit_for_bit(built, name, make, steps, every):
    scene = make()
    w = PhysicsWorld(scene)
    o = oracle_lib.OracleWorld(scene)
    for i in range(steps):
        w.step()
        o.step()
        if i % every == every - 1 or i < 2:
            d = compare_worlds(w, o)
            assert is_exact(d), f"{name}: step {i}: {d}"


def test_many_pyramids_full_size_matches_oracle(built):
    """b3d_many_pyramids (reference file: 10 780 cubes): first steps bit-exact against the oracle."""
    scene = scenes.many_pyramids()
    w = PhysicsWorld(scene)
    o = oracle_lib.OracleWorld(scene, threads=8)
    for i in range(6):
        w.step()
        o.step()
    d = compare_worlds(w, o)
    assert is_exact(d), d
    c = w.counters()
    assert c["num_pairs"] == 196 * 145 and c["num_active_manifolds"] == 196 * 145


def test_headline_workload_80x20_matches_oracle(built):
    """BASELINE.json configs[1] (80 pyramids x 20 levels, 16 800 cubes): islands of 590 manifolds do not fit
    shared memory, so their constraint rows are streamed from the L2 pool by bulk copies.  The first steps
    run in the small launch shape (4 lanes / constraint), later ones in the big shape (1 lane / constraint)
    once the device hint has reached the host: all of them must match the oracle bit for bit."""
    scene = scenes.many_pyramids_label()
    w = PhysicsWorld(scene)
    o = oracle_lib.OracleWorld(scene, threads=8)
    for i in range(7):
        w.step()
        o.step()
        if i in (0, 1, 6):
            d = compare_worlds(w, o, tables=(i == 6))
            assert is_exact(d), (i, d)
    st = w.debug_read("state", np.int32)
    assert st[19] == 80, "the 80 pyramids must have been streamed"
    c = w.counters()
    assert c["num_active_manifolds"] == 80 * 590


@pytest.mark.parametrize("shape", ["0", "1"])
def test_streamed_item_in_both_launch_shapes(built, monkeypatch, shape):
    """A 20-level pyramid (streamed item) forced through the small (4 lanes) or the big (1 lane) launch shape."""
    monkeypatch.setenv("RB_COOP_SHAPE", shape)
    scene = scenes.single_pyramid(20)
    w = PhysicsWorld(scene)
    o = oracle_lib.OracleWorld(scene)
    for i in range(25):
        w.step()
        o.step()
        if i % 8 == 7 or i < 2:
            assert is_exact(compare_worlds(w, o)), i
    assert w.debug_read("state", np.int32)[19] == 1


def test_many_pyramids_full_size_properties(built):
    """Size-independent properties after 300 steps at full size: pyramids stay standing, the ground
    carries the whole weight (sum of ground-contact impulses = M g dt), state is finite, and a second
    run reproduces the first bit for bit."""
    scene = scenes.many_pyramids()
    w = PhysicsWorld(scene)
    p0, _ = w.body_states()
    w.step(300)
    p1, v1 = w.body_states()
    assert np.isfinite(p1).all() and np.isfinite(v1).all()
    assert np.abs(p1[:, :3] - p0[:, :3]).max() < 0.05
    assert np.abs(v1).max() < 0.02
    cp = w.contact_pairs()
    ground = cp["colliders"][:, 0] == 0
    total_weight_impulse = 10780 * 100.0 * 10.0 / 60.0   # 1 m^3 cubes, density 100, g = 10, dt = 1/60
    got = float(cp["impulses"][ground].sum())
    assert abs(got - total_weight_impulse) / total_weight_impulse < 0.01
    w2 = PhysicsWorld(scene)
    w2.step(300)
    p2, v2 = w2.body_states()
    assert (p1.view(np.uint32) == p2.view(np.uint32)).all() and (v1.view(np.uint32) == v2.view(np.uint32)).all()


def test_long_run_drift_vs_oracle(built):
    """north_star: <= 5 % pose drift vs the CPU path after 1000 steps (here the drift is exactly 0)."""
    scene = scenes.pyramids(2, 3, 10)
    w = PhysicsWorld(scene)
    o = oracle_lib.OracleWorld(scene, threads=4)
    w.step(1000)
    o.step(1000)
    pg, _ = w.body_states()
    po, _ = o.body_states()
    height = 10.0
    drift = np.linalg.norm(pg[:, :3] - po[:, :3], axis=1).max() / height
    assert drift <= 0.05
    assert (pg.view(np.uint32) == po.view(np.uint32)).all()


def test_full_size_dynamic_scene_1000_steps(built):
    """north_star drift contract on a full-size DYNAMIC scene: a falling, collapsing pile of 2 000 tilted cubes
    plus a jointed chain (broad phase, full narrow phase, colouring and schedule rebuilt on most steps),
    1 000 steps, GPU vs oracle: <= 5 % of the scene height, and in fact identical bits."""
    scene = scenes.box_pile(10, 10, 20)
    w = PhysicsWorld(scene)
    o = oracle_lib.OracleWorld(scene, threads=8)
    for _ in range(10):
        w.step(100)
        o.step(100)
        pg, vg = w.body_states()
        po, vo = o.body_states()
        assert np.isfinite(pg).all()
        height = float(po[:, 1].max() - po[:, 1].min()) or 1.0
        assert np.linalg.norm(pg[:, :3] - po[:, :3], axis=1).max() / height <= 0.05
    assert (pg.view(np.uint32) == po.view(np.uint32)).all() and (vg.view(np.uint32) == vo.view(np.uint32)).all()
    assert is_exact(compare_worlds(w, o))


def test_headline_1000_steps_drift(built):
    """north_star drift contract at the headline size (80 pyramids x 20 levels, 16 800 cubes), 1 000 steps."""
    scene = scenes.many_pyramids_label()
    w = PhysicsWorld(scene)
    o = oracle_lib.OracleWorld(scene, threads=16)
    w.step(1000)
    o.step(1000)
    pg, _ = w.body_states()
    po, _ = o.body_states()
    drift = np.linalg.norm(pg[:, :3] - po[:, :3], axis=1).max() / 20.0
    assert drift <= 0.05
    assert (pg.view(np.uint32) == po.view(np.uint32)).all()


def test_step_host_round_trip(built):
    """rb_world_step_host (host buffers in/out) gives the same trajectory as device-resident stepping."""
    scene = scenes.pyramids(1, 2, 6)
    a = PhysicsWorld(scene)
    b = PhysicsWorld(scene)
    b._flush()
    nb = len(scene.bodies)
    pose, vel = b.body_states()
    state = np.concatenate([pose, vel], axis=1).astype(np.float32).copy()
    out = np.zeros_like(state)
    for _ in range(20):
        a.step()
        b.physics_pipeline.step_host(scene.gravity, state, out)
        state, out = out, state
    pa, va = a.body_states()
    assert (state[:, :7].view(np.uint32) == pa.view(np.uint32)).all()
    assert (state[:, 7:].view(np.uint32) == va.view(np.uint32)).all()
    assert nb == state.shape[0]


def test_empty_and_ragged(built):
    from rapier_b200.sets import ColliderBuilder, RigidBodyBuilder
    PhysicsWorld(scenes.Scene("empty")).step(3)
    s = scenes.Scene("ragged")
    s.bodies.insert(RigidBodyBuilder.dynamic().translation((0, 5, 0)))
    s.insert(RigidBodyBuilder.dynamic().translation((3, 5, 0)), ColliderBuilder.ball(0.5))
    s.colliders.insert(ColliderBuilder.cuboid(5, 0.5, 5))
    w = PhysicsWorld(s)
    o = oracle_lib.OracleWorld(s)
    for _ in range(120):
        w.step()
        o.step()
    assert is_exact(compare_worlds(w, o))


from variant_cases import VARIANTS  # noqa: E402


GPU_VARIANTS = list(VARIANTS)


@pytest.mark.parametrize("name,make,params,steps,every", GPU_VARIANTS, ids=[v[0] for v in GPU_VARIANTS])
def test_cuda_variants_match_oracle(built, name, make, params, steps, every):
    """Parameter / feature edge cases (restitution, warm-start 0 / 0.5, friction in the bias pass,
    groups, disabled joint contacts, fixed joints, composite bodies, locked axes), bit for bit."""
    scene = make()
    w = PhysicsWorld(scene, integration_parameters=params)
    o = oracle_lib.OracleWorld(scene, params=params)
    for i in range(steps):
        w.step()
        o.step()
        if i % every == every - 1 or i < 2:
            d = compare_worlds(w, o)
            assert is_exact(d), f"{name}: step {i}: {d}"


def test_determinism_stress(built):
    """Race detector: the same full-size scene stepped asynchronously several times must give the
    same bits every time (same-colour constraints commute, islands are independent, so any run-to-run
    difference is a data race)."""
    scene = scenes.many_pyramids()
    ref = None
    for rep in range(6):
        w = PhysicsWorld(scene)
        w.step(150, sync=False)
        pose, vel = w.body_states()
        tables = {name: w.debug_read(name, dt).copy() for name, dt in [("pair_keys", np.uint64), ("pair_color", np.int32), ("pair_data", np.uint32)]}
        cur = (pose.view(np.uint32).copy(), vel.view(np.uint32).copy(), tables)
        if ref is None:
            ref = cur
        else:
            assert (cur[0] == ref[0]).all() and (cur[1] == ref[1]).all(), f"run {rep}: body state differs from run 0"
            for k in tables:
                assert cur[2][k].shape == ref[2][k].shape and (cur[2][k] == ref[2][k]).all(), f"run {rep}: table {k} differs"


def test_large_world_insertion_protocol(built):
    """b3d_large_world.rs:55-77 as written (reduced floor): one sphere inserted every 5 steps through rb_world_insert."""
    from incremental_cases import large_world_protocol_case
    large_world_protocol_case(grid=20, spheres=12, steps=100, every=10)


def test_large_world_full_size_insertion_protocol(built):
    """BASELINE configs[4] at the reference size: 10^6 static colliders (radix-sorted once), spheres inserted every
    5 steps, the first 60 steps bit for bit against the oracle."""
    from incremental_cases import large_world_protocol_case
    w, o = large_world_protocol_case(grid=1000, spheres=100, steps=60, every=20, threads=4)
    c = w.counters()
    assert c["num_colliders"] == 1000 * 1000 + 11 and c["num_bodies"] == 11


def test_body_removal_and_insertion(built):
    from incremental_cases import removal_case
    removal_case()


def test_capacity_overflow_is_reported_after_asynchronous_steps(built):
    from rapier_b200.sets import ColliderBuilder, RigidBodyBuilder
    from rapier_b200.world import RapierError
    s = scenes.Scene("crowd", gravity=(0.0, 0.0, 0.0))
    for i in range(330):
        s.insert(RigidBodyBuilder.dynamic().translation((0.001 * i, 0.0, 0.0)), ColliderBuilder.ball(0.5))
    w = PhysicsWorld(s)
    w.step(1, sync=False)
    with pytest.raises(RapierError, match="-4"):
        w.physics_pipeline.synchronize()
    w.physics_pipeline.synchronize()
    with pytest.raises(RapierError, match="-4"):
        w.physics_pipeline.step_host(s.gravity, None, None)


def test_whole_island_sleep(built):
    """crates/rapier3d/tests/whole_island_sleep.rs:40-83 on the CUDA path + wake by impact."""
    from test_oracle_kat import sleeping_stack_is_woken_by_an_impact, whole_island_blocks_partial_sleep
    whole_island_blocks_partial_sleep(lambda s: PhysicsWorld(s))
    sleeping_stack_is_woken_by_an_impact(lambda s: PhysicsWorld(s))


def test_sleep_and_wake_match_oracle_bit_for_bit(built):
    from rapier_b200.sets import ColliderBuilder, RigidBodyBuilder
    scene = scenes.box_pile(5, 5, 6)
    scene.insert(RigidBodyBuilder.dynamic().translation((0.0, 70.0, 0.3)).linvel((0.0, -2.0, 0.0)), ColliderBuilder.ball(0.5).density(5.0))
    w = PhysicsWorld(scene)
    o = oracle_lib.OracleWorld(scene, threads=4)
    seen_sleep = seen_wake = False
    prev = 0
    for i in range(400):
        w.step(); o.step()
        if i % 5 == 4:
            sw, so = w.sleeping(), o.sleeping()
            assert (sw == so).all(), i
            n = int(sw.sum())
            seen_sleep = seen_sleep or n > 20
            seen_wake = seen_wake or (prev > 20 and n < prev)
            prev = n
        if i % 40 == 39:
            assert is_exact(compare_worlds(w, o)), i
    assert seen_sleep and seen_wake


def test_quarantine(built):
    from test_oracle_kat import nan_force_is_quarantined
    nan_force_is_quarantined(lambda s: PhysicsWorld(s), expect_error=True)


def test_coulomb_friction_model(built):
    """FrictionModel::Coulomb on hardware: the friction-cone known answers through the C ABI, and the reference
    file's b3d_many_pyramids (10 780 cubes, every island through the streaming solve) bit-exact against the oracle."""
    from test_oracle_kat import _coulomb_params, coulomb_friction_cone
    coulomb_friction_cone(lambda s: PhysicsWorld(s, integration_parameters=_coulomb_params()))
    scene = scenes.many_pyramids()
    w = PhysicsWorld(scene, integration_parameters=_coulomb_params())
    o = oracle_lib.OracleWorld(scene, params=_coulomb_params(), threads=8)
    for i in range(4):
        w.step()
        o.step()
    d = compare_worlds(w, o)
    assert is_exact(d), d


def test_ccd_motion_clamping(built):
    """CCD known answers of the reference's test-suite (ccd_default_vs_fixed.rs, issue_217_ccd_large_dt_hitch.rs) through the C ABI."""
    from test_oracle_kat import ccd_bullet_still_hits_dynamic, ccd_default_tier, ccd_large_dt_no_mid_air_hitch
    mk = lambda s, p: PhysicsWorld(s, integration_parameters=p)
    ccd_default_tier(mk)
    ccd_large_dt_no_mid_air_hitch(mk)
    ccd_bullet_still_hits_dynamic(mk)


def test_events(built):
    """Collision / contact-force events through the C ABI: the reference's threshold-crossing known answer
    (contact_force_event_first_tick.rs) and event lists identical to the oracle's on a collapsing pile."""
    from test_oracle_kat import contact_force_event_started_marks_threshold_crossings
    from variant_cases import events_parity_case
    contact_force_event_started_marks_threshold_crossings(lambda s: PhysicsWorld(s))
    events_parity_case(lambda s: PhysicsWorld(s), lambda s: oracle_lib.OracleWorld(s))


def test_kinematic_bodies(built):
    """Kinematic bodies through the C ABI: the behavioural known answers (platform / lift / pair filter, issue_287) and a
    bit-exact run of a turntable, a conveyor and a position-driven lift against the oracle."""
    from test_oracle_kat import kinematic_bodies, moving_kinematic_wakes_jointed_dynamic
    from variant_cases import kinematic_parity_case
    kinematic_bodies(lambda s: PhysicsWorld(s))
    moving_kinematic_wakes_jointed_dynamic(lambda s: PhysicsWorld(s))
    kinematic_parity_case(lambda s: PhysicsWorld(s), lambda s: oracle_lib.OracleWorld(s))


def test_joint_limits_and_motors(built):
    """Joint limits and motors through the C ABI: issue_499_angular_limits.rs cases, prismatic limits / position spring /
    bounded motor, and the generic-row scene bit-exact against the oracle under both friction models."""
    from test_oracle_kat import angular_limits_are_reached, prismatic_limits_and_position_motor
    from variant_cases import joint_limits_parity_case
    angular_limits_are_reached(lambda s: PhysicsWorld(s), cases=((-45.0, 45.0), (0.0, 270.0), (45.0, 315.0)))
    prismatic_limits_and_position_motor(lambda s: PhysicsWorld(s))
    mk = lambda s, p=None: PhysicsWorld(s, integration_parameters=p)
    mo = lambda s, p=None: oracle_lib.OracleWorld(s, params=p)
    joint_limits_parity_case(mk, mo)
    joint_limits_parity_case(mk, mo, coulomb=True)


def test_capsules(built):
    """Capsules through the C ABI (SHAPES = 1 variant of the collision kernel): rest heights / two-point manifolds / total impulse."""
    from test_oracle_kat import capsules_rest
    capsules_rest(lambda s: PhysicsWorld(s))


def test_more_reference_known_answers(built):
    """sleep_wake.rs scenarios, issue_746 (prismatic axis frames) and issue_856 (stiff position motor on a re-oriented base)
    through the C ABI."""
    from test_oracle_kat import (motor_position_with_rotating_base_stays_finite, prismatic_joint_stays_bounded_for_all_axis_rotations,
                                 sleep_wake_scenarios)

    def mk(s):
        w = PhysicsWorld(s)
        w.reserve(32, 32)
        return w
    sleep_wake_scenarios(mk)
    mkp = lambda s, p: PhysicsWorld(s, integration_parameters=p)
    prismatic_joint_stays_bounded_for_all_axis_rotations(mkp)
    motor_position_with_rotating_base_stays_finite(mkp)


def test_dominance_groups_and_joint_warmstart(built):
    """RigidBodyDominance and IntegrationParameters::warmstart_joints: known answers through the C ABI and bit-exact parity
    with the oracle (twist and Coulomb friction; limits / motors and a locked-only joint grid on the generic joint path)."""
    from test_oracle_kat import dominance_groups, joint_warmstart
    from variant_cases import dominance_parity_case, joint_limits_parity_case
    mk = lambda s, p=None: PhysicsWorld(s, integration_parameters=p)
    mo = lambda s, p=None: oracle_lib.OracleWorld(s, params=p)
    dominance_groups(mk)
    dominance_parity_case(mk, mo)
    joint_warmstart(mk)
    joint_limits_parity_case(mk, mo, warmstart_joints=True)
    joint_limits_parity_case(mk, mo, warmstart_joints=True, coulomb=True, steps=60)
    joint_limits_parity_case(mk, mo, warmstart_joints=True, scene=scenes.joint_grid(20), steps=60)


def test_coupled_joint_axes(built):
    """GenericJoint::coupled_axes -- SpringJoint / RopeJoint (one row on the distance between the anchors) and the cone limit of
    two coupled angular axes: the reference's joint_contact_solve_order.rs and issue_792 scenarios plus closed-form answers
    through the C ABI, and a scene mixing them with contacts bit-exact against the oracle (twist and Coulomb + warmstart_joints)."""
    from test_oracle_kat import coupled_angular_spring_joint_stays_finite, heavy_cubes_rest_on_spring_jointed_balls, spring_and_rope_joints
    from variant_cases import coupled_axes_parity_case
    mk = lambda s, p=None: PhysicsWorld(s, integration_parameters=p)
    mo = lambda s, p=None: oracle_lib.OracleWorld(s, params=p)
    spring_and_rope_joints(mk)
    heavy_cubes_rest_on_spring_jointed_balls(mk)
    coupled_angular_spring_joint_stays_finite(mk)
    coupled_axes_parity_case(mk, mo)
    coupled_axes_parity_case(mk, mo, coulomb=True, warmstart_joints=True, steps=60)


def test_additional_solver_iterations(built):
    """RigidBody::additional_solver_iterations (substep solve-groups, island_manager/substep_groups.rs): the scenarios of
    crates/rapier3d/tests/additional_solver_iterations.rs through the C ABI, and islands of three cadences (jointed, on a kinematic
    platform, bouncing, a grid-wide one) bit-exact against the oracle -- one pass of the general solve kernels per cadence."""
    from test_oracle_kat import additional_solver_iterations
    from variant_cases import substep_groups_parity_case
    mk = lambda s, p=None: PhysicsWorld(s, integration_parameters=p)
    mo = lambda s, p=None: oracle_lib.OracleWorld(s, params=p)
    additional_solver_iterations(mk)
    substep_groups_parity_case(mk, mo, steps=60)
    substep_groups_parity_case(mk, mo, steps=40, big=False, coulomb=True, warmstart_joints=True)


def test_sensors(built):
    """Collider::is_sensor: intersection-only pairs with CollisionEventFlags::SENSOR events -- free fall through a zone bit-exact
    with the zone-less twin, an aura that reports the ground early, a thin sensor wall a fast body ignores -- through the C ABI, and a
    pile falling through zones of every shape family bit-exact against the oracle (states, pair tables, flagged event lists)."""
    from test_oracle_kat import sensors
    from variant_cases import sensors_parity_case
    sensors(lambda s: PhysicsWorld(s))
    sensors_parity_case(lambda s: PhysicsWorld(s), lambda s: oracle_lib.OracleWorld(s))


def test_compound_bodies(built):
    """Multi-collider bodies with general mass properties (full tensor diagonalised by the host): sleep_wide_bodies.rs (64 U-shaped
    compounds all asleep after 300 steps, a still far-reaching body neither moves nor stays awake), a diagonal dumbbell spinning
    about its principal axis / conserving angular momentum, an L shape at rest -- through the C ABI; a pile of U / L / dumbbell /
    three-shape compounds bit-exact against the oracle."""
    from test_oracle_kat import compound_bodies
    from variant_cases import compound_pile
    compound_bodies(lambda s: PhysicsWorld(s))
    s = compound_pile()
    w, o = PhysicsWorld(s), oracle_lib.OracleWorld(s)
    for i in range(200):
        w.step(); o.step()
        if i % 20 == 19 or i < 2:
            d = compare_worlds(w, o)
            assert is_exact(d), (i, d)


def test_convex_polyhedra(built):
    """ColliderBuilder::{convex_hull, round_convex_hull}: known answers through the C ABI (the convex_pile parity variants
    run with the other variants) and the reference's examples3d/convex_polyhedron3.rs drop (reduced: 5 x 5 x 4 round hulls of
    10 random points) bit-exact against the oracle."""
    from test_oracle_kat import convex_polyhedra
    convex_polyhedra(lambda s: PhysicsWorld(s))
    s = scenes.convex_polyhedra(4)
    w, o = PhysicsWorld(s), oracle_lib.OracleWorld(s, threads=8)
    for i in range(150):
        w.step(); o.step()
        if i % 15 == 14 or i < 2:
            d = compare_worlds(w, o)
            assert is_exact(d), (i, d)
