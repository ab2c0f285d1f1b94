"""Scene / parameter variants shared by the emulated-kernel tests (CPU) and the CUDA parity tests (GPU):
edge cases of the reference's own test-suite for this path (restitution, warm-start coefficients,
friction in the bias pass, collision groups, joints with contacts disabled, multi-collider bodies,
user forces, damping, locked axes)."""
import math

from rapier_b200 import _abi as A
from rapier_b200 import scenes
import numpy as np
from rapier_b200.sets import ColliderBuilder, FixedJointBuilder, RigidBodyBuilder, SphericalJointBuilder


def _params(**kw):
    p = A.RbIntegrationParameters.default()
    for k, v in kw.items():
        setattr(p, k, v)
    return p


def bouncing_balls():
    """issue_974_restitution.rs scene with several restitutions side by side + a bouncing cube."""
    s = scenes.Scene("bounce")
    s.insert(RigidBodyBuilder.fixed(), ColliderBuilder.cuboid(30.0, 0.1, 30.0).restitution(0.7))
    for i, e in enumerate((0.0, 0.3, 0.8, 1.0)):
        s.insert(RigidBodyBuilder.dynamic().translation((2.0 * i, 1.3 + 0.2 * i, 0.0)), ColliderBuilder.ball(0.2).restitution(e))
    s.insert(RigidBodyBuilder.dynamic().translation((-3.0, 1.0, 0.0)).rotation((0.3, 0.2, 0.1)),
             ColliderBuilder.cuboid(0.3, 0.2, 0.25).restitution(0.6))
    return s


def groups_and_joints():
    """Collision-group filtering (interaction_groups.rs:148-154), a joint with contacts disabled
    (pair_update.rs:193-202), a fixed joint (6 locked axes), user force/torque, damping, locked axes."""
    s = scenes.Scene("groups")
    s.insert(RigidBodyBuilder.fixed().translation((0, -0.5, 0)), ColliderBuilder.cuboid(20, 0.5, 20))
    # two overlapping boxes that must NOT collide (disjoint groups) and one that collides with both
    a = s.insert(RigidBodyBuilder.dynamic().translation((0, 1.0, 0)), ColliderBuilder.cuboid(0.5, 0.5, 0.5).collision_groups(0b01, 0b01))
    b = s.insert(RigidBodyBuilder.dynamic().translation((0.3, 1.2, 0)), ColliderBuilder.cuboid(0.5, 0.5, 0.5).collision_groups(0b10, 0b10))
    s.insert(RigidBodyBuilder.dynamic().translation((0.1, 3.0, 0.1)), ColliderBuilder.cuboid(0.4, 0.4, 0.4))
    # overlapping balls joined by a spherical joint with contacts disabled
    c = s.insert(RigidBodyBuilder.dynamic().translation((5, 2, 0)), ColliderBuilder.ball(0.5))
    d = s.insert(RigidBodyBuilder.dynamic().translation((5.6, 2, 0)).linear_damping(0.5).angular_damping(1.0), ColliderBuilder.ball(0.5))
    s.joints.insert(c, d, SphericalJointBuilder().local_anchor1((0.3, 0, 0)).local_anchor2((-0.3, 0, 0)).contacts_enabled(False))
    # a fixed joint welding two boxes, one of them pushed by a user force / torque
    e = s.insert(RigidBodyBuilder.dynamic().translation((-5, 2, 0)), ColliderBuilder.cuboid(0.5, 0.25, 0.25))
    f = s.insert(RigidBodyBuilder.dynamic().translation((-4, 2, 0)), ColliderBuilder.cuboid(0.5, 0.25, 0.25))
    s.joints.insert(e, f, FixedJointBuilder().local_anchor1((0.5, 0, 0)).local_anchor2((-0.5, 0, 0)))
    s.bodies.descs[f].user_force[:] = (0.0, 0.0, 3.0)
    s.bodies.descs[f].user_torque[:] = (0.2, 0.0, 0.0)
    # a body with two colliders (composite mass properties) and a body with locked rotations
    g = s.insert(RigidBodyBuilder.dynamic().translation((0, 2, 6)), ColliderBuilder.cuboid(0.5, 0.2, 0.2).translation((0.6, 0, 0)))
    s.colliders.insert_with_parent(ColliderBuilder.cuboid(0.5, 0.2, 0.2).translation((-0.6, 0, 0)), g)
    s.insert(RigidBodyBuilder.dynamic().translation((3, 2, 6)).rotation((0.4, 0.0, 0.2)).locked_axes(A.RB_BODY_LOCK_RX | A.RB_BODY_LOCK_RZ),
             ColliderBuilder.cuboid(0.3, 0.6, 0.3))
    assert a >= 0 and b >= 0
    return s


def additional_mass_twins():
    """issue_78 / issue_666 scenes side by side: mass from collider density vs RigidBody::additional_mass on a
    massless collider (host-side mass properties; tests/test_oracle_kat.py holds the known-answer checks)."""
    from rapier_b200.sets import ColliderBuilder, RigidBodyBuilder
    s = scenes.Scene("additional_mass_twins")
    s.insert(RigidBodyBuilder.fixed().translation((0.0, -0.5, 0.0)), ColliderBuilder.cuboid(30.0, 0.5, 30.0).friction(0.5))
    for i, additional in enumerate((False, True, True)):
        body = RigidBodyBuilder.dynamic().translation((-6.0 + 6.0 * i, 1.0, 0.0)).rotation((0.0, 0.0, 0.7))
        col = ColliderBuilder.cuboid(0.5, 0.5, 0.5).friction(0.5)
        if additional:
            body = body.additional_mass(100.0)
            col = col.density(0.0 if i == 1 else 50.0)   # massless collider / extra mass on top of a dense one
        else:
            col = col.density(100.0)
        s.insert(body, col)
    s.insert(RigidBodyBuilder.dynamic().translation((12.0, 3.0, 0.0)).additional_mass(2.0).linvel((0.0, 0.0, 1.0)), ColliderBuilder.ball(0.4).density(0.0))
    return s


def plate_with_overflow_colour():
    """A dynamic plate carrying 156 small cubes: the plate has more contacts than the 120 dyn-dyn colours, so 36 of
    them land in the overflow colour 128 (contact_pair.rs:155), which is solved serially after the parallel colours."""
    from rapier_b200.sets import ColliderBuilder, RigidBodyBuilder
    s = scenes.Scene("plate_overflow")
    s.insert(RigidBodyBuilder.fixed().translation((0.0, -0.5, 0.0)), ColliderBuilder.cuboid(20.0, 0.5, 20.0))
    s.insert(RigidBodyBuilder.dynamic().translation((0.0, 0.3, 0.0)), ColliderBuilder.cuboid(7.0, 0.25, 7.0).density(5.0))
    for i in range(13):
        for k in range(12):
            s.insert(RigidBodyBuilder.dynamic().translation((-6.0 + i * 1.0, 0.86, -5.5 + k * 1.0)), ColliderBuilder.cuboid(0.3, 0.3, 0.3))
    return s


def shuffled_collider_order():
    """Bodies inserted first, colliders attached afterwards in REVERSE body order, so the pair table's
    (collider1, collider2) order is the opposite of the reference's colouring order (min body, max body)
    (contacts.rs:366-385): a stack of boxes next to a column of balls, all touching from step 0."""
    s = scenes.Scene("shuffled_colliders")
    ground = s.bodies.insert(RigidBodyBuilder.fixed().translation((0.0, -0.5, 0.0)))
    hs = []
    for k in range(3):
        for i in range(4):
            for j in range(4 - i):
                hs.append((s.bodies.insert(RigidBodyBuilder.dynamic().translation((i * 0.5 + j * 1.0, 0.5 + i * 1.0, k * 1.0))), "box"))
    for i in range(5):
        hs.append((s.bodies.insert(RigidBodyBuilder.dynamic().translation((8.0, 0.4 + 0.8 * i, 0.0))), "ball"))
    for h, kind in reversed(hs):
        s.colliders.insert_with_parent(ColliderBuilder.cuboid(0.5, 0.5, 0.5) if kind == "box" else ColliderBuilder.ball(0.4), h)
    s.colliders.insert_with_parent(ColliderBuilder.cuboid(20.0, 0.5, 20.0), ground)
    return s


def ccd_barrage():
    """Fast spinning cubes and balls fired at thin fixed walls, a tiled floor (many narrow static colliders + one wide
    slab: both static lists of the broad phase) and each other: CCD motion clamping (src/dynamics/ccd) of several
    bodies per step, single- and multi-collider bodies, impacts from the first step on."""
    s = scenes.Scene("ccd_barrage", gravity=(0.0, -9.81, 0.0))
    s.insert(RigidBodyBuilder.fixed().translation((0.0, -0.55, 0.0)), ColliderBuilder.cuboid(40.0, 0.05, 40.0))
    for i in range(12):
        for k in range(12):
            s.colliders.insert(ColliderBuilder.cuboid(0.5, 0.04, 0.5).translation((-6.0 + i * 1.0, -0.46, -6.0 + k * 1.0)))
    s.insert(RigidBodyBuilder.fixed().translation((8.0, 2.0, 0.0)), ColliderBuilder.cuboid(0.04, 3.0, 8.0))
    s.insert(RigidBodyBuilder.fixed().translation((-8.0, 2.0, 0.0)).rotation((0.0, 0.0, 0.3)), ColliderBuilder.cuboid(0.04, 3.0, 8.0))
    for i in range(6):
        s.insert(RigidBodyBuilder.dynamic().translation((-2.0 + 0.7 * i, 1.0 + 0.5 * i, -2.0 + i)).linvel((150.0 - 40.0 * i, -30.0 * i, 10.0))
                 .angvel((3.0, 1.0 * i, -2.0)), ColliderBuilder.cuboid(0.1, 0.15, 0.2))
        s.insert(RigidBodyBuilder.dynamic().translation((1.0 - 0.7 * i, 3.0 + 0.5 * i, 2.0 - i)).linvel((-120.0 + 30.0 * i, -60.0, -5.0 * i)).ccd_enabled(i % 2 == 0),
                 ColliderBuilder.ball(0.12 + 0.02 * i))   # (every other ball is a bullet: it also sweeps against the moving bodies)
    g = s.insert(RigidBodyBuilder.dynamic().translation((0.0, 6.0, 4.0)).linvel((90.0, -90.0, 0.0)).angvel((0.0, 0.0, 8.0)),
                 ColliderBuilder.cuboid(0.3, 0.05, 0.05).translation((0.3, 0.0, 0.0)))
    s.colliders.insert_with_parent(ColliderBuilder.ball(0.08).translation((-0.3, 0.0, 0.0)), g)
    for i in range(4):   # slow targets in the line of fire, and a bullet cube aimed at them
        s.insert(RigidBodyBuilder.dynamic().translation((3.0 + 0.9 * i, 0.0, 5.0)), ColliderBuilder.cuboid(0.3, 0.45, 0.3))
    s.insert(RigidBodyBuilder.dynamic().translation((-5.0, 0.1, 5.0)).linvel((180.0, 0.0, 0.0)).angvel((0.0, 0.0, 5.0)).ccd_enabled(True), ColliderBuilder.cuboid(0.1, 0.12, 0.08))
    return s


def events_scene():
    """A pile collapsing onto a slab, a bouncing ball and a body that leaves the scene: every collider asks for collision
    and contact-force events (thresholds from 0 to 40 N), so starts, stops, pairs that leave the broad phase while
    touching and force-threshold crossings all occur."""
    ev = A.RB_EVENT_COLLISION | A.RB_EVENT_CONTACT_FORCE
    s = scenes.Scene("events")
    s.insert(RigidBodyBuilder.fixed().translation((0.0, -0.5, 0.0)), ColliderBuilder.cuboid(6.0, 0.5, 6.0).active_events(ev).contact_force_event_threshold(5.0))
    for i in range(4):
        for k in range(3):
            s.insert(RigidBodyBuilder.dynamic().translation((0.3 * i - 0.4, 0.6 + 1.05 * i, 0.35 * k)).rotation((0.1 * k, 0.2, 0.05 * i)),
                     ColliderBuilder.cuboid(0.5, 0.5, 0.5).active_events(ev).contact_force_event_threshold(10.0 * (i % 3)))
    s.insert(RigidBodyBuilder.dynamic().translation((3.0, 2.0, 0.0)), ColliderBuilder.ball(0.3).restitution(0.8).active_events(A.RB_EVENT_COLLISION))
    s.insert(RigidBodyBuilder.dynamic().translation((5.5, 0.5, 0.0)).linvel((6.0, 0.0, 0.0)), ColliderBuilder.cuboid(0.4, 0.4, 0.4).active_events(ev))
    return s


def events_parity_case(make_world, make_oracle, steps=150):
    """The drained event lists of both worlds must be identical (ids, started flags, step, and every float bit)."""
    s = events_scene()
    w, o = make_world(s), make_oracle(s)
    nc = nf = 0
    for i in range(steps):
        w.step(); o.step()
        if i % 7 == 6 or i == steps - 1:   # (drained every few steps: the buffers accumulate across steps)
            cw, co = w.collision_events(), o.collision_events()
            assert cw == co, (i, cw[:4], co[:4])
            fw, fo = w.contact_force_events(), o.contact_force_events()
            assert len(fw) == len(fo), (i, len(fw), len(fo))
            for a, b in zip(fw, fo):
                assert a == b, (i, a, b)
            nc += len(cw); nf += len(fw)
    assert nc > 20 and nf > 50, (nc, nf)
    return nc, nf


def sensors_scene():
    """Sensor colliders of every shape family among a falling pile: fixed zones (cuboid, ball, capsule, round hull), a zone that
    leaves the scene with its kinematic carrier, massless and massive sensor auras on dynamic bodies (sensor-sensor pairs too), a
    sensor floor tile next to solid ones, bodies that sleep inside a zone."""
    ev = A.RB_EVENT_COLLISION
    s = scenes.Scene("sensors")
    s.insert(RigidBodyBuilder.fixed().translation((0.0, -0.5, 0.0)), ColliderBuilder.cuboid(12.0, 0.5, 12.0))
    s.colliders.insert(ColliderBuilder.cuboid(3.0, 0.6, 3.0).translation((0.0, 2.0, 0.0)).sensor(True).active_events(ev))
    s.colliders.insert(ColliderBuilder.ball(1.2).translation((2.5, 1.0, 0.5)).sensor(True).active_events(ev))
    s.colliders.insert(ColliderBuilder.capsule_x(1.5, 0.4).translation((-2.0, 0.8, 1.0)).sensor(True))
    s.colliders.insert(ColliderBuilder.round_convex_hull([(-1, 0, -1), (1, 0, -1), (1, 0, 1), (-1, 0, 1), (0, 1.5, 0)], 0.05).translation((0.0, 0.2, -3.0)).sensor(True).active_events(ev))
    s.insert(RigidBodyBuilder.kinematic_velocity_based().translation((-5.0, 1.5, 0.0)).linvel((1.5, 0.0, 0.0)), ColliderBuilder.cuboid(0.5, 1.5, 4.0).sensor(True).active_events(ev))
    k = 0
    for layer in range(4):
        for i in range(3):
            for j in range(3):
                pos = (-1.5 + 1.5 * i + 0.11 * layer, 3.5 + 1.2 * layer, -1.5 + 1.5 * j + 0.07 * i)
                b = RigidBodyBuilder.dynamic().translation(pos).rotation((0.2 * (k % 3), 0.1 * (k % 5), 0.15 * (k % 4)))
                kind = k % 4
                c = (ColliderBuilder.cuboid(0.4, 0.3, 0.35) if kind == 0 else ColliderBuilder.ball(0.35) if kind == 1 else
                     ColliderBuilder.capsule_y(0.3, 0.2) if kind == 2 else ColliderBuilder.cuboid(0.3, 0.3, 0.3))
                h = s.insert(b, c.active_events(ev if k % 2 else 0))
                if k % 5 == 0:
                    s.colliders.insert_with_parent(ColliderBuilder.ball(0.8).density(0.0 if k % 10 else 0.5).sensor(True).active_events(ev), h)
                k += 1
    return s


def sensors_parity_case(make_world, make_oracle, steps=200, every=20):
    """Body states, pair tables and the drained event lists (with their SENSOR flags) of both worlds must be identical."""
    from parity_util import compare_worlds, is_exact
    s = sensors_scene()
    w, o = make_world(s), make_oracle(s)
    nsens = 0
    for i in range(steps):
        w.step(); o.step()
        if i % 7 == 6 or i == steps - 1:
            cw, co = w.collision_events(with_flags=True), o.collision_events(with_flags=True)
            assert cw == co, (i, cw[:4], co[:4])
            nsens += sum(1 for e in cw if e[4])
        if i % every == every - 1 or i < 2:
            d = compare_worlds(w, o)
            assert is_exact(d), (i, d)
    assert nsens > 30, nsens


def compound_pile():
    """Compound bodies with general mass properties (full inertia tensor diagonalised on the host): U shapes (stress_tests/compound3.rs),
    L shapes of two densities, rotated dumbbells, a ball-and-stick of three shape families -- dropped in a pile with spin."""
    s = scenes.Scene("compound_pile")
    s.insert(RigidBodyBuilder.fixed().translation((0.0, -0.5, 0.0)), ColliderBuilder.cuboid(20.0, 0.5, 20.0))
    k = 0
    for layer in range(3):
        for i in range(3):
            for j in range(2):
                pos = (-3.0 + 3.0 * i + 0.2 * layer, 1.5 + 2.6 * layer, -1.5 + 3.0 * j + 0.1 * i)
                b = s.bodies.insert(RigidBodyBuilder.dynamic().translation(pos).rotation((0.3 * (k % 3), 0.2 * (k % 4), 0.1 * (k % 5))).angvel((0.5 * (k % 2), 0.0, 0.3 * (k % 3))))
                kind = k % 4
                if kind == 0:    # U
                    s.colliders.insert_with_parent(ColliderBuilder.cuboid(1.0, 0.15, 0.15), b)
                    s.colliders.insert_with_parent(ColliderBuilder.cuboid(0.15, 0.6, 0.15).translation((1.0, 0.6, 0.0)), b)
                    s.colliders.insert_with_parent(ColliderBuilder.cuboid(0.15, 0.6, 0.15).translation((-1.0, 0.6, 0.0)), b)
                elif kind == 1:  # L, heavy foot
                    s.colliders.insert_with_parent(ColliderBuilder.cuboid(0.8, 0.15, 0.15), b)
                    s.colliders.insert_with_parent(ColliderBuilder.cuboid(0.15, 0.15, 0.6).translation((0.65, 0.0, 0.75)).density(3.0), b)
                elif kind == 2:  # dumbbell on a diagonal, parts rotated
                    s.colliders.insert_with_parent(ColliderBuilder.cuboid(0.4, 0.2, 0.2).translation((0.6, 0.6, 0.0)).rotation((0.0, 0.0, 0.7853981633974483)), b)
                    s.colliders.insert_with_parent(ColliderBuilder.cuboid(0.4, 0.2, 0.2).translation((-0.6, -0.6, 0.0)).rotation((0.0, 0.0, 0.7853981633974483)), b)
                else:            # ball + capsule + box
                    s.colliders.insert_with_parent(ColliderBuilder.ball(0.3).translation((0.0, 0.7, 0.2)), b)
                    s.colliders.insert_with_parent(ColliderBuilder.capsule_y(0.4, 0.12), b)
                    s.colliders.insert_with_parent(ColliderBuilder.cuboid(0.3, 0.1, 0.3).translation((0.1, -0.5, 0.0)).rotation((0.0, 0.4, 0.0)), b)
                k += 1
    return s


def kinematic_parity_case(make_world, make_oracle, steps=150, every=15):
    """A velocity-based kinematic turntable and conveyor carrying boxes and balls, and a position-based lift driven
    along a curve with a new target every step: bit-exact against the oracle."""
    import math
    from parity_util import compare_worlds, is_exact
    s = scenes.Scene("kinematic_parity")
    s.insert(RigidBodyBuilder.fixed().translation((0.0, -1.0, 0.0)), ColliderBuilder.cuboid(30.0, 0.5, 30.0))
    s.insert(RigidBodyBuilder.kinematic_velocity_based().translation((0.0, 0.0, 0.0)).angvel((0.0, 0.8, 0.0)), ColliderBuilder.cuboid(3.0, 0.2, 3.0).friction(0.9))
    s.insert(RigidBodyBuilder.kinematic_velocity_based().translation((8.0, 0.0, 0.0)).linvel((0.0, 0.0, 1.0)), ColliderBuilder.cuboid(1.5, 0.2, 6.0))
    lift = s.insert(RigidBodyBuilder.kinematic_position_based().translation((-8.0, 0.0, 0.0)), ColliderBuilder.cuboid(1.5, 0.2, 1.5))
    for i in range(4):
        s.insert(RigidBodyBuilder.dynamic().translation((1.2 * math.cos(i * 1.6), 0.75 + 0.05 * i, 1.2 * math.sin(i * 1.6))), ColliderBuilder.cuboid(0.4, 0.5, 0.3))
        s.insert(RigidBodyBuilder.dynamic().translation((8.0 + 0.3 * i, 0.7 + 0.9 * i, -2.0 + 0.2 * i)), ColliderBuilder.ball(0.35) if i % 2 else ColliderBuilder.cuboid(0.3, 0.4, 0.3))
        s.insert(RigidBodyBuilder.dynamic().translation((-8.0 + 0.1 * i, 0.65 + 0.85 * i, 0.1 * i)), ColliderBuilder.cuboid(0.4, 0.4, 0.4))
    w, o = make_world(s), make_oracle(s)
    for i in range(steps):
        t = (i + 1) / 60.0
        h = 0.5 * math.sin(0.5 * t)
        target = [(-8.0 + 0.5 * math.sin(t), 1.0 - math.cos(1.5 * t), 0.3 * t, 0.0, math.sin(h), 0.0, math.cos(h))]
        for world in (w, o):
            world.set_next_kinematic_positions([lift], target)
            world.step()
        if i % every == every - 1 or i < 2:
            d = compare_worlds(w, o)
            assert is_exact(d), (i, d)
    return w


def joint_limits_scene():
    """Every generic joint row kind next to contacts: revolute joints with angular limits and velocity / position motors
    (force- and acceleration-based), prismatic joints with linear limits and motors, a spherical joint with angular motors on
    all three axes, locked-only joints (fixed, spherical) in the same world, a jointed chain resting on the ground."""
    from rapier_b200.sets import PrismaticJointBuilder, RevoluteJointBuilder
    s = scenes.Scene("joint_limits")
    s.insert(RigidBodyBuilder.fixed().translation((0.0, -0.5, 0.0)), ColliderBuilder.cuboid(30.0, 0.5, 30.0))
    base = s.bodies.insert(RigidBodyBuilder.fixed().translation((0.0, 4.0, 0.0)))
    prev = base
    for i in range(5):   # a chain of bars on limited revolute joints, the last ones motorised
        b = s.insert(RigidBodyBuilder.dynamic().translation((1.0 + 1.0 * i, 4.0, 0.0)), ColliderBuilder.cuboid(0.45, 0.1, 0.1))
        j = RevoluteJointBuilder((0.0, 0.0, 1.0)).local_anchor1((0.5 if i else 0.0, 0.0, 0.0)).local_anchor2((-0.5, 0.0, 0.0)).limits(3, -0.6 - 0.1 * i, 0.3 + 0.2 * i)
        if i == 3:
            j = j.motor_velocity(3, 2.0, 5.0).motor_max_force(3, 40.0)
        if i == 4:
            j = j.motor_position(3, 0.4, 30.0, 3.0).motor_model(3, 1)
        s.joints.insert(prev, b, j)
        prev = b
    for k in range(3):   # sliders: limits, a position spring, a velocity motor against a limit
        b = s.insert(RigidBodyBuilder.dynamic().translation((-4.0, 1.0 + 1.5 * k, 2.0 * k)), ColliderBuilder.cuboid(0.3, 0.3, 0.3))
        j = PrismaticJointBuilder((0.3, 1.0, 0.1)).local_anchor1((-4.0, 1.5 + 1.5 * k, 2.0 * k)).limits(0, -0.8, 0.6)
        if k == 1:
            j = j.motor_position(0, 0.3, 60.0, 6.0)
        if k == 2:
            j = j.motor_velocity(0, 1.5, 10.0)
        s.joints.insert(base, b, j)
    ball = s.insert(RigidBodyBuilder.dynamic().translation((5.0, 3.0, 0.0)), ColliderBuilder.cuboid(0.4, 0.2, 0.3))
    j = SphericalJointBuilder().local_anchor1((5.0, -0.5, 0.0)).local_anchor2((0.0, 0.5, 0.0))
    for ax in (3, 4, 5):
        j = j.motor(ax, 0.2 * (ax - 3), 0.5, 20.0, 2.0).limits(ax, -1.0, 1.2) if ax != 4 else j.motor_velocity(ax, 1.0, 4.0)
    s.joints.insert(base, ball, j)
    a = s.insert(RigidBodyBuilder.dynamic().translation((8.0, 0.5, 0.0)), ColliderBuilder.cuboid(0.5, 0.5, 0.5))
    b = s.insert(RigidBodyBuilder.dynamic().translation((9.0, 0.5, 0.0)), ColliderBuilder.cuboid(0.5, 0.5, 0.5))
    s.joints.insert(a, b, FixedJointBuilder().local_anchor1((0.5, 0.0, 0.0)).local_anchor2((-0.5, 0.0, 0.0)))
    c = s.insert(RigidBodyBuilder.dynamic().translation((9.0, 1.6, 0.0)), ColliderBuilder.ball(0.5))
    s.joints.insert(b, c, SphericalJointBuilder().local_anchor1((0.0, 0.55, 0.0)).local_anchor2((0.0, -0.55, 0.0)))
    return s


def coupled_axes_scene():
    """Coupled joint axes next to contacts and the other generic rows: spring joints (force- and acceleration-based, with and
    without a distance limit), rope joints between dynamic bodies and to the world, a net of springs carrying a plate that boxes
    fall on, cone limits (two coupled angular axes, all three pairings) on bodies that tumble, a spring with only two coupled
    linear axes next to a limited third one."""
    from rapier_b200.sets import GenericJointBuilder, RopeJointBuilder, SpringJointBuilder
    s = scenes.Scene("coupled_axes")
    s.insert(RigidBodyBuilder.fixed().translation((0.0, -0.5, 0.0)), ColliderBuilder.cuboid(30.0, 0.5, 30.0))
    base = s.bodies.insert(RigidBodyBuilder.fixed().translation((0.0, 6.0, 0.0)))
    prev = base
    for i in range(5):   # a rope of rope joints released sideways, then a chain of springs under it
        b = s.insert(RigidBodyBuilder.dynamic().translation((0.9 * (i + 1), 6.0, 0.0)).can_sleep(False), ColliderBuilder.ball(0.2))
        s.joints.insert(prev, b, RopeJointBuilder(1.0))
        prev = b
    for i in range(3):
        b = s.insert(RigidBodyBuilder.dynamic().translation((0.9 * 5 + 0.3, 5.0 - i, 0.0)).can_sleep(False), ColliderBuilder.cuboid(0.2, 0.2, 0.2))
        j = SpringJointBuilder(0.8, 120.0, 4.0 + i)
        if i == 1:
            j = j.motor_model(0, 0)
        if i == 2:
            j = j.limits(0, 0.0, 1.1)   # spring with a hard maximum length (motor_linear_coupled clamps its target velocity too)
        s.joints.insert(prev, b, j)
        prev = b
    plate = s.insert(RigidBodyBuilder.dynamic().translation((-6.0, 2.0, 0.0)).can_sleep(False), ColliderBuilder.cuboid(1.5, 0.1, 1.5))
    for k, (dx, dz) in enumerate(((-1.4, -1.4), (1.4, -1.4), (-1.4, 1.4), (1.4, 1.4))):
        s.joints.insert(base, plate, SpringJointBuilder(1.0, 300.0, 10.0 + k).local_anchor1((-6.0 + 1.3 * dx, -2.5, 1.3 * dz)).local_anchor2((dx, 0.0, dz)))
    for k in range(3):
        s.insert(RigidBodyBuilder.dynamic().translation((-6.3 + 0.3 * k, 3.0 + 0.9 * k, 0.1 * k)), ColliderBuilder.cuboid(0.3, 0.3, 0.3))
    for k, mask in enumerate((0b101000, 0b011000, 0b110000)):   # cone limits about Y, Z, X
        b = s.insert(RigidBodyBuilder.dynamic().translation((6.0 + 2.5 * k, 3.0, 4.0)).angvel((1.5 - k, 0.4 * k, 1.0 + 0.3 * k)).can_sleep(False),
                     ColliderBuilder.cuboid(0.5, 0.3, 0.4))
        first = 3 if mask & 0b001000 else 4
        j = GenericJointBuilder(0b000111).local_anchor1((6.0 + 2.5 * k, -2.0, 4.0)).local_anchor2((0.0, 1.0, 0.0)).coupled_axes(mask).limits(first, 0.0 if k else -0.1, 0.4 + 0.1 * k)
        if k == 2:
            j = j.motor(4, 0.0, 0.0, 5.0, 0.5)   # coupled angular motor: no row
        s.joints.insert(base, b, j)
    b = s.insert(RigidBodyBuilder.dynamic().translation((0.0, 3.0, -5.0)).can_sleep(False), ColliderBuilder.ball(0.4))
    s.joints.insert(base, b, GenericJointBuilder(0).local_anchor1((0.0, -1.5, -5.0)).coupled_axes(0b000101).motor_position(0, 1.0, 80.0, 5.0).limits(1, -0.3, 0.6))
    return s


def coupled_axes_parity_case(make_world, make_oracle, steps=150, every=15, **kw):
    joint_limits_parity_case(make_world, make_oracle, steps=steps, every=every, scene=coupled_axes_scene(), **kw)


def substep_groups_scene(big=True):
    """Substep solve-groups: islands of three cadences side by side (0, 3 and 8 extra substeps; the key of an island is the
    largest additional_solver_iterations among its members), one of them jointed, one riding a kinematic platform, one with
    restitution, a wall of bricks large enough for the grid-wide item with 2 extra substeps on a single brick, and plain stacks."""
    s = scenes.Scene("substep_groups", gravity=(0.0, -9.81, 0.0))
    s.insert(RigidBodyBuilder.fixed().translation((0.0, -0.5, 0.0)), ColliderBuilder.cuboid(60.0, 0.5, 60.0))
    for k, extra in enumerate((0, 3, 8, 0, 3)):   # heavy-on-light stacks
        x = -12.0 + 4.0 * k
        s.insert(RigidBodyBuilder.dynamic().translation((x, 0.5, 0.0)), ColliderBuilder.cuboid(0.5, 0.5, 0.5))
        s.insert(RigidBodyBuilder.dynamic().translation((x + 0.05, 1.6, 0.02)), ColliderBuilder.cuboid(0.45, 0.5, 0.45).restitution(0.4 if k == 1 else 0.0))
        s.insert(RigidBodyBuilder.dynamic().translation((x, 3.0, 0.0)).additional_solver_iterations(extra), ColliderBuilder.cuboid(0.5, 0.5, 0.5).density(50.0))
    base = s.bodies.insert(RigidBodyBuilder.fixed().translation((0.0, 8.0, 6.0)))
    prev = base
    for i in range(6):   # rope with a heavy end: 8 extra substeps for the whole chain, a box resting against it joins the island
        last = i == 5
        b = RigidBodyBuilder.dynamic().translation((0.6 * (i + 1), 8.0, 6.0))
        if last:
            b = b.additional_solver_iterations(8)
        link = s.insert(b, ColliderBuilder.ball(0.25).density(60.0 if last else 1.0))
        s.joints.insert(prev, link, SphericalJointBuilder().local_anchor1((0.3 if i else 0.0, 0.0, 0.0)).local_anchor2((-0.3, 0.0, 0.0)))
        prev = link
    s.kin = s.insert(RigidBodyBuilder.kinematic_velocity_based().translation((10.0, 1.0, 6.0)).linvel((0.4, 0.0, 0.0)), ColliderBuilder.cuboid(2.0, 0.2, 2.0))
    s.insert(RigidBodyBuilder.dynamic().translation((10.0, 1.7, 6.0)), ColliderBuilder.cuboid(0.5, 0.5, 0.5))
    s.insert(RigidBodyBuilder.dynamic().translation((10.0, 2.75, 6.0)).additional_solver_iterations(3), ColliderBuilder.cuboid(0.5, 0.5, 0.5).density(40.0))
    if big:   # a brick wall: one island above the item caps
        for i in range(14):
            for j in range(24):
                b = RigidBodyBuilder.dynamic().translation((-20.0 + 1.0 * j + (0.5 if i % 2 else 0.0), 0.25 + 0.5 * i, -8.0))
                if i == 3 and j == 5:
                    b = b.additional_solver_iterations(2)
                s.insert(b, ColliderBuilder.cuboid(0.5, 0.25, 0.3))
    return s


def substep_groups_parity_case(make_world, make_oracle, steps=60, every=10, big=True, **kw):
    joint_limits_parity_case(make_world, make_oracle, steps=steps, every=every, scene=substep_groups_scene(big), **kw)


def joint_limits_parity_case(make_world, make_oracle, steps=150, every=15, coulomb=False, warmstart_joints=False, scene=None):
    from parity_util import compare_worlds, is_exact
    s = scene if scene is not None else joint_limits_scene()
    kw = {}
    if coulomb:
        kw["friction_model"] = 1
    if warmstart_joints:
        kw["warmstart_joints"] = 1
    p = _params(**kw) if kw else None
    w, o = (make_world(s, p), make_oracle(s, p)) if kw else (make_world(s), make_oracle(s))
    for i in range(steps):
        w.step(); o.step()
        if i % every == every - 1 or i < 2:
            d = compare_worlds(w, o)
            assert is_exact(d), (i, d)
            ji_w, ji_o = w.debug_read("joint_impulses", np.float32), o.debug_read("joint_impulses", np.float32)
            assert (ji_w.view(np.uint32) == ji_o.view(np.uint32)).all(), i
    pose, _ = w.body_states()
    assert np.isfinite(pose).all()


def capsule_pile():
    """Capsules of all three axes, cuboids and balls dropped together: capsule-capsule (crossed and parallel), capsule-cuboid
    (face clipping, edges, ends) and capsule-ball manifolds, recycling and warm start across them, a two-collider body
    (capsule + box), parentless capsules as static geometry."""
    s = scenes.Scene("capsule_pile")
    s.insert(RigidBodyBuilder.fixed().translation((0.0, -0.5, 0.0)), ColliderBuilder.cuboid(15.0, 0.5, 15.0))
    s.colliders.insert(ColliderBuilder.capsule_z(4.0, 0.4).translation((-2.5, 0.4, 0.0)))
    s.colliders.insert(ColliderBuilder.capsule_z(4.0, 0.4).translation((2.5, 0.4, 0.0)))
    k = 0
    for layer in range(4):
        for i in range(3):
            for j in range(2):
                pos = (-1.2 + 1.2 * i + 0.07 * layer, 0.8 + 1.1 * layer, -0.7 + 1.4 * j + 0.05 * i)
                b = RigidBodyBuilder.dynamic().translation(pos).rotation((0.3 * ((k * 7) % 5 - 2), 0.2 * ((k * 3) % 7 - 3), 0.25 * ((k * 5) % 3 - 1)))
                kind = k % 5
                c = (ColliderBuilder.capsule_y(0.35, 0.2) if kind == 0 else ColliderBuilder.capsule_x(0.5, 0.25) if kind == 1 else
                     ColliderBuilder.cuboid(0.4, 0.3, 0.35) if kind == 2 else ColliderBuilder.ball(0.3) if kind == 3 else ColliderBuilder.capsule_z(0.3, 0.3))
                s.insert(b, c.restitution(0.2 if k % 4 == 0 else 0.0))
                k += 1
    g = s.insert(RigidBodyBuilder.dynamic().translation((0.0, 6.0, 0.0)), ColliderBuilder.capsule_x(0.6, 0.2).translation((0.0, 0.3, 0.0)))
    s.colliders.insert_with_parent(ColliderBuilder.cuboid(0.3, 0.15, 0.3).translation((0.0, -0.1, 0.0)), g)
    return s


def convex_pile():
    """Convex polyhedra against everything: random hulls (sharp and round) tumbling into a heap on a floor of a cuboid plus
    a static prism, hull-cubes stacked on hull-cubes / cuboids, a ball and capsules of all axes dropped among them,
    a tetrahedron, a hexagonal prism (6-vertex faces: clipped polygons beyond eight points), a kinematic hull sweeping
    through, a hull on a spherical joint."""
    s = scenes.Scene("convex_pile")
    s.insert(RigidBodyBuilder.fixed().translation((0.0, -0.5, 0.0)), ColliderBuilder.cuboid(15.0, 0.5, 15.0))
    hexagon = [(math.cos(k * math.pi / 3.0), y, math.sin(k * math.pi / 3.0)) for k in range(6) for y in (-0.3, 0.3)]
    s.colliders.insert(ColliderBuilder.convex_hull([(2.0 * x, 1.5 * y + 0.45, 2.0 * z) for x, y, z in hexagon]).translation((0.0, 0.0, 4.0)))
    cube = [(x, y, z) for x in (-0.5, 0.5) for y in (-0.5, 0.5) for z in (-0.5, 0.5)]
    tet = [(0.0, 0.0, 0.0), (1.0, 0.0, 0.0), (0.0, 1.0, 0.0), (0.0, 0.0, 1.0)]
    r = np.random.default_rng(11)
    k = 0
    for layer in range(5):
        for i in range(3):
            for j in range(2):
                pos = (-1.3 + 1.3 * i + 0.07 * layer, 0.9 + 1.25 * layer, -0.7 + 1.4 * j + 0.05 * i)
                b = RigidBodyBuilder.dynamic().translation(pos).rotation(tuple(float(x) for x in r.uniform(-1.0, 1.0, 3)))
                kind = k % 6
                if kind == 0:
                    c = ColliderBuilder.convex_hull(r.uniform(-0.55, 0.55, (10, 3)))
                elif kind == 1:
                    c = ColliderBuilder.round_convex_hull(r.uniform(-0.45, 0.45, (8, 3)), 0.08)
                elif kind == 2:
                    c = ColliderBuilder.convex_hull([(0.6 * x, 0.5 * y, 0.6 * z) for x, y, z in hexagon])
                elif kind == 3:
                    c = ColliderBuilder.cuboid(0.4, 0.3, 0.5) if layer % 2 else ColliderBuilder.ball(0.45)
                elif kind == 4:
                    c = [ColliderBuilder.capsule_x, ColliderBuilder.capsule_y, ColliderBuilder.capsule_z][layer % 3](0.4, 0.25)
                else:
                    c = ColliderBuilder.convex_hull(tet).density(3.0)
                s.insert(b, c.friction(0.3 + 0.1 * (k % 5)).restitution(0.2 if k % 7 == 0 else 0.0))
                k += 1
    for i in range(3):   # stacks: hull cube on hull cube on cuboid
        s.insert(RigidBodyBuilder.dynamic().translation((6.0, 0.5 + 1.0 * i, 0.0)), ColliderBuilder.convex_hull(cube) if i else ColliderBuilder.cuboid(0.5, 0.5, 0.5))
    s.insert(RigidBodyBuilder.kinematic_velocity_based().translation((-8.0, 0.6, 0.0)).linvel((1.5, 0.0, 0.0)), ColliderBuilder.convex_hull([(0.4 * x, 1.2 * y, 3.0 * z) for x, y, z in cube]))
    anchor = s.bodies.insert(RigidBodyBuilder.fixed().translation((-4.0, 5.0, -5.0)))
    swing = s.insert(RigidBodyBuilder.dynamic().translation((-2.5, 5.0, -5.0)), ColliderBuilder.convex_hull(r.uniform(-0.5, 0.5, (12, 3))))
    s.joints.insert(anchor, swing, SphericalJointBuilder().local_anchor2((-1.5, 0.0, 0.0)))
    return s


VARIANTS = [
    ("restitution", bouncing_balls, None, 150, 25),
    ("groups_joints_forces", groups_and_joints, None, 150, 25),
    ("warmstart_half", lambda: scenes.pyramids(1, 2, 6), _params(warmstart_coefficient=0.5), 40, 10),
    ("warmstart_zero", lambda: scenes.pyramids(1, 2, 6), _params(warmstart_coefficient=0.0), 40, 10),
    ("friction_in_bias_pass", lambda: scenes.box_pile(3, 3, 3), _params(friction_in_bias_pass=1), 120, 20),
    ("two_pgs_no_relax", lambda: scenes.box_pile(3, 2, 3), _params(num_internal_pgs_iterations=2, num_internal_stabilization_iterations=0), 100, 20),
    ("six_substeps_no_recycling", lambda: scenes.pyramids(1, 1, 8), _params(num_solver_iterations=6, contact_recycling=0), 40, 10),
    ("length_unit_10", lambda: scenes.box_pile(2, 3, 2), _params(length_unit=10.0), 80, 20),
    # scenes of the reference's joint known-answer tests (tests/test_oracle_kat.py): revolute joints on bodies with an
    # offset centre of mass (issue_952), and the 1000:1 chain with 20 substeps (substep_chain_high_mass_ratio)
    ("revolute_offset_com", lambda: scenes.offset_com_pendulums(16), None, 80, 20),
    ("heavy_chain_20_substeps", scenes.heavy_end_chain, _params(num_solver_iterations=20), 60, 20),
    ("additional_mass_twins", additional_mass_twins, None, 120, 20),
    ("overflow_colour", plate_with_overflow_colour, None, 60, 15),
    ("shuffled_collider_order", shuffled_collider_order, None, 60, 15),
    # FrictionModel::Coulomb (contact_with_coulomb_friction.rs): one coupled tangent part per contact point
    ("coulomb_pile_with_joints", lambda: scenes.box_pile(3, 3, 3), _params(friction_model=1), 120, 20),
    ("coulomb_restitution", bouncing_balls, _params(friction_model=1), 120, 20),
    ("coulomb_groups_joints_forces", groups_and_joints, _params(friction_model=1), 120, 20),
    ("coulomb_pyramids_warmstart_half", lambda: scenes.pyramids(1, 2, 6), _params(friction_model=1, warmstart_coefficient=0.5), 40, 10),
    ("coulomb_warmstart_zero_friction_in_bias", lambda: scenes.box_pile(2, 3, 2), _params(friction_model=1, warmstart_coefficient=0.0, friction_in_bias_pass=1), 60, 15),
    ("coulomb_large_island", lambda: scenes.pyramid3(9), _params(friction_model=1), 25, 5),
    ("coulomb_overflow_colour", plate_with_overflow_colour, _params(friction_model=1), 40, 10),
    # capsules (SHAPES = 1 variant of the collision kernel)
    ("capsule_pile", capsule_pile, None, 180, 20),
    ("capsule_pile_coulomb_no_recycling", capsule_pile, _params(friction_model=1, contact_recycling=0), 60, 15),
    # convex polyhedra (rb_poly.cuh)
    ("convex_pile", convex_pile, None, 200, 20),
    ("convex_pile_coulomb_no_recycling", convex_pile, _params(friction_model=1, contact_recycling=0), 60, 15),
    # CCD motion clamping (src/dynamics/ccd): fast bodies against thin fixed walls / a tiled floor; and switched off
    ("ccd_barrage", ccd_barrage, None, 90, 10),
    ("ccd_barrage_ccd_off", ccd_barrage, _params(max_ccd_substeps=0), 30, 10),
    ("ccd_barrage_coulomb_dt_large", ccd_barrage, _params(friction_model=1, dt=1.0 / 30.0), 40, 10),
]


def dominance_scene():
    """RigidBodyDominance next to ordinary contacts: a stack whose middle box dominates (its neighbours see it as immovable,
    it still rests on the one below), a heavy dominated box dropped on a light dominating one, a dynamic box dominating the
    kinematic platform that sweeps into it, dominated balls rolling against a dominating box, two jointed boxes of
    different groups, and a small pile of mixed groups (a group changes the constraint's world-attached side, not the
    colouring)."""
    s = scenes.Scene("dominance")
    s.insert(RigidBodyBuilder.fixed().translation((0.0, -0.5, 0.0)), ColliderBuilder.cuboid(30.0, 0.5, 30.0))
    for k, g in enumerate((0, 5, 0)):
        s.insert(RigidBodyBuilder.dynamic().translation((0.02 * k, 0.5 + 1.0 * k, 0.0)).dominance_group(g), ColliderBuilder.cuboid(0.5, 0.5, 0.5))
    s.insert(RigidBodyBuilder.dynamic().translation((4.0, 0.3, 0.0)).dominance_group(10), ColliderBuilder.cuboid(0.3, 0.3, 0.3).density(0.5))
    s.insert(RigidBodyBuilder.dynamic().translation((4.1, 2.0, 0.05)).dominance_group(-3), ColliderBuilder.cuboid(0.6, 0.6, 0.6).density(20.0))
    s.insert(RigidBodyBuilder.kinematic_velocity_based().translation((-6.0, 0.5, 0.0)).linvel((1.5, 0.0, 0.0)).dominance_group(-1),
             ColliderBuilder.cuboid(0.4, 0.5, 1.0))
    s.insert(RigidBodyBuilder.dynamic().translation((-4.0, 0.4, 0.0)).dominance_group(2), ColliderBuilder.cuboid(0.4, 0.4, 0.4))
    s.insert(RigidBodyBuilder.dynamic().translation((-2.5, 0.4, 0.2)), ColliderBuilder.cuboid(0.4, 0.4, 0.4))
    s.insert(RigidBodyBuilder.dynamic().translation((0.0, 0.6, 5.0)).dominance_group(127), ColliderBuilder.cuboid(0.6, 0.6, 0.6))
    for i in range(3):
        s.insert(RigidBodyBuilder.dynamic().translation((-3.0 - 1.2 * i, 0.4, 5.0 + 0.1 * i)).linvel((4.0, 0.0, 0.0)).dominance_group(-128 + i),
                 ColliderBuilder.ball(0.4))
    a = s.insert(RigidBodyBuilder.dynamic().translation((8.0, 0.5, 0.0)).dominance_group(1), ColliderBuilder.cuboid(0.5, 0.5, 0.5))
    b = s.insert(RigidBodyBuilder.dynamic().translation((9.1, 0.5, 0.0)).dominance_group(-1), ColliderBuilder.cuboid(0.5, 0.5, 0.5))
    s.joints.insert(a, b, SphericalJointBuilder().local_anchor1((0.55, 0.5, 0.0)).local_anchor2((-0.55, 0.5, 0.0)))
    s.insert(RigidBodyBuilder.dynamic().translation((8.5, 1.6, 0.0)), ColliderBuilder.cuboid(0.45, 0.45, 0.45))
    k = 0
    for layer in range(3):
        for i in range(3):
            for j in range(2):
                pos = (-1.0 + 1.05 * i + 0.06 * layer, 0.5 + 1.02 * layer, -10.0 + 1.05 * j + 0.04 * i)
                s.insert(RigidBodyBuilder.dynamic().translation(pos).dominance_group((k * 5) % 4 - 1),
                         ColliderBuilder.cuboid(0.5, 0.5, 0.5) if k % 3 else ColliderBuilder.ball(0.5))
                k += 1
    return s


def dominance_parity_case(make_world, make_oracle, steps=180, every=12):
    from parity_util import compare_worlds, is_exact
    s = dominance_scene()
    for coulomb in (False, True):
        p = _params(friction_model=1) if coulomb else None
        w, o = (make_world(s, p), make_oracle(s, p)) if coulomb else (make_world(s), make_oracle(s))
        for i in range(steps):
            w.step(); o.step()
            if i % every == every - 1 or i < 2:
                d = compare_worlds(w, o)
                assert is_exact(d), (coulomb, i, d)
        pose, _ = w.body_states()
        assert np.isfinite(pose).all()
