"""Unit-level vectors taken from the reference tree (tests/golden/ref_vectors.json, generator and provenance in
tests/golden/make_ref_vectors.py) against the three implementations of this repo: the CPU oracle and the emulated
kernels (CPU suite) and the CUDA library (-m gpu).  "asserted" = what the Rust test asserts; "derived" = the cited
Rust function evaluated by hand (numpy f32, no FMA), compared within the stated tolerance."""
import ctypes as C
import json
import os

import numpy as np
import pytest

import emul_lib
import oracle_lib

HERE = os.path.dirname(os.path.abspath(__file__))
DOC = json.load(open(os.path.join(HERE, "golden", "ref_vectors.json")))
NOUT = {"pose_drift": 1, "reduce_manifold": 5, "normal_solve": 13, "tangent_solve": 14, "generate": 39, "recentered_angle": 1, "combine_coeff": 1}


def kat_oracle(name, x):
    x = np.ascontiguousarray(x, np.float32)
    out = np.zeros(NOUT[name], np.float32)
    rc = oracle_lib.lib().orc_kat(name.encode(), x.ctypes.data, len(x), out.ctypes.data, len(out))
    assert rc == 0, (name, rc)
    return out


def kat_lib(L):
    def run(name, x):
        x = np.ascontiguousarray(x, np.float32)
        out = np.zeros(NOUT[name], np.float32)
        rc = L.rb_debug_kat(name.encode(), x.ctypes.data, len(x), out.ctypes.data, len(out))
        assert rc == 0, (name, rc, L.rb_last_error())
        return out
    return run


def close(a, b, tol):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return bool((np.abs(a - b) <= tol * np.maximum(1.0, np.abs(b))).all())


def check_all(run):
    # generate(): contact_with_twist_friction.rs:845-981
    for v in DOC["generate"]:
        o = run("generate", v["input"])
        d, a, tol = v["derived"], v["asserted"], v["tol"]
        n = d["num_contacts"]
        assert int(o[0]) == n
        assert close(o[1:4], d["dir1"], 0) and close(o[4:7], d["tangent1"], 0) and close(o[7], d["limit"], 0)
        assert close(o[8:12], d["impulse"], tol) and close(o[12:16], d["impulse_accumulator"], tol) and close(o[16:20], d["r"], tol)
        assert close(o[20:24], d["dist"], tol) and close(o[24:28], d["twist_dists"], tol)
        assert close(o[28:30], d["tangent_impulse"], tol) and close(o[30:32], [-x for x in d["tangent_impulse"]], tol)
        assert close(o[32], d["twist_impulse"], tol) and close(o[33], -d["twist_impulse"], tol)
        assert [int(x) for x in o[35:39]] == d["manifold_contact_id"]
        for k in range(n, 4):   # the Rust's literal assertions on inactive slots
            assert o[16 + k] == a["inactive_slots"]["r"] and o[8 + k] == a["inactive_slots"]["impulse"] and o[24 + k] == a["inactive_slots"]["twist_dist"]
            assert int(o[35 + k]) == a["inactive_slots"]["manifold_contact_id"]
        if n == 1:
            assert o[32] == a["single_point_twist_impulse"]
    # relative_pose_drift(): contact_pair.rs:869-900
    v = DOC["pose_drift"]
    worst = max(float(run("pose_drift", x)[0]) for x in v["inputs"])
    assert worst <= v["asserted"]["max_drift"], worst
    # ContactConstraintNormalPart::solve: contact_constraint_element.rs:481-504
    for v in DOC["normal_solve"]:
        o = run("normal_solve", v["input"])
        assert close(o, v["derived"], v["tol"]), (v["name"], o, v["derived"])
        if "impulse" in v["asserted"]:
            assert o[0] == v["asserted"]["impulse"], v["name"]
        if v["asserted"].get("velocities_unchanged"):
            assert (o[1:13] == np.asarray(v["input"][25:37], np.float32)).all(), v["name"]
    # ContactConstraintTangentPart::solve: contact_constraint_element.rs:903-938
    for v in DOC["tangent_solve"]:
        o = run("tangent_solve", v["input"])
        assert np.isfinite(o).all()
        assert (o[0:2] == np.asarray(v["asserted"]["impulse"], np.float32)).all(), o[:2]
        assert (o[2:14] == np.asarray(v["asserted"]["velocities"], np.float32)).all(), o
    # reduce_manifold_naive: manifold_reduction.rs:4-84
    for v in DOC["reduce_manifold"]:
        o = run("reduce_manifold", v["input"])
        assert int(o[0]) == v["derived"]["num_selected"], (v["input"], o)
        assert [int(x) for x in o[1:5]] == v["derived"]["selected"], (v["input"], o)


def test_reference_vectors_oracle():
    check_all(kat_oracle)


def test_reference_vectors_emulated_kernels():
    check_all(kat_lib(emul_lib.lib()))


@pytest.mark.gpu
def test_reference_vectors_cuda(built):
    from rapier_b200._lib import lib
    check_all(kat_lib(lib()))


# ---- the colouring walk (narrow_phase/mod.rs:87-152 applied in the order of contacts.rs:366-385) ----
def colour_walk(pairs, is_dynamic):
    """Sequential restatement: `pairs` = [(body1, body2)] of the pairs that begin touching (None = parentless
    collider), processed in (min body, max body, index) order; returns the colour of each."""
    U = 0xffffffff
    order = sorted(range(len(pairs)), key=lambda i: (min(U if b is None else b for b in pairs[i]), max(U if b is None else b for b in pairs[i]), i))
    masks = {}
    colour = [255] * len(pairs)
    for i in order:
        b1, b2 = pairs[i]
        c1 = b1 if (b1 is not None and is_dynamic[b1]) else None
        c2 = b2 if (b2 is not None and is_dynamic[b2]) else None
        if c1 is None and c2 is None:
            colour[i] = 128
            continue
        if c1 is not None and c2 is not None:
            mask = masks.get(c1, 0) | masks.get(c2, 0)
            free = ~mask & ((1 << 120) - 1)
            col = (free & -free).bit_length() - 1 if free else 128
            bodies = [c1, c2]
        else:
            b = c1 if c1 is not None else c2
            inv = ~masks.get(b, 0) & ((1 << 128) - 1)
            col = inv.bit_length() - 1 if inv else 128   # 127 - leading_zeros(!mask)
            bodies = [b]
        if col >= 128:
            colour[i] = 128
            continue
        for b in bodies:
            masks[b] = masks.get(b, 0) | (1 << col)
        colour[i] = col
    return colour


def _colour_case(world_factory):
    from rapier_b200 import scenes
    from variant_cases import plate_with_overflow_colour, shuffled_collider_order
    for make in (lambda: scenes.pyramids(2, 2, 6), shuffled_collider_order, plate_with_overflow_colour, lambda: scenes.keva(1)):
        scene = make()
        w = world_factory(scene)
        w.step()   # every touching pair begins touching in the first step: one colouring walk from empty masks
        keys = w.debug_read("pair_keys", np.uint64)
        nsc = w.debug_read("pair_nsc", np.int32)
        col = w.debug_read("pair_color", np.int32)
        parent = [c.parent for c in scene.colliders.descs]
        dyn = [b.body_type == 0 for b in scene.bodies.descs]
        touching = [i for i in range(len(keys)) if nsc[i] > 0]
        pairs = []
        for i in touching:
            c1, c2 = int(keys[i] >> np.uint64(32)), int(keys[i] & np.uint64(0xffffffff))
            pairs.append((parent[c1] if parent[c1] >= 0 else None, parent[c2] if parent[c2] >= 0 else None))
        want = colour_walk(pairs, dyn)
        got = [int(col[i]) for i in touching]
        assert got == want, scene.name
        assert all(int(col[i]) == 255 for i in range(len(keys)) if nsc[i] == 0)


def test_colouring_walk_oracle():
    _colour_case(lambda s: oracle_lib.OracleWorld(s))


def test_colouring_walk_emulated_kernels():
    from rapier_b200.world import PhysicsWorld
    _colour_case(lambda s: PhysicsWorld(s, _lib=emul_lib.lib()))


@pytest.mark.gpu
def test_colouring_walk_cuda(built):
    from rapier_b200.world import PhysicsWorld
    _colour_case(lambda s: PhysicsWorld(s))


# ---- the colour buckets of the constraint schedule (solver_contact_graph.rs:113-233; its tests :245-334 assert
# that every live entry sits in the bucket of its colour exactly once, through inserts, removals and recolours) ----
def _schedule_invariants(w):
    nsc = w.debug_read("pair_nsc", np.int32)
    col = w.debug_read("pair_color", np.int32)
    cons_pair = w.debug_read("sched_cons_pair", np.int32)
    starts = w.debug_read("sched_item_cons_start", np.int32)
    offs = w.debug_read("sched_item_color_off", np.int32).reshape(len(starts) - 1, 130)
    cpos = w.debug_read("sched_color_pos", np.int32)
    active = set(np.nonzero(nsc > 0)[0].tolist())
    assert len(cons_pair) == len(active) and set(cons_pair.tolist()) == active     # every solver-active manifold exactly once
    pos_to_colour = {int(p): c for c, p in enumerate(cpos[:129]) if p >= 0}
    for it in range(len(starts) - 1):
        for stage, colour in pos_to_colour.items():
            for q in range(starts[it] + offs[it, stage], starts[it] + offs[it, stage + 1]):
                assert min(int(col[cons_pair[q]]), 128) == colour               # ... in the bucket of its colour


def _schedule_case(factory):
    from rapier_b200 import scenes
    from variant_cases import plate_with_overflow_colour
    for make, steps in ((lambda: scenes.box_pile(4, 4, 5), (1, 20, 45, 90)), (plate_with_overflow_colour, (2, 30)), (lambda: scenes.pyramid3(9), (1, 12, 30))):
        w = factory(make())
        done = 0
        for s in steps:     # contacts begin, end and change colour as the pile falls and settles
            w.step(s - done)
            done = s
            _schedule_invariants(w)


def test_schedule_colour_buckets_emulated_kernels():
    from rapier_b200.world import PhysicsWorld
    _schedule_case(lambda s: PhysicsWorld(s, _lib=emul_lib.lib()))


@pytest.mark.gpu
def test_schedule_colour_buckets_cuda(built):
    from rapier_b200.world import PhysicsWorld
    _schedule_case(lambda s: PhysicsWorld(s))


def test_recentered_angle_slope_is_one_everywhere():
    """joint_constraint_helper.rs:804-848 (the reference's own unit test of the angular-limit measure), on the oracle's
    restatement -- whose arctangent is the explicit polynomial the kernels share: zero at the centre of the range, +half_range at
    the upper limit, slope one everywhere on the circle except across the wrap at the antipode of the centre."""
    import math
    ang = lambda theta, lo, hi: float(kat_oracle("recentered_angle", [theta, lo, hi])[0])
    for center_deg in (-180, -135, -90, 0, 45, 135, 180):
        lo, hi = math.radians(center_deg - 45.0), math.radians(center_deg + 45.0)
        assert abs(ang(math.radians(center_deg), lo, hi)) < 1.0e-5, center_deg
        assert abs(ang(math.radians(center_deg + 45.0), lo, hi) - math.radians(45.0)) < 1.0e-4, center_deg
        for theta_deg in range(-350, 351, 7):
            theta, eps = math.radians(theta_deg), 1.0e-3
            a0, a1 = ang(theta - eps, lo, hi), ang(theta + eps, lo, hi)
            if abs(a1 - a0) > 1.0:
                continue   # the wrap discontinuity itself
            assert abs((a1 - a0) / (2.0 * eps) - 1.0) < 1.0e-2, (center_deg, theta_deg, (a1 - a0) / (2.0 * eps))


def test_coefficient_combine_rules():
    """coefficient_combine_rule.rs:96-130 (the reference's unit tests, literal values): the geometric mean, its clamping of
    negative coefficients, and rule priority (the larger rule wins)."""
    from rapier_b200 import _abi as A
    comb = lambda c1, c2, r1, r2: float(kat_oracle("combine_coeff", [c1, c2, r1, r2])[0])
    g, avg = A.RB_COMBINE_GEOMETRIC_MEAN, A.RB_COMBINE_AVERAGE
    assert comb(0.25, 1.0, g, g) == 0.5
    assert comb(0.7, 0.0, g, g) == 0.0
    assert comb(-0.5, 0.5, g, g) == 0.0
    assert comb(-0.5, -0.5, g, g) == 0.0
    assert comb(0.25, 1.0, g, avg) == 0.5
    assert comb(0.25, 1.0, avg, avg) == 0.625
