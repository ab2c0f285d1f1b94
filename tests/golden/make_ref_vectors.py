"""Generator of tests/golden/ref_vectors.json: unit-level vectors taken from the REFERENCE tree.

The reference binary cannot run here (no Rust toolchain), so these are not outputs of it.  Each vector carries
  * the literal INPUTS of a unit test that lives in the reference's own test modules (file:line cited), restated
    in the float layouts of `rb_debug_kat` / `orc_kat` below, and
  * the values that test ASSERTS ("asserted": literals or properties written in the Rust source), plus, where the
    Rust only asserts a relation between two of its own code paths, the values obtained by evaluating the cited
    Rust function by hand in f32 with numpy, WITHOUT fused multiply-adds, exactly as rustc compiles it
    ("derived": an independent restatement, compared within a tolerance because this engine places explicit FMAs).
tests/test_ref_vectors.py checks the oracle and the emulated kernels (CPU suite) and the CUDA library (-m gpu).

Float layouts (shared by rb_debug_kat and orc_kat):
  pose_drift       in  base t3 q4, cur t3 q4, max_extent                                   out drift
  reduce_manifold  in  n, prediction, local_n1(3), n x (local_p1(3), dist)                 out nsel, sel[4] (-1 = none)
  normal_solve     in  dir3 im1(3) im2(3) torque_dir1(3) torque_dir2(3) ii_torque_dir1(3) ii_torque_dir2(3)
                       r rhs impulse cfm_factor v1(3) w1(3) v2(3) w2(3)                    out impulse v1 w1 v2 w2
  tangent_solve    in  dir3 t1(3) t2(3) im1(3) im2(3) torque_dir1[2](6) torque_dir2[2](6) ii_torque_dir1[2](6)
                       ii_torque_dir2[2](6) r[3] rhsvec(3: rhs_j = rhsvec . t_j) impulse[2] limit v1 w1 v2 w2
                                                                                             out impulse[2] v1 w1 v2 w2
  generate         in  normal3 friction restitution n, n x (anchor1(3) anchor2(3) contact_id impulse warmstart_impulse
                       warmstart_twist warmstart_tangent_world(3) solver_dp1(3) solver_dp2(3)); both sides world-attached
                   out num_contacts dir1(3) tangent1(3) limit impulse[4] impulse_accumulator[4] r[4] builder dist[4]
                       twist_dists[4] tangent impulse[2] tangent accumulator[2] twist impulse, accumulator, twist r,
                       manifold_contact_id[4] (255 = inactive slot)
"""
import json
import os

import numpy as np

F = np.float32
HERE = os.path.dirname(os.path.abspath(__file__))


def f(x):
    return float(F(x))


def vect(x, y):   # contact_with_twist_friction.rs:838-841 (test helper `vect`)
    x, y = F(x), F(y)
    return [f(x), f(y), f(F(0.3) * x - F(0.1) * y)]


def test_manifold(n, seed):
    """contact_with_twist_friction.rs:845-880 `test_manifold(n, seed)`."""
    seed = F(seed)
    pts = []
    for k in range(n):
        kf = F(k)
        pts.append(dict(anchor1=vect(F(0.3) * kf + seed, 0.5), anchor2=vect(F(0.3) * kf + seed, -0.5), contact_id=k,
                        impulse=1.0, warmstart_impulse=f(seed + kf + F(0.25)), warmstart_twist=f(seed * F(0.25) + kf),
                        warmstart_tangent_world=[0.0, 0.0, 0.0],   # the test sets the (unused in 3-D) component pair, the world vector stays default
                        solver_dp1=[0.0, 0.0, 0.0], solver_dp2=[0.0, 0.0, 0.0]))
    return dict(normal=[0.0, 1.0, 0.0], friction=f(0.7), restitution=0.0, points=pts)


def generate_vectors():
    out = []
    for n, seed in ((4, 1.0), (1, 2.0), (3, 0.5)):
        m = test_manifold(n, seed)
        flat = m["normal"] + [m["friction"], m["restitution"], float(n)]
        for p in m["points"]:
            flat += p["anchor1"] + p["anchor2"] + [float(p["contact_id"]), p["impulse"], p["warmstart_impulse"], p["warmstart_twist"]]
            flat += p["warmstart_tangent_world"] + p["solver_dp1"] + p["solver_dp2"]
        # hand evaluation of generate() (:58-424) for a world-attached manifold: poses identity, masses zero
        ws = [p["warmstart_impulse"] for p in m["points"]]
        tw = F(0.0)
        for p in m["points"]:
            tw = F(tw + F(p["warmstart_twist"]) * F(F(1.0) / F(n)))
        exp = dict(num_contacts=n, dir1=[0.0, -1.0, 0.0], tangent1=[0.0, 0.0, 1.0], limit=m["friction"],
                   impulse=ws + [0.0] * (4 - n), impulse_accumulator=[-x for x in ws] + [0.0] * (4 - n),
                   r=[0.0] * 4, dist=[-1.0] * n + [0.0] * (4 - n), twist_dists=[0.0] * 4, tangent_impulse=[0.0, 0.0],
                   twist_impulse=f(tw) if n > 1 else 0.0, manifold_contact_id=list(range(n)) + [255] * (4 - n))
        out.append(dict(name=f"test_manifold({n}, {seed})", function="generate",
                        source="src/dynamics/solver/contact_constraint/contact_with_twist_friction.rs:845-880 (fixture), :901-981 (assertions)",
                        input=[f(x) for x in flat], derived=exp,
                        asserted=dict(inactive_slots=dict(r=0.0, impulse=0.0, twist_dist=0.0, manifold_contact_id=255),
                                      single_point_twist_impulse=0.0, note="Rust asserts: inactive slots hold the neutral fill; single-point lanes have no twist warm-start; num_contacts = lane count"),
                        tol=2e-6))
    return out


def quat_from_scaled_axis(v):   # glam Quat::from_scaled_axis
    v = np.asarray(v, F)
    ang = F(np.sqrt(F(v[0] * v[0] + v[1] * v[1] + v[2] * v[2])))
    if ang == 0:
        return [0.0, 0.0, 0.0, 1.0]
    ax = v / ang
    s, c = F(np.sin(np.float64(ang) * 0.5)), F(np.cos(np.float64(ang) * 0.5))
    return [f(ax[0] * s), f(ax[1] * s), f(ax[2] * s), f(c)]


def pose_drift_vectors():
    """src/geometry/contact_pair.rs:869-900 `same_pose_never_drifts`."""
    axis = np.array([0.3, -0.7, 0.15], F)
    axis = axis / F(np.sqrt(F((axis * axis).sum())))
    cases = []
    for i in range(40):
        for j in range(10):
            angle = F(i) * F(0.157)
            t = [f(F(j) * F(3.7)), 1.0, -2.0]
            q = quat_from_scaled_axis(axis * angle)
            cases.append(t + q + t + q + [4.0])
    return dict(name="same_pose_never_drifts", function="pose_drift", source="src/geometry/contact_pair.rs:869-900",
                inputs=cases, asserted=dict(max_drift=f(F(8.0) * np.finfo(F).eps * F(4.0)),
                                            note="a pose compared against itself reports no drift beyond 8 eps max_extent"))


def ang(x):   # contact_constraint_element.rs:771-779
    x = F(x)
    return [f(x), f(F(0.3) * x), f(F(-0.7) * x)]


VELS = [0.3, -1.2, 0.5, 0.7, 0.1, -0.3, -0.4, 0.9, -0.6, -0.2, 0.4, 0.6]   # contact_constraint_element.rs:806-824 `vels()`
DIR_IM = ([0.0, 1.0, 0.0], [0.5, 0.5, 0.5], [0.25, 0.25, 0.25])           # :826-842 `dir_im()`


def dot_nofma(a, b):
    a, b = np.asarray(a, F), np.asarray(b, F)
    return F(F(F(a[0] * b[0]) + F(a[1] * b[1])) + F(a[2] * b[2]))


def normal_solve_ref(dir1, im1, im2, td1, td2, itd1, itd2, r, rhs, imp, cfm, v1, w1, v2, w2):
    """contact_constraint_element.rs:481-504, f32, no FMA (as rustc compiles glam's scalar path)."""
    dir1, im1, im2, td1, td2, itd1, itd2, v1, w1, v2, w2 = [np.asarray(x, F) for x in (dir1, im1, im2, td1, td2, itd1, itd2, v1, w1, v2, w2)]
    r, rhs, imp, cfm = F(r), F(rhs), F(imp), F(cfm)
    dvel = F(F(F(F(dot_nofma(dir1, v1) + dot_nofma(td1, w1)) - dot_nofma(dir1, v2)) + dot_nofma(td2, w2)) + rhs)
    new = F(cfm * max(F(imp - F(r * dvel)), F(0.0)))
    dl = F(new - imp)
    v1 = v1 + (dir1 * im1) * dl
    w1 = w1 + itd1 * dl
    v2 = v2 + (dir1 * im2) * (-dl)
    w2 = w2 + itd2 * dl
    return [f(new)] + [f(x) for x in np.concatenate([v1, w1, v2, w2]).astype(F)]


def normal_solve_vectors():
    """The scalar `solve` calls of `degraded_solve_pair_matches_scalar_solve` (:844-898): the Rust asserts that its
    2x2 block path equals this scalar path; the scalar path is what this engine implements."""
    out = []
    dir1, im1, im2 = DIR_IM
    for (r, rhs, imp, what) in ((0.8, -2.0, 0.5, "pushes: unclamped branch"), (0.8, 5.0, 0.1, "separates: clamps to zero"),
                                (0.0, -1.0, 0.0, "massless point: inert"), (1.5, -0.3, 2.0, "warm-started, mild correction")):
        for cfm in (1.0, 0.7):
            td1, td2 = ang(0.9), ang(-0.4)
            flat = dir1 + im1 + im2 + td1 + td2 + td1 + td2 + [f(r), f(rhs), f(imp), f(cfm)] + VELS
            exp = normal_solve_ref(dir1, im1, im2, td1, td2, td1, td2, r, rhs, imp, cfm, VELS[0:3], VELS[3:6], VELS[6:9], VELS[9:12])
            a = {}
            if what.startswith("separates"):
                a = dict(impulse=0.0)
            if what.startswith("massless"):
                a = dict(impulse=0.0, velocities_unchanged=True)
            out.append(dict(name=f"scalar solve r={r} rhs={rhs} impulse={imp} cfm={cfm} ({what})", function="normal_solve",
                            source="src/dynamics/solver/contact_constraint/contact_constraint_element.rs:481-504 (function), :844-898 (literal inputs)",
                            input=[f(x) for x in flat], derived=exp, asserted=a, tol=2e-6))
    return out


def tangent_solve_vectors():
    """contact_constraint_element.rs:903-938 `inactive_tangent_slot_is_finite_noop`: impulse 0, limit 0, r = [1, 1, 0],
    rhs = [3, -2], tangents (1,0,0), (0,0,1); must stay a finite no-op."""
    _, im1, im2 = DIR_IM
    t1, t2 = [1.0, 0.0, 0.0], [0.0, 0.0, 1.0]
    dirv = [0.0, -1.0, 0.0]    # dir x t1 = t2
    z6 = [0.0] * 6
    flat = dirv + t1 + t2 + im1 + im2 + z6 + z6 + z6 + z6 + [1.0, 1.0, 0.0] + [3.0, 0.0, -2.0] + [0.0, 0.0] + [0.0] + VELS
    return [dict(name="inactive_tangent_slot_is_finite_noop", function="tangent_solve",
                 source="src/dynamics/solver/contact_constraint/contact_constraint_element.rs:903-938",
                 input=[f(x) for x in flat], asserted=dict(impulse=[0.0, 0.0], velocities=VELS, finite=True))]


def reduce_ref(pts, n1, prediction):
    """manifold_reduction.rs:4-84 restated line by line (f32, no FMA).  pts: list of (p1 xyz, dist)."""
    n = len(pts)
    if n <= 4:
        return n, list(range(n))
    MAXF = np.finfo(F).max
    sel = [-1] * 4
    deepest = MAXF
    for i, p in enumerate(pts):
        if F(p[3]) < deepest:
            deepest = F(p[3]); sel[0] = i
    if sel[0] < 0:
        return 0, []
    a = np.asarray(pts[sel[0]][:3], F)
    furthest = -MAXF
    for i, p in enumerate(pts):
        d = np.asarray(p[:3], F) - a
        dist = dot_nofma(d, d)
        if i != sel[0] and F(p[3]) <= F(prediction) and dist > furthest:
            furthest = dist; sel[1] = i
    if sel[1] < 0:
        return 1, sel[:1]
    b = np.asarray(pts[sel[1]][:3], F)
    if (a == b).all():
        return 1, sel[:1]
    ab = b - a
    n1 = np.asarray(n1, F)
    tangent = np.array([F(F(ab[1] * n1[2]) - F(ab[2] * n1[1])), F(F(ab[2] * n1[0]) - F(ab[0] * n1[2])), F(F(ab[0] * n1[1]) - F(ab[1] * n1[0]))], F)
    mn, mx = MAXF, -MAXF
    for i, p in enumerate(pts):
        if i == sel[0] or i == sel[1] or F(p[3]) > F(prediction):
            continue
        d = dot_nofma(np.asarray(p[:3], F) - a, tangent)
        if d < mn:
            mn = d; sel[2] = i
        if d > mx:
            mx = d; sel[3] = i
    if sel[2] < 0:
        return 2, sel[:2]
    if sel[2] == sel[3]:
        return 3, sel[:3]
    return 4, sel


def reduce_vectors():
    rng = np.random.default_rng(7)
    cases = []
    # an octagon-ish clipped face (what two overlapping quads give), all within prediction
    oct_pts = [(1.0, 0.0, 0.5, -0.01), (0.7, 0.0, 0.9, -0.02), (0.0, 0.0, 1.0, -0.005), (-0.7, 0.0, 0.8, -0.03), (-1.0, 0.0, 0.0, -0.01),
               (-0.6, 0.0, -0.7, 0.0), (0.1, 0.0, -1.0, 0.01), (0.8, 0.0, -0.6, -0.015)]
    cases.append((oct_pts, (0.0, 1.0, 0.0), 0.02))
    cases.append(([(0, 0, 0, 0.5), (1, 0, 0, 0.6), (0, 0, 1, 0.7), (1, 0, 1, 0.8), (2, 0, 2, 0.9)], (0.0, 1.0, 0.0), 0.02))   # only the deepest qualifies
    cases.append(([(0, 0, 0, -0.1), (0, 0, 0, 0.0), (0, 0, 0, 0.01), (0, 0, 0, 0.0), (0, 0, 0, 0.0)], (0.0, 1.0, 0.0), 0.02))   # coincident points
    cases.append(([(0, 0, 0, -0.1), (1, 0, 0, 0.0), (2, 0, 0, 0.01), (3, 0, 0, 0.0), (4, 0, 0, 0.0)], (0.0, 1.0, 0.0), 0.02))   # collinear
    cases.append((oct_pts[:4], (0.0, 1.0, 0.0), 0.02))   # n <= 4: untouched
    for _ in range(40):
        n = int(rng.integers(5, 9))
        pts = [tuple(float(F(x)) for x in (rng.uniform(-1, 1), rng.uniform(-0.01, 0.01), rng.uniform(-1, 1), rng.uniform(-0.05, 0.04))) for _ in range(n)]
        nn = np.array([rng.uniform(-0.2, 0.2), 1.0, rng.uniform(-0.2, 0.2)], F)
        nn = nn / F(np.sqrt(F((nn * nn).sum())))
        cases.append((pts, tuple(float(x) for x in nn), 0.02))
    out = []
    for pts, n1, pred in cases:
        nsel, sel = reduce_ref(pts, n1, pred)
        flat = [float(len(pts)), f(pred)] + [f(x) for x in n1]
        for p in pts:
            flat += [f(x) for x in p]
        out.append(dict(function="reduce_manifold", source="src/geometry/manifold_reduction.rs:4-84", input=flat,
                        derived=dict(num_selected=nsel, selected=list(sel) + [-1] * (4 - len(sel)))))
    return out


def main():
    doc = {"note": __doc__.split("\n\n")[1], "generate": generate_vectors(), "pose_drift": pose_drift_vectors(),
           "normal_solve": normal_solve_vectors(), "tangent_solve": tangent_solve_vectors(), "reduce_manifold": reduce_vectors()}
    with open(os.path.join(HERE, "ref_vectors.json"), "w") as fh:
        json.dump(doc, fh, indent=1)
    print({k: (len(v) if isinstance(v, list) else len(v["inputs"])) for k, v in doc.items() if k != "note"})


if __name__ == "__main__":
    main()
