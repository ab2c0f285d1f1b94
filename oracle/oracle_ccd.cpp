// TEST INFRASTRUCTURE ONLY -- CPU oracle: continuous collision detection (motion clamping), see oracle.h header note.
//
// Follows src/dynamics/ccd/ccd_solver.rs (update_ccd_active_flags :49-92, solve_continuous pass 1 :162-238,
// apply_clamps :325-340), src/dynamics/ccd/sweeps.rs (sweep_fast_body :503-626, cast_collider_pair :282-352,
// cast_sub_shape :354-447), the fast-body criterion of rigid_body_components.rs:1101-1156 and the call site
// src/pipeline/physics_pipeline/substep.rs:492-520.
//
// PARITY UNPINNED for the time of impact: it lives in the reference's parry fork (parry::query::sweep_toi -- Sweep,
// ToiProxy, sweep_time_of_impact, a Box2D-v3-style b2TimeOfImpact over point-cloud proxies), which is not under
// /root/reference.  Restated from the published algorithm's contract (target = max(slop, r1 + r2 - slop), tolerance
// 0.25 slop, pairs that start within the target report fraction 0 and are ignored) as conservative advancement on a
// separating-axis lower bound of the core distance.  max_ccd_substeps > 1 is not restated.
#include <algorithm>
#include "oracle_internal.h"

namespace orc {

static const int CCD_MAX_ITERS = 32;

float ccd_atan01(float z) {   // atan on [0, 1], explicit arithmetic (the kernels evaluate the same polynomial)
    const float s = z * z;
    float p = -0.0117212f;
    p = fma_(p, s, 0.05265332f);
    p = fma_(p, s, -0.11643287f);
    p = fma_(p, s, 0.19354346f);
    p = fma_(p, s, -0.33262347f);
    p = fma_(p, s, 0.99997726f);
    return p * z;
}
float ccd_quat_angle(float vlen, float w) {
    const float aw = w < 0.0f ? -w : w;
    if (vlen == 0.0f) return 0.0f;
    const float half = vlen <= aw ? ccd_atan01(vlen / aw) : 1.5707964f - ccd_atan01(aw / vlen);
    return half * 2.0f;
}

struct Sweep { V3 c1, c2; Q4 q1, q2; V3 lc; };   // parry Sweep::from_poses (Box2D b2Sweep)
static inline Sweep sweep_from_poses(const Pose& a, const Pose& b, V3 lc) {
    Sweep s;
    s.c1 = pose_point(a, lc); s.c2 = pose_point(b, lc); s.q1 = a.q; s.q2 = b.q; s.lc = lc;
    if (qdot(a.q, b.q) < 0.0f) s.q2 = Q4{-b.q.x, -b.q.y, -b.q.z, -b.q.w};
    return s;
}
static inline Pose sweep_transform_at(const Sweep& s, float beta) {   // Sweep::transform_at
    const float om = 1.0f - beta;
    Q4 q{fma_(s.q2.x, beta, s.q1.x * om), fma_(s.q2.y, beta, s.q1.y * om), fma_(s.q2.z, beta, s.q1.z * om), fma_(s.q2.w, beta, s.q1.w * om)};
    q = qnormalize(q);
    const V3 c = madd(s.c1 * om, s.c2, beta);
    return Pose{q, c - qrot(q, s.lc)};
}

static inline float fabs1(float x) { return x < 0.0f ? -x : x; }

static bool core_distance(int shA, V3 heA, const Pose& pA, int shB, V3 heB, const Pose& pB, float& d, V3& n) {
    if (shA == RB_SHAPE_BALL && shB == RB_SHAPE_BALL) {
        const V3 dl = pB.t - pA.t;
        d = length(dl);
        if (!(d > 0.0f)) return false;
        n = dl * (1.0f / d);
        return true;
    }
    if (shA == RB_SHAPE_BALL || shB == RB_SHAPE_BALL) {
        const bool box_is_a = shB == RB_SHAPE_BALL;
        const Pose& pbox = box_is_a ? pA : pB;
        const V3 he = box_is_a ? heA : heB;
        const V3 p = pose_inv_point(pbox, box_is_a ? pB.t : pA.t);
        const V3 q = V3{fclamp(p.x, -he.x, he.x), fclamp(p.y, -he.y, he.y), fclamp(p.z, -he.z, he.z)};
        const V3 dl = p - q;
        d = length(dl);
        if (!(d > 0.0f)) return false;
        const V3 nw = qrot(pbox.q, dl * (1.0f / d));
        n = box_is_a ? nw : -nw;
        return true;
    }
    const M3 ra = qto_mat(pA.q), rb = qto_mat(pB.q);
    const V3 ax[3] = {ra.c0, ra.c1, ra.c2}, bx[3] = {rb.c0, rb.c1, rb.c2};
    const V3 dc = pB.t - pA.t;
    float best = -3.4028235e38f;
    V3 bn = V3{0.0f, 1.0f, 0.0f};
    auto test = [&](V3 a) {
        float s = dot(a, dc);
        if (s < 0.0f) { a = -a; s = -s; }
        const float ea = fma_(heA.z, fabs1(dot(a, ax[2])), fma_(heA.y, fabs1(dot(a, ax[1])), heA.x * fabs1(dot(a, ax[0]))));
        const float eb = fma_(heB.z, fabs1(dot(a, bx[2])), fma_(heB.y, fabs1(dot(a, bx[1])), heB.x * fabs1(dot(a, bx[0]))));
        const float sep = s - ea - eb;
        if (sep > best) { best = sep; bn = a; }
    };
    for (int i = 0; i < 3; ++i) test(ax[i]);
    for (int i = 0; i < 3; ++i) test(bx[i]);
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) {
            const V3 c = cross(ax[i], bx[j]);
            const float l2 = length_sq(c);
            if (l2 > 1.0e-6f) test(c * (1.0f / sqrtf(l2)));
        }
    d = best;
    n = bn;
    return best > 0.0f;
}

// sweep_time_of_impact (restated): fraction at which the moving shape B comes within the target distance of the fixed
// shape A; -1 = not before the end of the sweep; 0 = within it from the start.
static float toi(int shA, V3 heA, const Pose& pA, int shB, V3 heB, const Sweep& sw, float slop) {
    const float total = (shA == RB_SHAPE_BALL ? heA.x : 0.0f) + (shB == RB_SHAPE_BALL ? heB.x : 0.0f);
    const float target = fmax2(slop, total - slop), tol = 0.25f * slop;
    const V3 vc = sw.c2 - sw.c1;
    const Q4 dq = qmul(sw.q2, qconj(sw.q1));
    float x = length(V3{dq.x, dq.y, dq.z});
    if (x > 1.0f) x = 1.0f;
    const float wrate = 4.0f * x / (1.0f + sqrtf(fmax2(1.0f - x * x, 0.0f)));
    const V3 al = V3{fabs1(sw.lc.x), fabs1(sw.lc.y), fabs1(sw.lc.z)};
    const float rmax = shB == RB_SHAPE_BALL ? length(al) : length(heB + al);
    float t = 0.0f;
    for (int it = 0; it < CCD_MAX_ITERS; ++it) {
        const Pose pB = sweep_transform_at(sw, t);
        float d;
        V3 n;
        if (!core_distance(shA, heA, pA, shB, heB, pB, d, n)) return t == 0.0f ? 0.0f : t;
        if (d < target + tol) return t;
        const float bound = fmax2(-dot(vc, n), 0.0f) + wrate * rmax;
        if (!(bound > 0.0f)) return -1.0f;
        t = t + (d - target) / bound;
        if (!(t < 1.0f)) return -1.0f;
    }
    return t;
}

static bool is_moving_fast_with_next_position(const World& w, const Body& b) {   // rigid_body_components.rs:1125-1156
    const RbIntegrationParameters& P = w.params.p;
    const float inv_dt = P.dt == 0.0f ? 0.0f : 1.0f / P.dt;
    const V3 dcom = pose_point(b.next_pos, b.local_com) - pose_point(b.pos, b.local_com);
    const Q4 dq = qmul(b.next_pos.q, qconj(b.pos.q));
    const float x = length(V3{dq.x, dq.y, dq.z});
    const float max_delta_position = length(dcom) + 2.0f * x * b.max_extent;
    const float max_velocity = length(dcom * inv_dt) + (ccd_quat_angle(x, dq.w) * inv_dt) * b.max_extent;   // ccd_vels = interpolate_velocity (:147-196)
    return fmax2(max_delta_position, max_velocity * P.dt) > 0.5f * b.ccd_thickness;
}

// substep.rs:492-520 with max_ccd_substeps = 1: flags, then CCDSolver::solve_continuous pass 1 + apply_clamps.
void ccd_motion_clamping(World& w) {
    const RbIntegrationParameters& P = w.params.p;
    if (P.max_ccd_substeps == 0) return;
    const float slop = P.normalized_allowed_linear_error * P.length_unit;
  for (int pass = 0; pass < 2; ++pass) {   // pass 0: non-bullets vs fixed targets; pass 1: bullets vs everything but bullets (ccd_solver.rs:190-263)
    for (int bi = 0; bi < (int)w.bodies.size(); ++bi) {
        Body& b = w.bodies[bi];
        if (!b.is_awake() || !b.is_strict_dynamic()) continue;
        const V3& nt = b.next_pos.t;
        if (!(std::isfinite(nt.x) && std::isfinite(nt.y) && std::isfinite(nt.z))) continue;   // (left to the quarantine chokepoint)
        if (!is_moving_fast_with_next_position(w, b)) continue;
        const bool bullet = (b.flags & RB_BODY_CCD_ENABLED) != 0;   // is_bullet (sweeps.rs:31-33)
        if (bullet != (pass == 1)) continue;
        float frac = 1.0f;
        for (int ci = 0; ci < (int)w.colliders.size(); ++ci) {
            const Collider& c1 = w.colliders[ci];
            if (c1.parent != bi || c1.shape < 0 || c1.shape >= RB_SHAPE_CAPSULE || c1.sensor) continue;   // (capsules and polyhedra are not swept; sensors never: ccd_solver.rs:92-93)
            const Pose cs = pose_mul(b.pos, c1.pos_wrt_parent), ce = pose_mul(b.next_pos, c1.pos_wrt_parent);
            const Sweep sw = sweep_from_poses(cs, ce, pose_inv_point(c1.pos_wrt_parent, b.local_com));
            const Aabb a1 = shape_aabb(c1.shape, c1.he, cs), a2 = shape_aabb(c1.shape, c1.he, ce);
            Aabb swept;
            swept.mins = V3{fmin2(a1.mins.x, a2.mins.x), fmin2(a1.mins.y, a2.mins.y), fmin2(a1.mins.z, a2.mins.z)};
            swept.maxs = V3{fmax2(a1.maxs.x, a2.maxs.x), fmax2(a1.maxs.y, a2.maxs.y), fmax2(a1.maxs.z, a2.maxs.z)};
            for (const Collider& c2 : w.colliders) {
                if (c2.shape < 0 || c2.shape >= RB_SHAPE_CAPSULE || c2.sensor) continue;
                Pose target_pose = c2.pos;
                if (c2.parent >= 0 && w.bodies[c2.parent].type != RB_BODY_FIXED) {   // tier_allows (sweeps.rs:36-42)
                    const Body& t = w.bodies[c2.parent];
                    if (!bullet || c2.parent == bi || !t.is_dynamic() || (t.is_strict_dynamic() && (t.flags & RB_BODY_CCD_ENABLED))) continue;
                    target_pose = pose_mul(t.is_awake() ? t.next_pos : t.pos, c2.pos_wrt_parent);   // target_collider_pose (:101-109)
                }
                const Aabb& f = c2.fat;   // (any superset of the colliders within the prediction distance gives the same minimum)
                if (!(swept.mins.x <= f.maxs.x && swept.mins.y <= f.maxs.y && swept.mins.z <= f.maxs.z && swept.maxs.x >= f.mins.x &&
                      swept.maxs.y >= f.mins.y && swept.maxs.z >= f.mins.z)) continue;
                if (!((c1.memberships & c2.filter) != 0 && (c2.memberships & c1.filter) != 0)) continue;
                const float fr = toi(c2.shape, c2.he, target_pose, c1.shape, c1.he, sw, slop);
                if (fr > 0.0f && fr < frac) frac = fr;
            }
        }
        if (frac < 1.0f) b.next_pos = sweep_transform_at(sweep_from_poses(b.pos, b.next_pos, b.local_com), frac);
    }
  }
}

}  // namespace orc
