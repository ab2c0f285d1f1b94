// TEST INFRASTRUCTURE ONLY -- CPU oracle: the velocity solver + integrator (see oracle.h header note).
//
// Follows src/pipeline/physics_pipeline/solve.rs:159-401 (build_islands_and_solve_velocity_constraints)
// -> src/dynamics/solver/staged_island_solver/{init.rs:30-544, worker.rs:32-898, solve.rs:12-209}
// with the constraint arithmetic of src/dynamics/solver/contact_constraint/
// {contact_with_twist_friction.rs, contact_constraint_element.rs} and the joint rows of
// src/dynamics/solver/joint_constraint/{joint_constraint_builder.rs, joint_constraint_helper.rs,
// joint_velocity_constraint.rs}.  One constraint at a time, colours in the reference's stage order.
#include <algorithm>
#include <cstring>
#include <string>
#include "oracle_internal.h"
#include "opool.h"

namespace orc {

void set_threads(int n) { Pool::get().set_threads(n); }
int get_threads() { return Pool::get().threads(); }

// integration_parameters.rs:85-149
struct Spring {
    float natural_frequency, damping_ratio;
    float angular_frequency() const { return natural_frequency * 6.283185307179586f; }
    float erp_inv_dt(float dt) const {
        float w = angular_frequency();
        return w / (dt * w + 2.0f * damping_ratio);
    }
    float erp(float dt) const { return dt * erp_inv_dt(dt); }
    float cfm_coeff(float dt) const {
        float e = erp(dt);
        if (e == 0.0f) return 0.0f;
        float inv_erp_minus_one = 1.0f / e - 1.0f;
        return inv_erp_minus_one * inv_erp_minus_one /
               ((1.0f + inv_erp_minus_one) * 4.0f * damping_ratio * damping_ratio);
    }
    float cfm_factor(float dt) const { return 1.0f / (1.0f + cfm_coeff(dt)); }
};

struct GatheredBody {  // solver_body.rs:11-33,338-350: world-attached side = identity / zero
    V3 lin, ang;
    Pose pose;
    Sdp3 ii;
    V3 im;
};
static inline GatheredBody gather(const World& w, uint32_t id) {
    GatheredBody g;
    if (id == NO_BODY) {
        g.lin = vzero(); g.ang = vzero(); g.pose = pose_identity(); g.ii = sdp_zero(); g.im = vzero();
    } else {
        const SolverBody& s = w.sb[id];
        g.lin = s.lin; g.ang = s.ang; g.pose = s.pose; g.ii = s.ii; g.im = s.im;
    }
    return g;
}
static inline void scatter_vel(World& w, uint32_t id, V3 lin, V3 ang) {
    if (id == NO_BODY) return;
    w.sb[id].lin = lin;
    w.sb[id].ang = ang;
}

static inline float is_bouncy(float restitution, bool is_new) {  // contact_pair.rs:773-779
    return is_new ? (restitution > 0.0f ? 1.0f : 0.0f) : (restitution >= 1.0f ? 1.0f : 0.0f);
}

// contact_with_twist_friction.rs:58-424 (one lane)
static void generate(World& w, int pair_index, Constraint& c) {
    const Pair& p = w.pairs[pair_index];
    memset(&c, 0, sizeof(c));
    c.pair = pair_index;
    c.coulomb = w.params.p.friction_model == 1;   // FrictionModel::Coulomb (init.rs:419)
    const int rel_dom = relative_dominance(p.b1 >= 0 ? &w.bodies[p.b1] : nullptr, p.b2 >= 0 ? &w.bodies[p.b2] : nullptr);
    bool d1 = p.b1 >= 0 && w.bodies[p.b1].is_awake() && rel_dom <= 0;   // contact_with_twist_friction.rs:71-84
    bool d2 = p.b2 >= 0 && w.bodies[p.b2].is_awake() && rel_dom >= 0;
    c.id1 = d1 ? (uint32_t)p.b1 : NO_BODY;
    c.id2 = d2 ? (uint32_t)p.b2 : NO_BODY;
    GatheredBody g1 = gather(w, c.id1), g2 = gather(w, c.id2);
    V3 world_com1 = g1.pose.t, world_com2 = g2.pose.t;
    V3 force_dir1 = -p.normal;
    int count = p.nsc < MAX_MANIFOLD_POINTS ? p.nsc : MAX_MANIFOLD_POINTS;
    V3 t1 = orthonormal_vector(force_dir1);   // contact_constraint/mod.rs:26-47
    V3 t2 = cross(force_dir1, t1);
    float inv_num_points = 1.0f / (float)count;
    c.dir1 = force_dir1;
    c.im1 = g1.im; c.im2 = g2.im; c.ii1 = g1.ii; c.ii2 = g2.ii;
    c.b_restitution = p.restitution;
    c.num_contacts = count;
    c.tangent1 = t1;
    c.limit = p.friction;

    V3 friction_center = vzero(), friction_center2 = vzero();
    float twist_warmstart = 0.0f;
    float tangent_warmstart[2] = {0.0f, 0.0f};
    V3 tangent_vel = vzero();
    V3 points[MAX_MANIFOLD_POINTS];
    for (int k = 0; k < count; ++k) {
        float weight = inv_num_points;
        const SolverContact& sc = p.sc[k];
        const Point& pt = p.pts[sc.cid];
        float warmstart_impulse = pt.warmstart_impulse;
        V3 wt = pt.warmstart_tangent_world;
        float ws_t0 = dot(wt, t1), ws_t1 = dot(wt, t2);
        float warmstart_twist = pt.warmstart_twist;
        bool is_new = pt.impulse == 0.0f;
        float bouncy = is_bouncy(p.restitution, is_new);
        V3 p1 = pose_point(g1.pose, sc.anchor1);
        V3 p2 = pose_point(g2.pose, sc.anchor2);
        float dist = dot(p1 - p2, force_dir1);
        V3 dp1 = pt.dp1, dp2 = pt.dp2;
        V3 point = world_com1 + dp1;
        points[k] = point;
        friction_center = madd(friction_center, point, weight);
        friction_center2 = madd(friction_center2, world_com2 + dp2, weight);
        V3 vel1 = g1.lin + cross(g1.ang, dp1);
        V3 vel2 = g2.lin + cross(g2.ang, dp2);
        twist_warmstart = fma_(warmstart_twist, weight, twist_warmstart);
        tangent_warmstart[0] = fma_(ws_t0, weight, tangent_warmstart[0]);
        tangent_warmstart[1] = fma_(ws_t1, weight, tangent_warmstart[1]);
        tangent_vel = tangent_vel + vzero() * weight;
        c.cids[k] = sc.cid;
        NormalPart& n = c.normal[k];
        n.torque_dir1 = cross(dp1, force_dir1);
        n.torque_dir2 = cross(dp2, -force_dir1);
        n.ii_torque_dir1 = sdp_mul(g1.ii, n.torque_dir1);
        n.ii_torque_dir2 = sdp_mul(g2.ii, n.torque_dir2);
        V3 imsum = g1.im + g2.im;
        float projected_mass = inv_or_zero(dot(force_dir1, cmul(imsum, force_dir1)) + dot(n.ii_torque_dir1, n.torque_dir1) +
                                           dot(n.ii_torque_dir2, n.torque_dir2));
        float projected_velocity = dot(vel1 - vel2, force_dir1);
        float restitution_seed = bouncy * p.restitution * projected_velocity;
        n.impulse = warmstart_impulse;
        n.impulse_accumulator = -n.impulse;
        n.r = projected_mass;
        c.b_local_p1[k] = pose_inv_point(g1.pose, point);
        c.b_local_p2[k] = pose_inv_point(g2.pose, world_com2 + dp2);
        c.b_dist[k] = dist - dot(point - (world_com2 + dp2), force_dir1);
        c.b_restitution_seed[k] = restitution_seed;
        if (c.coulomb) {   // contact_with_coulomb_friction.rs:255-302: one tangent part per point, arms = the point's own
            TangentPart& tp = c.tangent[k];
            tp.impulse[0] = ws_t0; tp.impulse[1] = ws_t1;
            tp.impulse_acc[0] = -ws_t0; tp.impulse_acc[1] = -ws_t1;
            const V3 tj[2] = {t1, t2};
            for (int j = 0; j < 2; ++j) {
                tp.torque_dir1[j] = cross(dp1, tj[j]);
                tp.torque_dir2[j] = cross(dp2, -tj[j]);
                tp.ii_torque_dir1[j] = sdp_mul(g1.ii, tp.torque_dir1[j]);
                tp.ii_torque_dir2[j] = sdp_mul(g2.ii, tp.torque_dir2[j]);
                tp.r[j] = dot(tj[j], cmul(imsum, tj[j])) + dot(tp.ii_torque_dir1[j], tp.torque_dir1[j]) + dot(tp.ii_torque_dir2[j], tp.torque_dir2[j]);
                tp.rhs_wo_bias[j] = 0.0f;   // tangent_velocity . t_j, identically zero without contact-modification hooks
                tp.rhs[j] = 0.0f;
            }
            tp.r[2] = 2.0f * (dot(tp.ii_torque_dir1[0], tp.torque_dir1[1]) + dot(tp.ii_torque_dir2[0], tp.torque_dir2[1]));
        }
    }
    if (c.coulomb) return;   // no friction centre, no twist row (contact_with_coulomb_friction.rs:41-49)
    c.t_impulse[0] = tangent_warmstart[0];
    c.t_impulse[1] = tangent_warmstart[1];
    c.t_impulse_acc[0] = -tangent_warmstart[0];
    c.t_impulse_acc[1] = -tangent_warmstart[1];
    c.w_impulse = count > 1 ? twist_warmstart : 0.0f;
    c.w_impulse_acc = -c.w_impulse;
    c.b_lfc1 = pose_inv_point(g1.pose, friction_center);
    c.b_lfc2 = pose_inv_point(g2.pose, friction_center2);
    c.b_tangent_vel = tangent_vel;
    V3 dp1 = friction_center - world_com1;
    V3 dp2 = friction_center2 - world_com2;
    if (count > 1) {
        for (int k = 0; k < count; ++k) c.twist_dists[k] = length(friction_center - points[k]);
        V3 ii_twist_dir1 = sdp_mul(g1.ii, force_dir1);
        V3 ii_twist_dir2 = sdp_mul(g2.ii, -force_dir1);
        c.w_rhs = 0.0f;
        c.w_r = inv_or_zero(dot(ii_twist_dir1, force_dir1) + dot(ii_twist_dir2, -force_dir1));
    }
    c.t_dp1 = dp1;
    c.t_dp2 = dp2;
    V3 tangents[2] = {t1, t2};
    for (int j = 0; j < 2; ++j) {
        V3 td1 = cross(dp1, tangents[j]);
        V3 td2 = cross(dp2, -tangents[j]);
        V3 itd1 = sdp_mul(g1.ii, td1);
        V3 itd2 = sdp_mul(g2.ii, td2);
        V3 imsum = g1.im + g2.im;
        float r = dot(tangents[j], cmul(imsum, tangents[j])) + dot(itd1, td1) + dot(itd2, td2);
        float rhs_wo_bias = 0.0f;  // = tangent_vel . t with tangent_velocity identically zero (no hooks)
        c.t_torque_dir1[j] = td1; c.t_torque_dir2[j] = td2;
        c.t_ii_torque_dir1[j] = itd1; c.t_ii_torque_dir2[j] = itd2;
        c.t_rhs_wo_bias[j] = rhs_wo_bias;
        c.t_rhs[j] = rhs_wo_bias;
        c.t_r[j] = r;
    }
    c.t_r[2] = 2.0f * (dot(c.t_ii_torque_dir1[0], c.t_torque_dir1[1]) + dot(c.t_ii_torque_dir2[0], c.t_torque_dir2[1]));
}

struct SubParams {
    float dt, inv_dt;
    float dyn_cfm, static_cfm, dyn_erp, static_erp;
    float max_corrective_velocity, warmstart_coeff;
};

// contact_with_twist_friction.rs:426-522
static void update(const World& w, Constraint& c, const SubParams& sp, float solved_dt) {
    float is_static = (c.id1 == NO_BODY || c.id2 == NO_BODY) ? 1.0f : 0.0f;
    float cfm_factor = sp.dyn_cfm + is_static * (sp.static_cfm - sp.dyn_cfm);
    float erp_inv_dt = sp.dyn_erp + is_static * (sp.static_erp - sp.dyn_erp);
    GatheredBody g1 = gather(w, c.id1), g2 = gather(w, c.id2);
    V3 tangents[2] = {c.tangent1, cross(c.dir1, c.tangent1)};
    (void)solved_dt;  // tangent_velocity is identically zero without contact-modification hooks (out of scope)
    for (int k = 0; k < c.num_contacts; ++k) {
        NormalPart& n = c.normal[k];
        V3 p1 = pose_point(g1.pose, c.b_local_p1[k]);
        V3 p2 = pose_point(g2.pose, c.b_local_p2[k]);
        float dist = c.b_dist[k] + dot(p1 - p2, c.dir1);
        float rhs_wo_bias = fmax2(dist, 0.0f) * sp.inv_dt;
        float rhs_bias = fclamp(dist * erp_inv_dt, -sp.max_corrective_velocity, 0.0f);
        n.rhs_wo_bias = rhs_wo_bias;
        n.rhs = rhs_wo_bias + rhs_bias;
        n.cfm_factor = dist <= 0.0f ? cfm_factor : 1.0f;
        n.impulse_accumulator = n.impulse_accumulator + n.impulse;
        n.impulse = n.impulse * sp.warmstart_coeff;
        if (c.coulomb) {   // contact_with_coulomb_friction.rs:438-447
            TangentPart& tp = c.tangent[k];
            for (int j = 0; j < 2; ++j) {
                tp.impulse_acc[j] = tp.impulse_acc[j] + tp.impulse[j];
                tp.impulse[j] = tp.impulse[j] * sp.warmstart_coeff;
                float bias = dot(p1 - p2, tangents[j]) * sp.inv_dt;
                tp.rhs[j] = tp.rhs_wo_bias[j] + bias;
            }
        }
    }
    if (!c.coulomb) {
        V3 p1 = pose_point(g1.pose, c.b_lfc1);
        V3 p2 = pose_point(g2.pose, c.b_lfc2);
        for (int j = 0; j < 2; ++j) {
            float bias = dot(p1 - p2, tangents[j]) * sp.inv_dt;
            c.t_rhs[j] = c.t_rhs_wo_bias[j] + bias;
        }
        c.t_impulse_acc[0] = c.t_impulse_acc[0] + c.t_impulse[0];
        c.t_impulse_acc[1] = c.t_impulse_acc[1] + c.t_impulse[1];
        c.t_impulse[0] = c.t_impulse[0] * sp.warmstart_coeff;
        c.t_impulse[1] = c.t_impulse[1] * sp.warmstart_coeff;
        c.w_impulse_acc = c.w_impulse_acc + c.w_impulse;
        c.w_impulse = c.w_impulse * sp.warmstart_coeff;
    }
}

// contact_with_twist_friction.rs:529-554
static void refresh_rhs_wo_bias(const World& w, Constraint& c, const SubParams& sp, float solved_dt) {
    GatheredBody g1 = gather(w, c.id1), g2 = gather(w, c.id2);
    (void)solved_dt;  // tangent_velocity is identically zero without contact-modification hooks (out of scope)
    for (int k = 0; k < c.num_contacts; ++k) {
        V3 p1 = pose_point(g1.pose, c.b_local_p1[k]);
        V3 p2 = pose_point(g2.pose, c.b_local_p2[k]);
        float dist = c.b_dist[k] + dot(p1 - p2, c.dir1);
        c.normal[k].rhs = fmax2(dist, 0.0f) * sp.inv_dt;
        c.normal[k].cfm_factor = 1.0f;
        if (c.coulomb) { c.tangent[k].rhs[0] = c.tangent[k].rhs_wo_bias[0]; c.tangent[k].rhs[1] = c.tangent[k].rhs_wo_bias[1]; }
    }
    c.t_rhs[0] = c.t_rhs_wo_bias[0];
    c.t_rhs[1] = c.t_rhs_wo_bias[1];
}

// contact_with_twist_friction.rs:633-678 + contact_constraint_element.rs:465-478,627-647,720-732
static void warmstart(World& w, Constraint& c) {
    GatheredBody g1 = gather(w, c.id1), g2 = gather(w, c.id2);
    V3 v1 = g1.lin, w1 = g1.ang, v2 = g2.lin, w2 = g2.ang;
    for (int k = 0; k < c.num_contacts; ++k) {
        const NormalPart& n = c.normal[k];
        v1 = madd(v1, cmul(c.dir1, c.im1), n.impulse);
        w1 = madd(w1, n.ii_torque_dir1, n.impulse);
        v2 = madd(v2, cmul(c.dir1, c.im2), -n.impulse);
        w2 = madd(w2, n.ii_torque_dir2, n.impulse);
    }
    V3 t0 = c.tangent1, t1 = cross(c.dir1, c.tangent1);
    if (c.coulomb) {   // contact_with_coulomb_friction.rs:584-592 + contact_constraint_element.rs:64-98, point by point
        for (int k = 0; k < c.num_contacts; ++k) {
            const TangentPart& tp = c.tangent[k];
            v1 = maddv(v1, madd(t0 * tp.impulse[0], t1, tp.impulse[1]), c.im1);
            w1 = madd(madd(w1, tp.ii_torque_dir1[0], tp.impulse[0]), tp.ii_torque_dir1[1], tp.impulse[1]);
            v2 = maddv(v2, madd(t0 * (-tp.impulse[0]), t1, -tp.impulse[1]), c.im2);
            w2 = madd(madd(w2, tp.ii_torque_dir2[0], tp.impulse[0]), tp.ii_torque_dir2[1], tp.impulse[1]);
        }
        scatter_vel(w, c.id1, v1, w1);
        scatter_vel(w, c.id2, v2, w2);
        return;
    }
    v1 = maddv(v1, madd(t0 * c.t_impulse[0], t1, c.t_impulse[1]), c.im1);
    w1 = madd(madd(w1, c.t_ii_torque_dir1[0], c.t_impulse[0]), c.t_ii_torque_dir1[1], c.t_impulse[1]);
    v2 = maddv(v2, madd(t0 * (-c.t_impulse[0]), t1, -c.t_impulse[1]), c.im2);
    w2 = madd(madd(w2, c.t_ii_torque_dir2[0], c.t_impulse[0]), c.t_ii_torque_dir2[1], c.t_impulse[1]);
    if (c.num_contacts > 1) {
        V3 ii_twist_dir1 = sdp_mul(c.ii1, c.dir1);
        V3 ii_twist_dir2 = sdp_mul(c.ii2, c.dir1);
        w1 = madd(w1, ii_twist_dir1, c.w_impulse);
        w2 = madd(w2, ii_twist_dir2, -c.w_impulse);
    }
    scatter_vel(w, c.id1, v1, w1);
    scatter_vel(w, c.id2, v2, w2);
}

// contact_with_twist_friction.rs:680-781 + contact_constraint_element.rs:481-504,650-705,735-756
static void solve(World& w, Constraint& c, bool solve_friction) {
    GatheredBody g1 = gather(w, c.id1), g2 = gather(w, c.id2);
    V3 v1 = g1.lin, w1 = g1.ang, v2 = g2.lin, w2 = g2.ang;
    for (int k = 0; k < c.num_contacts; ++k) {
        NormalPart& n = c.normal[k];
        float dvel = dot(c.dir1, v1) + dot(n.torque_dir1, w1) - dot(c.dir1, v2) + dot(n.torque_dir2, w2) + n.rhs;
        float new_impulse = n.cfm_factor * fmax2(fma_(-n.r, dvel, n.impulse), 0.0f);
        float dlambda = new_impulse - n.impulse;
        n.impulse = new_impulse;
        v1 = madd(v1, cmul(c.dir1, c.im1), dlambda);
        w1 = madd(w1, n.ii_torque_dir1, dlambda);
        v2 = madd(v2, cmul(c.dir1, c.im2), -dlambda);
        w2 = madd(w2, n.ii_torque_dir2, dlambda);
    }
    if (solve_friction && c.coulomb) {   // contact_with_coulomb_friction.rs:659-676 + contact_constraint_element.rs:124-176
        V3 t0 = c.tangent1, t1 = cross(c.dir1, c.tangent1);
        for (int k = 0; k < c.num_contacts; ++k) {
            TangentPart& tp = c.tangent[k];
            const float limit = (0.0f + c.normal[k].impulse) * c.limit;
            float dvel_0 = dot(t0, v1) + dot(tp.torque_dir1[0], w1) - dot(t0, v2) + dot(tp.torque_dir2[0], w2) + tp.rhs[0];
            float dvel_1 = dot(t1, v1) + dot(tp.torque_dir1[1], w1) - dot(t1, v2) + dot(tp.torque_dir2[1], w2) + tp.rhs[1];
            float k11 = tp.r[0], k22 = tp.r[1], k12 = tp.r[2] * 0.5f;
            float inv_det = inv_or_zero(fma_(k11, k22, -(k12 * k12)));
            float d0 = fma_(k22, dvel_0, -(k12 * dvel_1)) * inv_det;
            float d1 = fma_(k11, dvel_1, -(k12 * dvel_0)) * inv_det;
            float n0 = tp.impulse[0] - d0, n1 = tp.impulse[1] - d1;
            float len = sqrtf(fma_(n1, n1, n0 * n0));
            if (len > limit) {
                float s = limit / len;
                n0 = n0 * s;
                n1 = n1 * s;
            }
            float dl0 = n0 - tp.impulse[0], dl1 = n1 - tp.impulse[1];
            tp.impulse[0] = n0;
            tp.impulse[1] = n1;
            v1 = maddv(v1, madd(t0 * dl0, t1, dl1), c.im1);
            w1 = madd(madd(w1, tp.ii_torque_dir1[0], dl0), tp.ii_torque_dir1[1], dl1);
            v2 = maddv(v2, madd(t0 * (-dl0), t1, -dl1), c.im2);
            w2 = madd(madd(w2, tp.ii_torque_dir2[0], dl0), tp.ii_torque_dir2[1], dl1);
        }
    } else if (solve_friction) {
        V3 t0 = c.tangent1, t1 = cross(c.dir1, c.tangent1);
        float tangent_limit = 0.0f, twist_limit = 0.0f;
        for (int k = 0; k < c.num_contacts; ++k) {
            tangent_limit = tangent_limit + c.normal[k].impulse;
            twist_limit = fma_(c.normal[k].impulse, c.twist_dists[k], twist_limit);
        }
        tangent_limit = tangent_limit * c.limit;
        twist_limit = twist_limit * c.limit;
        if (c.num_contacts > 1) {
            V3 ii_twist_dir1 = sdp_mul(c.ii1, c.dir1);
            V3 ii_twist_dir2 = sdp_mul(c.ii2, c.dir1);
            float dvel = dot(c.dir1, w1 - w2) + c.w_rhs;
            float new_impulse = fclamp(fma_(-c.w_r, dvel, c.w_impulse), -twist_limit, twist_limit);
            float dlambda = new_impulse - c.w_impulse;
            c.w_impulse = new_impulse;
            w1 = madd(w1, ii_twist_dir1, dlambda);
            w2 = madd(w2, ii_twist_dir2, -dlambda);
        }
        float dvel_0 = dot(t0, v1) + dot(c.t_torque_dir1[0], w1) - dot(t0, v2) + dot(c.t_torque_dir2[0], w2) + c.t_rhs[0];
        float dvel_1 = dot(t1, v1) + dot(c.t_torque_dir1[1], w1) - dot(t1, v2) + dot(c.t_torque_dir2[1], w2) + c.t_rhs[1];
        float k11 = c.t_r[0], k22 = c.t_r[1], k12 = c.t_r[2] * 0.5f;
        float inv_det = inv_or_zero(fma_(k11, k22, -(k12 * k12)));
        float d0 = fma_(k22, dvel_0, -(k12 * dvel_1)) * inv_det;
        float d1 = fma_(k11, dvel_1, -(k12 * dvel_0)) * inv_det;
        float n0 = c.t_impulse[0] - d0, n1 = c.t_impulse[1] - d1;
        // nalgebra simd_cap_magnitude: scale down to `limit` when longer.
        float len = sqrtf(fma_(n1, n1, n0 * n0));
        if (len > tangent_limit) {
            float s = tangent_limit / len;
            n0 = n0 * s;
            n1 = n1 * s;
        }
        float dl0 = n0 - c.t_impulse[0], dl1 = n1 - c.t_impulse[1];
        c.t_impulse[0] = n0;
        c.t_impulse[1] = n1;
        v1 = maddv(v1, madd(t0 * dl0, t1, dl1), c.im1);
        w1 = madd(madd(w1, c.t_ii_torque_dir1[0], dl0), c.t_ii_torque_dir1[1], dl1);
        v2 = maddv(v2, madd(t0 * (-dl0), t1, -dl1), c.im2);
        w2 = madd(madd(w2, c.t_ii_torque_dir2[0], dl0), c.t_ii_torque_dir2[1], dl1);
    }
    scatter_vel(w, c.id1, v1, w1);
    scatter_vel(w, c.id2, v2, w2);
}

// contact_with_twist_friction.rs:568-597 + contact_constraint_element.rs:508-534
static void apply_restitution(World& w, Constraint& c) {
    bool any = false;
    for (int k = 0; k < c.num_contacts; ++k) any = any || c.b_restitution_seed[k] < 0.0f;
    if (!any) return;
    GatheredBody g1 = gather(w, c.id1), g2 = gather(w, c.id2);
    V3 v1 = g1.lin, w1 = g1.ang, v2 = g2.lin, w2 = g2.ang;
    for (int k = 0; k < c.num_contacts; ++k) {
        NormalPart& n = c.normal[k];
        float seed = c.b_restitution_seed[k];
        float dvel = dot(c.dir1, v1) + dot(n.torque_dir1, w1) - dot(c.dir1, v2) + dot(n.torque_dir2, w2) + seed;
        bool gate = seed < 0.0f && (n.impulse_accumulator + n.impulse) > 0.0f;
        float new_impulse = fmax2(fma_(-n.r, dvel, n.impulse), 0.0f);
        if (!gate) new_impulse = n.impulse;
        float dlambda = new_impulse - n.impulse;
        n.impulse = new_impulse;
        v1 = madd(v1, cmul(c.dir1, c.im1), dlambda);
        w1 = madd(w1, n.ii_torque_dir1, dlambda);
        v2 = madd(v2, cmul(c.dir1, c.im2), -dlambda);
        w2 = madd(w2, n.ii_torque_dir2, dlambda);
    }
    scatter_vel(w, c.id1, v1, w1);
    scatter_vel(w, c.id2, v2, w2);
}

static inline float canon0(float x) { return x == 0.0f ? 0.0f : x; }  // utils::canonicalize_zero

// contact_with_twist_friction.rs:783-829
static void writeback_impulses(World& w, const Constraint& c) {
    Pair& p = w.pairs[c.pair];
    V3 t2 = cross(c.dir1, c.tangent1);
    if (c.coulomb) {   // contact_with_coulomb_friction.rs:683-740: per-point world tangent impulse, the twist slot is left alone
        for (int k = 0; k < c.num_contacts; ++k) {
            Point& pt = p.pts[c.cids[k]];
            float a0 = canon0(c.tangent[k].impulse[0]), a1 = canon0(c.tangent[k].impulse[1]);
            V3 tw = c.tangent1 * a0 + t2 * a1;
            pt.warmstart_impulse = canon0(c.normal[k].impulse);
            pt.impulse = canon0(c.normal[k].impulse_accumulator + c.normal[k].impulse);
            pt.warmstart_tangent_world = V3{canon0(tw.x), canon0(tw.y), canon0(tw.z)};
        }
        return;
    }
    float ti0 = canon0(c.t_impulse[0]), ti1 = canon0(c.t_impulse[1]);
    V3 tw = c.tangent1 * ti0 + t2 * ti1;
    tw = V3{canon0(tw.x), canon0(tw.y), canon0(tw.z)};
    float twist = canon0(c.w_impulse);
    for (int k = 0; k < c.num_contacts; ++k) {
        Point& pt = p.pts[c.cids[k]];
        pt.warmstart_impulse = canon0(c.normal[k].impulse);
        pt.impulse = canon0(c.normal[k].impulse_accumulator + c.normal[k].impulse);
        pt.warmstart_tangent_world = tw;
        pt.warmstart_twist = twist;
    }
}

// rigid_body.rs:2023-2046
static V3 gyroscopic_corrected_angvel(V3 angvel, Q4 principal_axes, V3 principal_inertia, V3 inv_principal_inertia, float dt) {
    V3 wl = qrot_inv(principal_axes, angvel);
    V3 curr_momentum = cmul(principal_inertia, wl);
    V3 explicit_gyro = (-cross(wl, curr_momentum)) * dt;
    V3 total = curr_momentum + explicit_gyro;
    float total_sq = length_sq(total);
    if (total_sq != 0.0f) {
        V3 capped = total * sqrtf(length_sq(curr_momentum) / total_sq);
        return qrot(principal_axes, cmul(inv_principal_inertia, capped));
    }
    return angvel;
}

// ------------------------------------------------------------------------------------------
// Joints (Appendix B of SURVEY.md): per-substep row rebuild + Gram-Schmidt + solve.
// ------------------------------------------------------------------------------------------
static int joint_rows_of(const Joint& j) {
    int n = 0;
    const uint32_t free_axes = ~j.locked_axes & 63u;
    const uint32_t coupled = j.coupled_axes, lim = j.limit_axes & free_axes, mot = j.motor_axes & free_axes;
    for (int i = 0; i < 6; ++i) n += ((j.locked_axes >> i) & 1) + (((lim & ~coupled) >> i) & 1) + (((mot & ~coupled) >> i) & 1);
    // coupled axes (joint_velocity_constraint.rs:224-250, :319-355): one motor row for the linear ones, one limit row per kind
    if (mot & coupled & 7u) n++;
    if ((coupled & 7u) && (lim & (1u << __builtin_ctz(coupled & 7u)))) n++;
    if ((coupled & 56u) && (lim & (1u << __builtin_ctz(coupled & 56u))) && __builtin_popcount(coupled & 56u) == 2) n++;
    return n;
}

float ccd_atan01(float z);   // oracle_ccd.cpp (explicit-arithmetic arctangent shared with the kernels)
static float atan2_poly(float y, float x) {   // atan2 in (-pi, pi] from the [0, 1] polynomial
    const float ax = x < 0.0f ? -x : x, ay = y < 0.0f ? -y : y;
    if (ax == 0.0f && ay == 0.0f) return 0.0f;
    float a = ay <= ax ? ccd_atan01(ay / ax) : 1.5707964f - ccd_atan01(ax / ay);
    if (x < 0.0f) a = 3.1415927f - a;
    return y < 0.0f ? -a : a;
}
static const float FMAX = 3.4028234663852886e38f, FINF = __builtin_inff();
// JointConstraintHelper::recentered_angle (joint_constraint_helper.rs:468-499): the joint angle about the limited axis measured
// from the centre of the allowed range, from the axis' imaginary part x and the real part w of the (sign-corrected) relative
// rotation and [cos, sin] of half the centre angle; wrapped to (-pi, pi].
static float recentered_angle(float x, float w, float c_cos, float c_sin) {
    const float sin_half = c_cos * x - c_sin * w, cos_half = c_cos * w + c_sin * x;
    float half = atan2_poly(sin_half, cos_half);
    if (fabsf(half) > 1.5707964f) half = half - copysignf(3.1415927f, half);
    return half * 2.0f;
}

// JointConstraintHelper::finalize_constraints (joint_constraint_helper.rs:676-722) on rows [a, b)
static void finalize_rows(JointRow* out, int a, int b, V3 imsum) {
    for (int jx = a; jx < b; ++jx) {
        JointRow& cj = out[jx];
        float dot_jj = dot(cj.lin_jac, cmul(imsum, cj.lin_jac)) + dot(cj.ii_ang_jac1, cj.ang_jac1) + dot(cj.ii_ang_jac2, cj.ang_jac2);
        float cfm_gain = dot_jj * cj.cfm_coeff + cj.cfm_gain;
        float inv_dot_jj = inv_or_zero(dot_jj);
        cj.inv_lhs = inv_or_zero(dot_jj + cfm_gain);
        cj.cfm_gain = cfm_gain;
        if (!(cj.lo == -FMAX && cj.hi == FMAX)) continue;   // rows with limited forces are not removed from the others
        for (int ix = jx + 1; ix < b; ++ix) {
            JointRow& ci = out[ix];
            float dot_ij = dot(ci.lin_jac, cmul(imsum, cj.lin_jac)) + dot(ci.ii_ang_jac1, cj.ang_jac1) + dot(ci.ii_ang_jac2, cj.ang_jac2);
            float coeff = dot_ij * inv_dot_jj;
            ci.lin_jac = ci.lin_jac - cj.lin_jac * coeff;
            ci.ang_jac1 = ci.ang_jac1 - cj.ang_jac1 * coeff;
            ci.ang_jac2 = ci.ang_jac2 - cj.ang_jac2 * coeff;
            ci.ii_ang_jac1 = ci.ii_ang_jac1 - cj.ii_ang_jac1 * coeff;
            ci.ii_ang_jac2 = ci.ii_ang_jac2 - cj.ii_ang_jac2 * coeff;
            ci.rhs_wo_bias = ci.rhs_wo_bias - cj.rhs_wo_bias * coeff;
            ci.rhs = ci.rhs - cj.rhs * coeff;
        }
    }
}

// joint_constraint_builder.rs:77-152 -> JointConstraint::update (joint_velocity_constraint.rs:145-357)
// restricted to locked axes; joint_constraint_helper.rs:95-164 (new), :411-461 (lock_linear),
// :628-675 (lock_angular), :676-722 (finalize_constraints).
static int joint_update(const World& w, const Joint& j, float sub_dt, JointRow* out) {
    GatheredBody g1 = gather(w, j.sid1), g2 = gather(w, j.sid2);
    Pose frame1 = pose_mul(g1.pose, j.sframe1);
    Pose frame2 = pose_mul(g2.pose, j.sframe2);
    Spring soft{j.natural_frequency, j.damping_ratio};
    float erp_inv_dt = soft.erp_inv_dt(sub_dt);
    float cfm_coeff = soft.cfm_coeff(sub_dt);
    M3 basis = qto_mat(frame1.q);
    V3 bcol[3] = {basis.c0, basis.c1, basis.c2};
    V3 lin_err = frame2.t - frame1.t;
    V3 new_center1 = frame2.t;
    for (int i = 0; i < 3; ++i)
        if (j.locked_axes & (1u << i)) new_center1 = new_center1 - bcol[i] * dot(lin_err, bcol[i]);
    frame1.t = new_center1;
    V3 r1 = frame1.t - g1.pose.t;
    V3 r2 = frame2.t - g2.pose.t;
    // ang_basis = diff_conj1_2(q1, q2)^T * sgn, ang_err = q1^-1 q2 * sgn (rotation_ops.rs:121-137)
    Q4 q1 = frame1.q, q2 = frame2.q;
    float sgn = copysignf(1.0f, qdot(q1, q2));
    Q4 ang_err = qmul(qconj(q1), q2);
    ang_err = Q4{ang_err.x * sgn, ang_err.y * sgn, ang_err.z * sgn, ang_err.w * sgn};
    V3 a = V3{q1.x, q1.y, q1.z}, b = V3{q2.x, q2.y, q2.z};
    float w1 = q1.w, w2 = q2.w;
    V3 cv = a * w2 + b * w1;
    // D = 0.5 * (a b^T + w1 w2 I - [cv]x + [a]x [b]x); [a]x[b]x = b a^T - (a.b) I
    float ab = dot(a, b);
    float D[3][3];
    float av[3] = {a.x, a.y, a.z}, bv[3] = {b.x, b.y, b.z}, cvv[3] = {cv.x, cv.y, cv.z};
    for (int r = 0; r < 3; ++r)
        for (int c = 0; c < 3; ++c) {
            float cx = 0.0f;  // [cv]x(r,c)
            if (r == 0 && c == 1) cx = -cvv[2]; else if (r == 0 && c == 2) cx = cvv[1];
            else if (r == 1 && c == 0) cx = cvv[2]; else if (r == 1 && c == 2) cx = -cvv[0];
            else if (r == 2 && c == 0) cx = -cvv[1]; else if (r == 2 && c == 1) cx = cvv[0];
            float diag = (r == c) ? (w1 * w2 - ab) : 0.0f;
            D[r][c] = (av[r] * bv[c] + diag - cx + bv[r] * av[c]) * 0.5f;
        }
    // ang_basis column i = row i of D (transpose), times sgn.
    V3 imsum = g1.im + g2.im;
    const uint32_t free_axes = ~j.locked_axes & 63u;
    const uint32_t motor_axes = j.motor_axes & free_axes, limit_axes = j.limit_axes & free_axes;
    const uint32_t coupled = j.coupled_axes;   // joint_velocity_constraint.rs:163-178
    const bool has_lin_coupling = (coupled & 7u) != 0, has_ang_coupling = (coupled & 56u) != 0;
    const int first_lin = has_lin_coupling ? __builtin_ctz(coupled & 7u) : 0, first_ang = has_ang_coupling ? __builtin_ctz(coupled & 56u) : 0;
    const float inv_dt = sub_dt == 0.0f ? 0.0f : 1.0f / sub_dt;
    const float max_bias = w.params.max_corrective_velocity();
    const float aerr[3] = {ang_err.x, ang_err.y, ang_err.z};
    auto lock_linear_row = [&](int i, float erp, float cfm, int dof, int kind) {   // lock_linear (joint_constraint_helper.rs:411-461)
        JointRow r;
        r.lin_jac = bcol[i];
        r.ang_jac1 = cross(r1, bcol[i]);
        r.ang_jac2 = cross(r2, bcol[i]);
        r.ii_ang_jac1 = sdp_mul(g1.ii, r.ang_jac1);
        r.ii_ang_jac2 = sdp_mul(g2.ii, r.ang_jac2);
        r.impulse = 0.0f; r.inv_lhs = 0.0f; r.cfm_coeff = cfm; r.cfm_gain = 0.0f;
        r.rhs_wo_bias = 0.0f;
        r.rhs = 0.0f + dot(bcol[i], lin_err) * erp;
        r.lo = -FMAX; r.hi = FMAX;
        r.dof = dof; r.kind = kind;
        return r;
    };
    auto motor_coeffs = [&](const RbJointMotor& m, float& m_erp, float& m_cfm_coeff, float& m_cfm_gain) {   // MotorModel::combine_coefficients
        m_erp = m.stiffness * inv_or_zero(sub_dt * m.stiffness + m.damping);
        const float c = inv_or_zero(sub_dt * sub_dt * m.stiffness + sub_dt * m.damping);
        m_cfm_coeff = m.model == 0 ? c : 0.0f;
        m_cfm_gain = m.model == 0 ? 0.0f : c;
    };
    int len = 0;
    // ---- motors (joint_velocity_constraint.rs:191-223): angular then linear, orthogonalised among themselves
    for (int i = 3; i < 6; ++i) {
        if (!((motor_axes & ~coupled) & (1u << i))) continue;
        const RbJointMotor& m = j.motors[i];
        float m_erp, m_cc, m_cg;
        motor_coeffs(m, m_erp, m_cc, m_cg);
        JointRow& r = out[len++];   // motor_angular (joint_constraint_helper.rs:566-626)
        const V3 ang_jac = bcol[i - 3];
        float rhs_wo_bias = 0.0f;
        if (m_erp != 0.0f) {
            const float ce = fclamp(aerr[i - 3], -1.0f, 1.0f);
            const float ang_dist = atan2_poly(ce, sqrtf(fmax2(1.0f - ce * ce, 0.0f))) * 2.0f;   // asin(clamped) * 2
            float s_err = ang_dist - m.target_pos;   // utils::smallest_abs_diff_between_angles
            const float sg = s_err > 0.0f ? 1.0f : (s_err < 0.0f ? -1.0f : 0.0f);
            const float comp = s_err - sg * 6.2831855f;
            if (!(fabsf(s_err) < fabsf(comp))) s_err = comp;
            rhs_wo_bias = rhs_wo_bias + s_err * m_erp;
        }
        rhs_wo_bias = rhs_wo_bias + -m.target_vel;
        r.lin_jac = vzero(); r.ang_jac1 = ang_jac; r.ang_jac2 = ang_jac;
        r.ii_ang_jac1 = sdp_mul(g1.ii, ang_jac); r.ii_ang_jac2 = sdp_mul(g2.ii, ang_jac);
        r.impulse = 0.0f; r.inv_lhs = 0.0f; r.cfm_coeff = m_cc; r.cfm_gain = m_cg;
        r.rhs = rhs_wo_bias; r.rhs_wo_bias = rhs_wo_bias;
        r.lo = -(m.max_force * sub_dt); r.hi = m.max_force * sub_dt;
        r.dof = i; r.kind = 2;
    }
    for (int i = 0; i < 3; ++i) {
        if (!((motor_axes & ~coupled) & (1u << i))) continue;
        const RbJointMotor& m = j.motors[i];
        float m_erp, m_cc, m_cg;
        motor_coeffs(m, m_erp, m_cc, m_cg);
        JointRow r = lock_linear_row(i, 0.0f, 0.0f, i, 2);   // motor_linear (joint_constraint_helper.rs:285-330)
        float rhs_wo_bias = 0.0f;
        if (m_erp != 0.0f) rhs_wo_bias = rhs_wo_bias + (dot(lin_err, r.lin_jac) - m.target_pos) * m_erp;
        float target_vel = m.target_vel;
        if (limit_axes & (1u << i)) {
            const float dist = dot(lin_err, r.lin_jac);
            target_vel = fclamp(target_vel, (j.limits[i][0] - dist) * inv_dt, (j.limits[i][1] - dist) * inv_dt);
        }
        rhs_wo_bias = rhs_wo_bias + -target_vel;
        r.cfm_coeff = m_cc; r.cfm_gain = m_cg;
        r.lo = -(m.max_force * sub_dt); r.hi = m.max_force * sub_dt;
        r.rhs = rhs_wo_bias; r.rhs_wo_bias = rhs_wo_bias;
        out[len++] = r;
    }
    // the distance row of the coupled linear axes (limit_linear_coupled / motor_linear_coupled, joint_constraint_helper.rs:210-283, :333-409)
    auto coupled_linear_row = [&](float& dist) {
        JointRow r;
        V3 lin_jac = vzero(), ang_jac1 = vzero(), ang_jac2 = vzero();
        for (int i = 0; i < 3; ++i) {
            if (!(coupled & (1u << i))) continue;
            const float coeff = dot(bcol[i], lin_err);
            lin_jac = lin_jac + bcol[i] * coeff;
            ang_jac1 = ang_jac1 + cross(r1, bcol[i]) * coeff;
            ang_jac2 = ang_jac2 + cross(r2, bcol[i]) * coeff;
        }
        dist = sqrtf(dot(lin_jac, lin_jac));
        const float inv_dist = inv_or_zero(dist);
        r.lin_jac = lin_jac * inv_dist; r.ang_jac1 = ang_jac1 * inv_dist; r.ang_jac2 = ang_jac2 * inv_dist;
        r.ii_ang_jac1 = sdp_mul(g1.ii, r.ang_jac1); r.ii_ang_jac2 = sdp_mul(g2.ii, r.ang_jac2);
        r.impulse = 0.0f; r.inv_lhs = 0.0f;
        return r;
    };
    // (motor_axes & coupled_axes) & ANG_AXES: "TODO: coupled angular motor constraint" in the reference -- no row (:224-226)
    if ((motor_axes & coupled) & 7u) {   // motor_linear_coupled (:228-250): the motor of the first coupled linear axis acts on the distance
        const RbJointMotor& m = j.motors[first_lin];
        float m_erp, m_cc, m_cg;
        motor_coeffs(m, m_erp, m_cc, m_cg);
        float dist;
        JointRow r = coupled_linear_row(dist);
        float rhs_wo_bias = 0.0f;
        if (m_erp != 0.0f) rhs_wo_bias = rhs_wo_bias + (dist - m.target_pos) * m_erp;
        float target_vel = m.target_vel;
        if (limit_axes & (1u << first_lin)) target_vel = fclamp(target_vel, (j.limits[first_lin][0] - dist) * inv_dt, (j.limits[first_lin][1] - dist) * inv_dt);
        rhs_wo_bias = rhs_wo_bias + -target_vel;
        r.cfm_coeff = m_cc; r.cfm_gain = m_cg;
        r.lo = -(m.max_force * sub_dt); r.hi = m.max_force * sub_dt;
        r.rhs = rhs_wo_bias; r.rhs_wo_bias = rhs_wo_bias;
        r.dof = first_lin; r.kind = 2;
        out[len++] = r;
    }
    finalize_rows(out, 0, len, imsum);
    const int start = len;
    // ---- locked axes, then limits (joint_velocity_constraint.rs:259-355), orthogonalised together
    for (int i = 3; i < 6; ++i) {
        if (!(j.locked_axes & (1u << i))) continue;
        int ax = i - 3;
        JointRow& r = out[len++];
        V3 ang_jac = V3{D[ax][0] * sgn, D[ax][1] * sgn, D[ax][2] * sgn};
        float imag = ax == 0 ? ang_err.x : (ax == 1 ? ang_err.y : ang_err.z);
        r.lin_jac = vzero();
        r.ang_jac1 = ang_jac;
        r.ang_jac2 = ang_jac;
        r.ii_ang_jac1 = sdp_mul(g1.ii, ang_jac);
        r.ii_ang_jac2 = sdp_mul(g2.ii, ang_jac);
        r.impulse = 0.0f; r.inv_lhs = 0.0f; r.cfm_coeff = cfm_coeff; r.cfm_gain = 0.0f;
        r.rhs_wo_bias = 0.0f;
        r.rhs = 0.0f + imag * erp_inv_dt;
        r.lo = -FMAX; r.hi = FMAX;
        r.dof = i; r.kind = 0;
    }
    for (int i = 0; i < 3; ++i) {
        if (!(j.locked_axes & (1u << i))) continue;
        out[len++] = lock_linear_row(i, erp_inv_dt, cfm_coeff, i, 0);
    }
    for (int i = 3; i < 6; ++i) {
        if (!((limit_axes & ~coupled) & (1u << i))) continue;
        const int ax = i - 3;
        // recentered_angle (joint_constraint_helper.rs:468-499) + limit_angular (:503-564)
        const float c_cos = j.ang_limit_center[ax][0], c_sin = j.ang_limit_center[ax][1], half_range = j.ang_limit_half_range[ax];
        const float ang = recentered_angle(aerr[ax], ang_err.w, c_cos, c_sin);
        const bool min_enabled = ang <= -half_range, max_enabled = half_range <= ang;
        JointRow& r = out[len++];
        const V3 ang_jac = bcol[ax];
        const float rhs_bias = fclamp((fmax2(ang - half_range, 0.0f) - fmax2(-half_range - ang, 0.0f)) * erp_inv_dt, -max_bias, max_bias);
        r.lin_jac = vzero(); r.ang_jac1 = ang_jac; r.ang_jac2 = ang_jac;
        r.ii_ang_jac1 = sdp_mul(g1.ii, ang_jac); r.ii_ang_jac2 = sdp_mul(g2.ii, ang_jac);
        r.impulse = 0.0f; r.inv_lhs = 0.0f; r.cfm_coeff = cfm_coeff; r.cfm_gain = 0.0f;
        r.rhs_wo_bias = 0.0f;
        r.rhs = 0.0f + rhs_bias;
        r.lo = min_enabled ? -FINF : 0.0f; r.hi = max_enabled ? FINF : 0.0f;
        r.dof = i; r.kind = 1;
    }
    for (int i = 0; i < 3; ++i) {
        if (!((limit_axes & ~coupled) & (1u << i))) continue;
        JointRow r = lock_linear_row(i, erp_inv_dt, cfm_coeff, i, 1);   // limit_linear (joint_constraint_helper.rs:166-207)
        const float dist = dot(lin_err, r.lin_jac);
        const bool min_enabled = dist <= j.limits[i][0], max_enabled = j.limits[i][1] <= dist;
        const float rhs_bias = fclamp((fmax2(dist - j.limits[i][1], 0.0f) - fmax2(j.limits[i][0] - dist, 0.0f)) * erp_inv_dt, -max_bias, max_bias);
        r.rhs = r.rhs_wo_bias + rhs_bias;
        r.cfm_coeff = cfm_coeff;
        r.lo = min_enabled ? -FINF : 0.0f; r.hi = max_enabled ? FINF : 0.0f;
        out[len++] = r;
    }
    if (has_ang_coupling && (limit_axes & (1u << first_ang))) {   // limit_angular_coupled (joint_constraint_helper.rs:725-798): exactly two coupled angular axes
        const uint32_t ac = coupled >> 3;
        const int not_coupled = (ac & 1u) == 0 ? 0 : ((ac & 2u) == 0 ? 1 : ((ac & 4u) == 0 ? 2 : 3));   // trailing_ones
        if (not_coupled < 3) {
            const M3 basis2 = qto_mat(frame2.q);
            const V3 b2col[3] = {basis2.c0, basis2.c1, basis2.c2};
            const V3 axis1 = bcol[not_coupled], axis2 = b2col[not_coupled];
            // Rot3::from_rotation_arc(axis1, axis2).to_axis_angle() (glam; third-party arithmetic restated, atan2 by the shared polynomial)
            const float d = dot(axis1, axis2);
            const float one_minus_eps = 1.0f - 2.0f * 1.1920929e-7f;
            Q4 rot;
            if (d > one_minus_eps) rot = Q4{0.0f, 0.0f, 0.0f, 1.0f};
            else if (d < -one_minus_eps) {   // half turn about any orthonormal vector
                const float sg = copysignf(1.0f, axis1.z), a = -1.0f / (sg + axis1.z), b = axis1.x * axis1.y * a;
                rot = Q4{b, sg + axis1.y * axis1.y * a, -axis1.y, -4.371139e-8f};
            } else {
                const V3 c = cross(axis1, axis2);
                const float ww = 1.0f + d;
                const float inv = 1.0f / sqrtf(c.x * c.x + c.y * c.y + c.z * c.z + ww * ww);
                rot = Q4{c.x * inv, c.y * inv, c.z * inv, ww * inv};
            }
            V3 ang_jac = V3{1.0f, 0.0f, 0.0f};
            float angle = 0.0f;
            const V3 v = V3{rot.x, rot.y, rot.z};
            const float vl = sqrtf(dot(v, v));
            if (vl >= 1.0e-8f) { angle = 2.0f * atan2_poly(vl, rot.w); ang_jac = v * (1.0f / vl); }
            if (angle == 0.0f) {   // axis1.orthonormal_basis()[0] (utils/orthonormal_basis.rs:76-86)
                const float sg = copysignf(1.0f, axis1.z), a = -1.0f / (sg + axis1.z), b = axis1.x * axis1.y * a;
                ang_jac = V3{1.0f + sg * axis1.x * axis1.x * a, sg * b, -sg * axis1.x};
            }
            const float lo = j.limits[first_ang][0], hi = j.limits[first_ang][1];
            const bool min_enabled = angle <= lo, max_enabled = hi <= angle;
            const float rhs_bias = fclamp((fmax2(angle - hi, 0.0f) - fmax2(lo - angle, 0.0f)) * erp_inv_dt, -max_bias, max_bias);
            JointRow& r = out[len++];
            r.lin_jac = vzero(); r.ang_jac1 = ang_jac; r.ang_jac2 = ang_jac;
            r.ii_ang_jac1 = sdp_mul(g1.ii, ang_jac); r.ii_ang_jac2 = sdp_mul(g2.ii, ang_jac);
            r.impulse = 0.0f; r.inv_lhs = 0.0f; r.cfm_coeff = cfm_coeff; r.cfm_gain = 0.0f;
            r.rhs_wo_bias = 0.0f;
            r.rhs = 0.0f + rhs_bias;
            r.lo = min_enabled ? -FINF : 0.0f; r.hi = max_enabled ? FINF : 0.0f;
            r.dof = first_ang; r.kind = 1;
        }
    }
    if (has_lin_coupling && (limit_axes & (1u << first_lin))) {   // limit_linear_coupled (:210-283): max distance only ("FIXME: handle min limit too")
        float dist;
        JointRow r = coupled_linear_row(dist);
        const float hi = j.limits[first_lin][1];
        r.rhs_wo_bias = fmin2(dist - hi, 0.0f) * inv_dt;
        const float rhs_bias = fclamp(fmax2(dist - hi, 0.0f) * erp_inv_dt, -max_bias, max_bias);
        r.rhs = r.rhs_wo_bias + rhs_bias;
        r.cfm_coeff = cfm_coeff; r.cfm_gain = 0.0f;
        r.lo = 0.0f; r.hi = FINF;
        r.dof = first_lin; r.kind = 1;
        out[len++] = r;
    }
    finalize_rows(out, start, len, imsum);
    return len;
}

// joint_velocity_constraint.rs:97-124
static void joint_solve(World& w, const Joint& j, JointRow* rows, int n, bool wo_bias, bool warm = false) {
    GatheredBody g1 = gather(w, j.sid1), g2 = gather(w, j.sid2);
    V3 v1 = g1.lin, w1 = g1.ang, v2 = g2.lin, w2 = g2.ang;
    for (int k = 0; k < n; ++k) {
        JointRow& r = rows[k];
        if (wo_bias) r.rhs = r.rhs_wo_bias;
        if (warm) {   // warmstart_generic (joint_velocity_constraint.rs:129-141), row by row right before its solve (solve.rs:41)
            V3 li = r.lin_jac * r.impulse;
            v1 = maddv(v1, li, g1.im);
            w1 = madd(w1, r.ii_ang_jac1, r.impulse);
            v2 = maddv(v2, -li, g2.im);
            w2 = madd(w2, r.ii_ang_jac2, -r.impulse);
        }
        float dlinvel = dot(r.lin_jac, v2 - v1);
        float dangvel = dot(r.ang_jac2, w2) - dot(r.ang_jac1, w1);
        float rhs = dlinvel + dangvel + r.rhs;
        float total = fclamp(r.impulse + r.inv_lhs * (rhs - r.cfm_gain * r.impulse), r.lo, r.hi);   // impulse_bounds
        float delta = total - r.impulse;
        r.impulse = total;
        V3 lin_impulse = r.lin_jac * delta;
        v1 = maddv(v1, lin_impulse, g1.im);
        w1 = madd(w1, r.ii_ang_jac1, delta);
        v2 = maddv(v2, -lin_impulse, g2.im);
        w2 = madd(w2, r.ii_ang_jac2, -delta);
    }
    scatter_vel(w, j.sid1, v1, w1);
    scatter_vel(w, j.sid2, v2, w2);
}

// Stage order (init.rs:163-254): colours with >= `min_count` members in ascending colour, then the
// serial tail: smaller colours ascending, then the overflow colour.
static void build_order(const std::vector<int>& colors_of, int min_count, std::vector<int>& order,
                        std::vector<int>& stage_start, int& num_colors) {
    std::vector<std::vector<int>> buckets(NUM_COLORS);
    for (int i = 0; i < (int)colors_of.size(); ++i) {
        int c = colors_of[i];
        if (c < 0) continue;
        if (c > 128) c = 128;
        buckets[c].push_back(i);
    }
    order.clear();
    stage_start.clear();
    num_colors = 0;
    for (int pass = 0; pass < 2; ++pass)
        for (int c = 0; c < 128; ++c) {
            bool big = (int)buckets[c].size() >= min_count;
            if (buckets[c].empty() || big != (pass == 0)) continue;
            stage_start.push_back((int)order.size());
            order.insert(order.end(), buckets[c].begin(), buckets[c].end());
            num_colors++;
        }
    // overflow: sequential, one stage per element.
    for (int i : buckets[128]) {
        stage_start.push_back((int)order.size());
        order.push_back(i);
    }
    if (!buckets[128].empty()) num_colors++;
    stage_start.push_back((int)order.size());
}


static inline float inv_exact0(float x) { return x == 0.0f ? 0.0f : 1.0f / x; }

float ccd_quat_angle(float vlen, float w);   // oracle_ccd.cpp

void solve_island(World& w, V3 gravity) {
    const RbIntegrationParameters& P = w.params.p;
    // interpolate_kinematic_velocities (substep.rs:242-265; rigid_body_components.rs:147-196)
    for (Body& b : w.bodies) {
        if (b.type != RB_BODY_KINEMATIC_POSITION_BASED || !b.is_awake()) continue;
        const float inv_dt = P.dt == 0.0f ? 0.0f : 1.0f / P.dt;
        const V3 dl = pose_point(b.kin_target, b.local_com) - pose_point(b.pos, b.local_com);
        const Q4 dq = qmul(b.kin_target.q, qconj(b.pos.q));
        const V3 dv = V3{dq.x, dq.y, dq.z};
        const float len = length(dv);
        V3 sa = vzero();
        if (len > 1.0e-12f) {
            float angle = ccd_quat_angle(len, dq.w);
            if (dq.w < 0.0f) angle = 6.2831855f - angle;
            sa = dv * (angle / len);
        }
        b.linvel = dl * inv_dt;
        b.angvel = sa * inv_dt;
    }
    const int nb = (int)w.bodies.size();
    const float inv_dt_full = P.dt == 0.0f ? 0.0f : 1.0f / P.dt;

    // Substep solve-groups (island_manager/substep_groups.rs; staged_island_solver/init.rs:52-91): every island runs
    // num_solver_iterations + key substeps of dt / that count, key = the largest additional_solver_iterations among its
    // awake members.  Islands share no solver body (kinematic bodies are island members here, oracle_internal.h), so the
    // groups are independent; they run in the reference's order, highest cadence first.
    std::vector<int> bkey(nb, 0);
    std::vector<int> group_keys;
    if ((int)w.island_of.size() == nb) {
        std::vector<int> ikey(nb, 0);
        for (int i = 0; i < nb; ++i) {
            const Body& b = w.bodies[i];
            const int extra = RB_BODY_EXTRA_ITERS_OF(b.flags);
            if (extra > 0 && b.is_awake() && w.island_of[i] >= 0) ikey[w.island_of[i]] = std::max(ikey[w.island_of[i]], extra);
        }
        bool seen[256] = {false};
        for (int i = 0; i < nb; ++i) {
            if (!w.bodies[i].is_awake() || w.island_of[i] < 0) continue;
            bkey[i] = ikey[w.island_of[i]];
            seen[bkey[i]] = true;
        }
        for (int k = 255; k >= 0; --k)
            if (seen[k]) group_keys.push_back(k);
    }
    if (group_keys.empty()) group_keys.push_back(0);
    const bool multi = group_keys.size() > 1 || group_keys[0] != 0;

    // a7 forces (solve.rs:234-291; rigid_body_components.rs:1030-1033) + S1 solver-body init
    // (solver_body.rs:82-121; worker.rs:46-104).
    w.sb.resize(nb);
    for (int i = 0; i < nb; ++i) {
        Body& b = w.bodies[i];
        SolverBody& s = w.sb[i];
        if (b.is_awake()) {
            V3 eff_mass = V3{inv_exact0(b.eff_inv_mass.x), inv_exact0(b.eff_inv_mass.y), inv_exact0(b.eff_inv_mass.z)};
            b.force = b.user_force + cmul(gravity, eff_mass) * b.gravity_scale;
            b.torque = b.user_torque;
        }
        s.flags = b.flags;
        s.lin = b.linvel;
        s.ang = b.angvel;
        s.pose = pose_prepend_translation(b.pos, b.local_com);
        if (b.is_awake()) {
            s.ii = b.eff_world_inv_inertia;
            s.im = b.eff_inv_mass;
        } else {
            s.ii = sdp_zero();
            s.im = vzero();
        }
        const float body_sub_dt = P.dt / (float)(P.num_solver_iterations + bkey[i]);   // the slot's group substep dt (worker.rs:66-77)
        s.incr_ang = sdp_mul(b.eff_world_inv_inertia, b.torque) * body_sub_dt;
        s.incr_lin = cmul(b.force, b.eff_inv_mass) * body_sub_dt;
        s.gyro = (b.flags & RB_BODY_GYROSCOPIC) && b.is_strict_dynamic() && b.is_awake();   // worker.rs:86
    }

    // Solver-active manifolds, in stage order (solver_graph.rs:129-361; init.rs:163-254).
    std::vector<int> colors_of(w.pairs.size(), -1);
    for (int i = 0; i < (int)w.pairs.size(); ++i) {
        const Pair& p = w.pairs[i];
        bool d1 = p.b1 >= 0 && w.bodies[p.b1].is_awake();
        bool d2 = p.b2 >= 0 && w.bodies[p.b2].is_awake();
        if (p.nsc > 0 && (d1 || d2)) colors_of[i] = p.color;
    }
    std::vector<int> pair_order;
    int num_colors = 0;
    build_order(colors_of, 125, pair_order, w.order_color_start, num_colors);  // ceil(n/4) >= 32
    w.last_num_colors = num_colors;
    const int ncons = (int)pair_order.size();
    w.cons.resize(ncons);
    w.order.resize(ncons);
    for (int i = 0; i < ncons; ++i) w.order[i] = i;
    const std::vector<int>& cs = w.order_color_start;
    const int nstages = (int)cs.size() - 1;

    // Joints: selection (impulse_joint_set.rs:504-572) + stage order (joints.rs:318-392).
    const int nj = (int)w.joints.size();
    std::vector<int> jcolors(nj, -1);
    std::vector<int> jrow_start(nj + 1, 0);
    for (int i = 0; i < nj; ++i) {
        Joint& j = w.joints[i];
        const Body& b1 = w.bodies[j.body1];
        const Body& b2 = w.bodies[j.body2];
        const bool active = !j.removed && (b1.is_awake() || b2.is_awake());   // joints of a sleeping island are not solved
        jcolors[i] = active ? j.color : -1;
        j.sid1 = b1.is_awake() ? (uint32_t)j.body1 : NO_BODY;
        j.sid2 = b2.is_awake() ? (uint32_t)j.body2 : NO_BODY;
        // generic_joint.rs:624-636 transform_to_solver_body_space
        j.sframe1 = j.local_frame1;
        j.sframe2 = j.local_frame2;
        if (!b1.is_awake()) j.sframe1 = pose_mul(b1.pos, j.local_frame1);
        else j.sframe1.t = j.local_frame1.t - b1.local_com;
        if (!b2.is_awake()) j.sframe2 = pose_mul(b2.pos, j.local_frame2);
        else j.sframe2.t = j.local_frame2.t - b2.local_com;
        jrow_start[i + 1] = jrow_start[i] + joint_rows_of(j);
    }
    int jnum_colors = 0;
    build_order(jcolors, 64, w.jorder, w.jorder_color_start, jnum_colors);
    w.jrows.resize(jrow_start[nj]);
    const std::vector<int>& js = w.jorder_color_start;
    const int jnstages = (int)js.size() - 1;

    Pool& pool = Pool::get();

    // S2 constraint generation (worker.rs:109-190)
    pool.parallel_for(0, ncons, 64, [&](int i) { generate(w, pair_order[i], w.cons[i]); });

    bool has_bouncy = false;
    for (int i = 0; i < ncons && !has_bouncy; ++i)
        for (int k = 0; k < w.cons[i].num_contacts; ++k)
            if (w.cons[i].b_restitution_seed[k] < 0.0f) has_bouncy = true;

    const bool fused_warmstart = P.warmstart_coefficient != 0.0f;
    const float max_lin = w.params.max_linear_velocity();
    const float max_ang = 0.7853981633974483f * inv_dt_full;  // MAX_ROTATION * base inv_dt (worker.rs:573-580)

    // group of every scheduled constraint / joint = the group of its awake body
    std::vector<int> ckey(ncons, 0), jkey(nj, 0);
    if (multi) {
        for (int q = 0; q < ncons; ++q) {
            const Pair& p = w.pairs[pair_order[q]];
            ckey[q] = bkey[(p.b1 >= 0 && w.bodies[p.b1].is_awake()) ? p.b1 : p.b2];
        }
        for (int i = 0; i < nj; ++i) {
            const Joint& j = w.joints[i];
            if (jcolors[i] >= 0) jkey[i] = bkey[w.bodies[j.body1].is_awake() ? j.body1 : j.body2];
        }
    }
    for (const int key : group_keys) {   // the group ring (worker.rs:193-207)
    const int num_substeps = P.num_solver_iterations + key;
    const float sub_dt = P.dt / (float)num_substeps;  // init.rs:64-77, :96-101
    SubParams sp;
    sp.dt = sub_dt;
    sp.inv_dt = sub_dt == 0.0f ? 0.0f : 1.0f / sub_dt;
    Spring dyn_soft{P.contact_natural_frequency, P.contact_damping_ratio};
    Spring static_soft{P.static_contact_natural_frequency, P.static_contact_damping_ratio};
    sp.dyn_cfm = dyn_soft.cfm_factor(sub_dt);
    sp.static_cfm = static_soft.cfm_factor(sub_dt);
    sp.dyn_erp = dyn_soft.erp_inv_dt(sub_dt);
    sp.static_erp = static_soft.erp_inv_dt(sub_dt);
    sp.max_corrective_velocity = w.params.max_corrective_velocity();
    sp.warmstart_coeff = P.warmstart_coefficient;

    auto solve_pass = [&](bool wo_bias, float solved_dt, bool warm_joints = false) {  // staged_island_solver/solve.rs:12-209
        bool solve_friction = wo_bias || P.friction_in_bias_pass || P.num_internal_stabilization_iterations == 0;
        for (int s = 0; s < jnstages; ++s) {
            pool.parallel_for(js[s], js[s + 1], 64, [&](int q) {
                int ji = w.jorder[q];
                if (multi && jkey[ji] != key) return;
                joint_solve(w, w.joints[ji], &w.jrows[jrow_start[ji]], jrow_start[ji + 1] - jrow_start[ji], wo_bias, warm_joints);
            });
        }
        for (int s = 0; s < nstages; ++s) {
            pool.parallel_for(cs[s], cs[s + 1], 64, [&](int q) {
                if (multi && ckey[q] != key) return;
                Constraint& c = w.cons[q];
                if (wo_bias) refresh_rhs_wo_bias(w, c, sp, solved_dt);
                solve(w, c, solve_friction);
            });
        }
    };

    for (int substep = 0; substep < num_substeps; ++substep) {
        float solved_dt = (float)substep * sub_dt;
        // S3 velocity increments + gyroscopic correction (worker.rs:235-284)
        pool.parallel_for(0, nb, 256, [&](int i) {
            SolverBody& s = w.sb[i];
            if (!w.bodies[i].is_awake() || (multi && bkey[i] != key)) return;
            s.lin = s.lin + s.incr_lin;
            s.ang = s.ang + s.incr_ang;
            if (s.gyro) {
                const Body& b = w.bodies[i];
                Q4 principal_axes = qmul(s.pose.q, b.principal_frame);
                s.ang = gyroscopic_corrected_angvel(s.ang, principal_axes, b.principal_inertia, b.inv_principal_inertia, sub_dt);
            }
        });
        // S4 joint rows rebuilt from the current poses (worker.rs:291-432); impulses restart from 0, or with
        // warmstart_joints from last step's written-back impulses (first substep) / the previous substep's rows, times
        // warmstart_coefficient (joint_constraint_builder.rs:116-150).
        pool.parallel_for(0, nj, 64, [&](int i) {
            if (jcolors[i] < 0 || (multi && jkey[i] != key)) return;
            JointRow* rows = &w.jrows[jrow_start[i]];
            const Joint& j = w.joints[i];
            float prev[24];
            const int nprev = jrow_start[i + 1] - jrow_start[i];
            if (P.warmstart_joints && substep > 0)
                for (int k = 0; k < nprev && k < 24; ++k) prev[k] = rows[k].impulse;
            const int len = joint_update(w, j, sub_dt, rows);
            if (P.warmstart_joints) {
                for (int k = 0; k < len && k < 24; ++k) {
                    JointRow& r = rows[k];
                    float seed = substep > 0 ? prev[k]
                                             : (r.kind == 0 ? j.impulses[r.dof] : r.kind == 1 ? j.limit_impulses[r.dof] : j.motor_impulses[r.dof]);
                    r.impulse = seed * P.warmstart_coefficient;
                }
            }
        });
        // S5 update + warmstart, colour by colour (worker.rs:438-539)
        if (!fused_warmstart) {
            for (int q = 0; q < ncons; ++q) { if (multi && ckey[q] != key) continue; update(w, w.cons[q], sp, solved_dt); }
        } else {
            for (int s = 0; s < nstages; ++s) {
                pool.parallel_for(cs[s], cs[s + 1], 64, [&](int q) {
                    if (multi && ckey[q] != key) return;
                    update(w, w.cons[q], sp, solved_dt);
                    warmstart(w, w.cons[q]);
                });
            }
        }
        // S6 biased solve (worker.rs:544-561)
        for (int it = 0; it < P.num_internal_pgs_iterations; ++it) solve_pass(false, solved_dt, P.warmstart_joints != 0 && it == 0);   // worker.rs:548
        // S7 integrate positions (worker.rs:568-631; rigid_body_components.rs:884-898)
        pool.parallel_for(0, nb, 256, [&](int i) {
            SolverBody& s = w.sb[i];
            if (!w.bodies[i].is_awake() || (multi && bkey[i] != key)) return;
            if (max_lin != 3.4028235e38f) {
                float n = length(s.lin);
                if (n > max_lin) s.lin = s.lin * (max_lin / n);
            }
            if (!(s.flags & RB_BODY_ALLOW_FAST_ROTATION)) {
                float n = length(s.ang);
                if (n > max_ang) s.ang = s.ang * (max_ang / n);
            }
            V3 hang = s.ang * (sub_dt * 0.5f);
            Q4 id_plus_hang = Q4{hang.x, hang.y, hang.z, 1.0f};
            s.pose.q = qnormalize(qmul(id_plus_hang, s.pose.q));
            s.pose.t = madd(s.pose.t, s.lin, sub_dt);
        });
        // S8 relax solve (worker.rs:636-649)
        for (int it = 0; it < P.num_internal_stabilization_iterations; ++it) solve_pass(true, solved_dt + sub_dt);
    }
    }   // (groups)

    // S9 restitution (worker.rs:657-734)
    if (has_bouncy) {
        for (int s = 0; s < nstages; ++s)
            for (int q = cs[s]; q < cs[s + 1]; ++q) apply_restitution(w, w.cons[q]);
    }
    // S10 impulse writeback (worker.rs:742-802)
    for (int q = 0; q < ncons; ++q) writeback_impulses(w, w.cons[q]);
    for (int i = 0; i < nj; ++i) {
        Joint& j = w.joints[i];
        if (jcolors[i] < 0) continue;
        for (int r = jrow_start[i]; r < jrow_start[i + 1]; ++r) {
            const JointRow& row = w.jrows[r];
            (row.kind == 0 ? j.impulses : (row.kind == 1 ? j.limit_impulses : j.motor_impulses))[row.dof] = row.impulse;
        }
    }
    // S11 body writeback (worker.rs:809-897; rigid_body_components.rs:835-841)
    for (int i = 0; i < nb; ++i) {
        Body& b = w.bodies[i];
        if (!b.is_awake()) continue;
        const SolverBody& s = w.sb[i];
        V3 lin = s.lin * (1.0f / (1.0f + P.dt * b.lin_damping));
        V3 ang = s.ang * (1.0f / (1.0f + P.dt * b.ang_damping));
        Pose np = pose_prepend_translation(s.pose, -b.local_com);
        auto fin3 = [](V3 v) { return std::isfinite(v.x) && std::isfinite(v.y) && std::isfinite(v.z); };
        if (!(fin3(lin) && fin3(ang) && fin3(np.t) && std::isfinite(np.q.x) && std::isfinite(np.q.y) && std::isfinite(np.q.z) && std::isfinite(np.q.w))) {
            // quarantine.rs:126-178: roll back to the last valid pose, zero the dynamics, disable, report
            b.linvel = vzero(); b.angvel = vzero(); b.user_force = vzero(); b.user_torque = vzero();
            b.next_pos = b.pos;
            b.type = 7;
            for (Collider& c : w.colliders)
                if (c.parent == i) c.shape = -1;
            w.quarantine.push_back(i);
            w.bp_dirty = true; w.static_dirty = true; w.islands_dirty = true;
            continue;
        }
        b.linvel = lin;
        b.angvel = ang;
        b.next_pos = b.type == RB_BODY_KINEMATIC_POSITION_BASED ? b.kin_target : np;   // worker.rs:836-842
    }
    w.counters.num_active_manifolds = ncons;
    w.counters.num_colors = num_colors;
}


// ---------------------------------------------------------------------------------------------
// Unit-level known-answer entry points (tests/test_ref_vectors.py; layouts documented in
// tests/golden/make_ref_vectors.py, shared with rb_debug_kat of the CUDA library).
// ---------------------------------------------------------------------------------------------
int kat_solver(const char* name_c, const float* in, int n_in, float* out, int n_out) {
    const std::string name(name_c);
    auto v3at = [&](int o) { return V3{in[o], in[o + 1], in[o + 2]}; };
    auto put3 = [&](int o, V3 v) { out[o] = v.x; out[o + 1] = v.y; out[o + 2] = v.z; };
    if (name == "generate") {   // contact_with_twist_friction.rs:58-424 on a synthetic, world-attached manifold
        if (n_in < 6 || n_out < 39) return -3;
        World w;
        Pair p{};
        p.c1 = 0; p.c2 = 1; p.b1 = -1; p.b2 = -1;
        p.normal = v3at(0); p.friction = in[3]; p.restitution = in[4];
        const int n = (int)in[5];
        if (n < 1 || n > MAX_MANIFOLD_POINTS || n_in < 6 + 19 * n) return -3;
        p.nsc = n; p.npts = n;
        for (int k = 0; k < n; ++k) {
            const int o = 6 + 19 * k;
            p.sc[k].anchor1 = v3at(o); p.sc[k].anchor2 = v3at(o + 3); p.sc[k].cid = (int)in[o + 6];
            Point& pt = p.pts[p.sc[k].cid];
            pt.impulse = in[o + 7]; pt.warmstart_impulse = in[o + 8]; pt.warmstart_twist = in[o + 9];
            pt.warmstart_tangent_world = v3at(o + 10); pt.dp1 = v3at(o + 13); pt.dp2 = v3at(o + 16);
        }
        w.pairs.push_back(p);
        Constraint c;
        generate(w, 0, c);
        for (int i = 0; i < 39; ++i) out[i] = 0.0f;
        out[0] = (float)c.num_contacts; put3(1, c.dir1); put3(4, c.tangent1); out[7] = c.limit;
        for (int k = 0; k < MAX_MANIFOLD_POINTS; ++k) {
            out[8 + k] = c.normal[k].impulse; out[12 + k] = c.normal[k].impulse_accumulator; out[16 + k] = c.normal[k].r;
            out[20 + k] = c.b_dist[k]; out[24 + k] = c.twist_dists[k];
            out[35 + k] = k < c.num_contacts ? (float)c.cids[k] : 255.0f;   // u8::MAX marks an inactive slot
        }
        out[28] = c.t_impulse[0]; out[29] = c.t_impulse[1]; out[30] = c.t_impulse_acc[0]; out[31] = c.t_impulse_acc[1];
        out[32] = c.w_impulse; out[33] = c.w_impulse_acc; out[34] = c.w_r;
        return 0;
    }
    if (name == "recentered_angle") {   // in: theta (rotation of frame 2 about X), limit min, limit max -> the re-centred angle
        if (n_in < 3 || n_out < 1) return -3;   // (JointConstraintHelper::new on identity / rotation-about-X frames + AngularLimitParams::new)
        const float theta = in[0], lo = in[1], hi = in[2];
        Q4 q1 = qidentity(), q2 = Q4{sinf(theta * 0.5f), 0.0f, 0.0f, cosf(theta * 0.5f)};
        const float sgn = copysignf(1.0f, qdot(q1, q2));
        Q4 e = qmul(qconj(q1), q2);
        const float half_range = (hi - lo) * 0.5f;
        float c_cos = 1.0f, c_sin = 0.0f;
        if (!(half_range >= 3.14159265358979323846f || half_range != half_range)) { const float c = (lo + hi) * 0.5f; c_cos = cosf(c * 0.5f); c_sin = sinf(c * 0.5f); }
        out[0] = recentered_angle(e.x * sgn, e.w * sgn, c_cos, c_sin);
        return 0;
    }
    if (name == "normal_solve" || name == "tangent_solve") {
        World w;
        w.sb.resize(2);
        Constraint c;
        memset(&c, 0, sizeof(c));
        c.id1 = 0; c.id2 = 1; c.num_contacts = 1;
        int o;
        if (name == "normal_solve") {   // contact_constraint_element.rs:481-504
            if (n_in < 37 || n_out < 13) return -3;
            c.dir1 = v3at(0); c.im1 = v3at(3); c.im2 = v3at(6);
            NormalPart& n = c.normal[0];
            n.torque_dir1 = v3at(9); n.torque_dir2 = v3at(12); n.ii_torque_dir1 = v3at(15); n.ii_torque_dir2 = v3at(18);
            n.r = in[21]; n.rhs = in[22]; n.impulse = in[23]; n.cfm_factor = in[24];
            o = 25;
        } else {                        // contact_constraint_element.rs:650-705 (coupled 2x2 tangent solve, no twist: one point)
            if (n_in < 59 || n_out < 14) return -3;
            c.dir1 = v3at(0); c.tangent1 = v3at(3);   // t2 = dir x t1 must equal the fixture's second tangent
            c.im1 = v3at(9); c.im2 = v3at(12);
            c.t_torque_dir1[0] = v3at(15); c.t_torque_dir1[1] = v3at(18); c.t_torque_dir2[0] = v3at(21); c.t_torque_dir2[1] = v3at(24);
            c.t_ii_torque_dir1[0] = v3at(27); c.t_ii_torque_dir1[1] = v3at(30); c.t_ii_torque_dir2[0] = v3at(33); c.t_ii_torque_dir2[1] = v3at(36);
            c.t_r[0] = in[39]; c.t_r[1] = in[40]; c.t_r[2] = in[41];
            const V3 lfc1 = v3at(42);   // rhs_j = lfc1 . t_j (unit sub-step inverse dt)
            c.t_rhs[0] = dot(lfc1, v3at(3)); c.t_rhs[1] = dot(lfc1, v3at(6));
            c.t_impulse[0] = in[45]; c.t_impulse[1] = in[46];
            c.limit = 0.0f;   // tangent limit = limit * sum(normal impulses); the fixture's limit rides on the normal impulse
            c.normal[0].impulse = in[47]; c.limit = 1.0f;
            c.normal[0].r = 0.0f; c.normal[0].cfm_factor = 1.0f; c.normal[0].rhs = 0.0f;   // normal row inert: r = 0 keeps its impulse
            o = 48;
        }
        w.sb[0].lin = v3at(o); w.sb[0].ang = v3at(o + 3); w.sb[1].lin = v3at(o + 6); w.sb[1].ang = v3at(o + 9);
        w.sb[0].im = c.im1; w.sb[1].im = c.im2;
        w.sb[0].pose = pose_identity(); w.sb[1].pose = pose_identity();
        w.sb[0].ii = sdp_zero(); w.sb[1].ii = sdp_zero();
        if (name == "normal_solve") {
            solve(w, c, false);
            out[0] = c.normal[0].impulse;
            put3(1, w.sb[0].lin); put3(4, w.sb[0].ang); put3(7, w.sb[1].lin); put3(10, w.sb[1].ang);
        } else {
            // the normal row must not move anything: impulse kept by max(impulse - 0 * dvel, 0) * 1
            solve(w, c, true);
            out[0] = c.t_impulse[0]; out[1] = c.t_impulse[1];
            put3(2, w.sb[0].lin); put3(5, w.sb[0].ang); put3(8, w.sb[1].lin); put3(11, w.sb[1].ang);
        }
        return 0;
    }
    return -100;   // not one of this file's
}

}  // namespace orc
