// rb_solver.cuh -- the velocity solver + integrator, fused per work item.
//
// One work item = a group of whole connected components ("islands").  An item is solved start to
// finish by ONE CTA with its bodies' velocities / poses / inverse inertias staged in shared memory
// (solve_item<BlockExec, SmemBodies>), or -- for islands too large for a CTA (item 0) -- by the
// whole grid with bodies in HBM and grid-wide barriers (solve_item<GridExec, GlobalBodies>).
// Either way the stage order is the reference's:
//   S1 solver-body init   staged_island_solver/worker.rs:46-104, solver_body.rs:82-121
//   S2 generate           contact_with_twist_friction.rs:58-424
//   per substep: S3 increments+gyro (worker.rs:235-284), S4 joint rows (joint_constraint_builder.rs:77-152),
//                S5 update+warmstart per colour (contact_with_twist_friction.rs:426-522, :633-678),
//                S6 biased solve: joints then contacts per colour (staged_island_solver/solve.rs:12-209),
//                S7 integrate (worker.rs:568-631), S8 relax solve (contact_with_twist_friction.rs:529-554, :680-781)
//   S9 restitution (worker.rs:657-734)  S10 impulse writeback (:742-802)  S11 body writeback (:809-897)
//   + advance_to_final_positions for the item's bodies (substep.rs:84-224; rigid_body_components.rs:528-572).
// Constraint rows hold only what cannot be recomputed cheaply: lever arms, effective masses,
// builder anchors and the accumulated impulses; jacobians, rhs and cfm are recomputed from the
// staged body state in every sweep (same expressions, hence the same bits, as storing them).
#pragma once
#include "rb_collide.cuh"

namespace rb {

constexpr int SB_STRIDE = 29;  // floats per staged body (odd: conflict-free across bodies)
// staged body layout: lin 0-2, ang 3-5, q 6-9, t 10-12, ii 13-18, im 19-21, incr_lin 22-24, incr_ang 25-27, flags 28

struct BodyState { vec3 lin, ang; pose p; sym3 ii; vec3 im; };

struct SmemBodies {
    float* s;
    RB_HD vec3 lin(int i) const { const float* b = s + i * SB_STRIDE; return mk3(b[0], b[1], b[2]); }
    RB_HD vec3 ang(int i) const { const float* b = s + i * SB_STRIDE; return mk3(b[3], b[4], b[5]); }
    RB_HD void set_vel(int i, vec3 l, vec3 a) const {
        float* b = s + i * SB_STRIDE;
        b[0] = l.x; b[1] = l.y; b[2] = l.z; b[3] = a.x; b[4] = a.y; b[5] = a.z;
    }
    RB_HD pose xf(int i) const {
        const float* b = s + i * SB_STRIDE;
        quat q; q.x = b[6]; q.y = b[7]; q.z = b[8]; q.w = b[9];
        return mkpose(q, mk3(b[10], b[11], b[12]));
    }
    RB_HD void set_xf(int i, const pose& p) const {
        float* b = s + i * SB_STRIDE;
        b[6] = p.q.x; b[7] = p.q.y; b[8] = p.q.z; b[9] = p.q.w; b[10] = p.t.x; b[11] = p.t.y; b[12] = p.t.z;
    }
    RB_HD sym3 ii(int i) const {
        const float* b = s + i * SB_STRIDE;
        sym3 m; m.xx = b[13]; m.xy = b[14]; m.xz = b[15]; m.yy = b[16]; m.yz = b[17]; m.zz = b[18];
        return m;
    }
    RB_HD vec3 im(int i) const { const float* b = s + i * SB_STRIDE; return mk3(b[19], b[20], b[21]); }
    RB_HD void set_mass(int i, const sym3& m, vec3 im_) const {
        float* b = s + i * SB_STRIDE;
        b[13] = m.xx; b[14] = m.xy; b[15] = m.xz; b[16] = m.yy; b[17] = m.yz; b[18] = m.zz;
        b[19] = im_.x; b[20] = im_.y; b[21] = im_.z;
    }
    RB_HD vec3 incr_lin(int i) const { const float* b = s + i * SB_STRIDE; return mk3(b[22], b[23], b[24]); }
    RB_HD vec3 incr_ang(int i) const { const float* b = s + i * SB_STRIDE; return mk3(b[25], b[26], b[27]); }
    RB_HD void set_incr(int i, vec3 l, vec3 a) const {
        float* b = s + i * SB_STRIDE;
        b[22] = l.x; b[23] = l.y; b[24] = l.z; b[25] = a.x; b[26] = a.y; b[27] = a.z;
    }
};

struct GlobalBodies {  // ids are global body indices; inverse masses are read from the body tables
    const World* w;
    RB_HD vec3 lin(int i) const { return xyz(w->s_lin[i]); }
    RB_HD vec3 ang(int i) const { return xyz(w->s_ang[i]); }
    RB_HD void set_vel(int i, vec3 l, vec3 a) const { w->s_lin[i] = f4(l, 0.f); w->s_ang[i] = f4(a, 0.f); }
    RB_HD pose xf(int i) const { return mkpose(mkq(w->s_q[i]), xyz(w->s_t[i])); }
    RB_HD void set_xf(int i, const pose& p) const { w->s_q[i] = f4(p.q); w->s_t[i] = f4(p.t, 0.f); }
    RB_HD sym3 ii(int i) const { return load_ii(*w, i); }
    RB_HD vec3 im(int i) const { return xyz(w->b_eim[i]); }
    RB_HD void set_mass(int, const sym3&, vec3) const {}
    RB_HD vec3 incr_lin(int i) const { return xyz(w->s_incr_lin[i]); }
    RB_HD vec3 incr_ang(int i) const { return xyz(w->s_incr_ang[i]); }
    RB_HD void set_incr(int i, vec3 l, vec3 a) const { w->s_incr_lin[i] = f4(l, 0.f); w->s_incr_ang[i] = f4(a, 0.f); }
};

template <class B>
RB_HD BodyState gather_body(const B& bd, int id) {  // world-attached side: identity / zero (solver_body.rs:11-33)
    BodyState g;
    if (id == NO_BODY) {
        g.lin = zero3(); g.ang = zero3(); g.p = pident(); g.ii = sym_zero(); g.im = zero3();
    } else {
        g.lin = bd.lin(id); g.ang = bd.ang(id); g.p = bd.xf(id); g.ii = bd.ii(id); g.im = bd.im(id);
    }
    return g;
}
template <class B>
RB_HD void scatter_vel(const B& bd, int id, vec3 l, vec3 a) {
    if (id != NO_BODY) bd.set_vel(id, l, a);
}

RB_HD float bouncy(float restitution, bool is_new) {  // contact_pair.rs:773-779
    return is_new ? (restitution > 0.0f ? 1.0f : 0.0f) : (restitution >= 1.0f ? 1.0f : 0.0f);
}

// S2: contact_with_twist_friction.rs:58-424 for the manifold scheduled at slot q.
template <class B>
RB_HD void cons_generate(const World& w, const B& bd, int q, int buf, int item) {
    int4 h = w.cons_hdr[q];
    const int p = h.x, id1 = h.y, id2 = h.z;
    BodyState g1 = gather_body(bd, id1), g2 = gather_body(bd, id2);
    vec3 com1 = g1.p.t, com2 = g2.p.t;
    float4 nrm = prow(w, buf, PR_NORMAL, p);
    float restitution = prow(w, buf, PR_LN2, p).w;
    int count = as_int(prow(w, buf, PR_INFO, p).z);
    if (count > MAX_PTS) count = MAX_PTS;
    vec3 dir = -xyz(nrm);
    vec3 t1 = ortho_vector(dir);
    vec3 t2 = cross3(dir, t1);
    float inv_n = 1.0f / (float)count;
    vec3 fc1 = zero3(), fc2 = zero3();
    float tws = 0.0f, tgs0 = 0.0f, tgs1 = 0.0f;
    vec3 pts[MAX_PTS];
    float imp[MAX_PTS] = {0.f, 0.f, 0.f, 0.f}, acc[MAX_PTS] = {0.f, 0.f, 0.f, 0.f};
    bool any_seed = false;
    for (int k = 0; k < count; ++k) {
        float4 a1 = prow(w, buf, PR_A1 + k, p), a2 = prow(w, buf, PR_A2 + k, p);
        int cid = as_int(a1.w);
        float4 pd = prow(w, buf, PR_PD + cid, p);
        vec3 wt = xyz(prow(w, buf, PR_TW + cid, p));
        vec3 dp1 = xyz(prow(w, buf, PR_DP1 + cid, p)), dp2 = xyz(prow(w, buf, PR_DP2 + cid, p));
        float ws_imp = pd.y, ws_twist = pd.z;
        float w0 = dot3(wt, t1), w1 = dot3(wt, t2);
        float bz = bouncy(restitution, pd.x == 0.0f);
        vec3 p1 = xform(g1.p, xyz(a1));
        vec3 p2 = xform(g2.p, xyz(a2));
        float dist = dot3(p1 - p2, dir);
        vec3 point = com1 + dp1;
        pts[k] = point;
        fc1 = fc1 + point * inv_n;
        fc2 = fc2 + (com2 + dp2) * inv_n;
        vec3 v1 = g1.lin + cross3(g1.ang, dp1);
        vec3 v2 = g2.lin + cross3(g2.ang, dp2);
        tws = tws + ws_twist * inv_n;
        tgs0 = tgs0 + w0 * inv_n;
        tgs1 = tgs1 + w1 * inv_n;
        vec3 td1 = cross3(dp1, dir), td2 = cross3(dp2, -dir);
        vec3 itd1 = smul(g1.ii, td1), itd2 = smul(g2.ii, td2);
        vec3 imsum = g1.im + g2.im;
        float r = safe_inv(dot3(dir, had(imsum, dir)) + dot3(itd1, td1) + dot3(itd2, td2));
        float pv = dot3(v1 - v2, dir);
        float seed = bz * restitution * pv;
        any_seed = any_seed || seed < 0.0f;
        imp[k] = ws_imp;
        acc[k] = -ws_imp;
        crow(w, CR_DP1 + k, q) = f4(dp1, r);
        crow(w, CR_DP2 + k, q) = f4(dp2, dist - dot3(point - (com2 + dp2), dir));
        crow(w, CR_LP1 + k, q) = f4(xform_inv(g1.p, point), seed);
        crow(w, CR_LP2 + k, q) = f4(xform_inv(g2.p, com2 + dp2), as_float_i(cid));
    }
    float wimp = count > 1 ? tws : 0.0f;
    vec3 tdp1 = fc1 - com1, tdp2 = fc2 - com2;
    float twd[MAX_PTS] = {0.f, 0.f, 0.f, 0.f};
    float wr = 0.0f;
    if (count > 1) {
        for (int k = 0; k < count; ++k) twd[k] = norm(fc1 - pts[k]);
        vec3 i1 = smul(g1.ii, dir), i2 = smul(g2.ii, -dir);
        wr = safe_inv(dot3(i1, dir) + dot3(i2, -dir));
    }
    float tr[3];
    vec3 itd1s[2], itd2s[2], td1s[2], td2s[2];
    for (int j = 0; j < 2; ++j) {
        vec3 tj = j == 0 ? t1 : t2;
        vec3 td1 = cross3(tdp1, tj), td2 = cross3(tdp2, -tj);
        vec3 itd1 = smul(g1.ii, td1), itd2 = smul(g2.ii, td2);
        vec3 imsum = g1.im + g2.im;
        tr[j] = dot3(tj, had(imsum, tj)) + dot3(itd1, td1) + dot3(itd2, td2);
        td1s[j] = td1; td2s[j] = td2; itd1s[j] = itd1; itd2s[j] = itd2;
    }
    tr[2] = 2.0f * (dot3(itd1s[0], td1s[1]) + dot3(itd2s[0], td2s[1]));
    crow(w, CR_DIR, q) = f4(dir, nrm.w);
    crow(w, CR_T1, q) = f4(t1, wr);
    crow(w, CR_TDP1, q) = f4(tdp1, tr[0]);
    crow(w, CR_TDP2, q) = f4(tdp2, tr[1]);
    crow(w, CR_LFC1, q) = f4(xform_inv(g1.p, fc1), tr[2]);
    crow(w, CR_LFC2, q) = f4(xform_inv(g2.p, fc2), 0.0f);
    crow(w, CR_TWD, q) = make_float4(twd[0], twd[1], twd[2], twd[3]);
    crow(w, CR_IMP, q) = make_float4(imp[0], imp[1], imp[2], imp[3]);
    crow(w, CR_ACC, q) = make_float4(acc[0], acc[1], acc[2], acc[3]);
    crow(w, CR_TI, q) = make_float4(tgs0, tgs1, -tgs0, -tgs1);
    crow(w, CR_WI, q) = make_float4(wimp, -wimp, 0.0f, 0.0f);
    h.w = count;
    w.cons_hdr[q] = h;
    if (any_seed) w.item_flags[item] = 1;
}

RB_HD float getk(float4 v, int k) { return k == 0 ? v.x : (k == 1 ? v.y : (k == 2 ? v.z : v.w)); }
RB_HD void setk(float4& v, int k, float x) { if (k == 0) v.x = x; else if (k == 1) v.y = x; else if (k == 2) v.z = x; else v.w = x; }

// Sweep modes
constexpr int MODE_WARMSTART = 0, MODE_BIASED = 1, MODE_RELAX = 2, MODE_RESTITUTION = 3;

// One constraint, one sweep.  MODE_WARMSTART = builder.update + constraint.warmstart (fused,
// worker.rs:438-539); MODE_BIASED / MODE_RELAX = (refresh_rhs_wo_bias +) solve.
template <class B>
RB_HD void cons_sweep(const World& w, const B& bd, int q, int mode, bool solve_friction) {
    const Params& P = w.prm;
    int4 h = w.cons_hdr[q];
    const int id1 = h.y, id2 = h.z, nc = h.w;
    BodyState g1 = gather_body(bd, id1), g2 = gather_body(bd, id2);
    vec3 v1 = g1.lin, w1 = g1.ang, v2 = g2.lin, w2 = g2.ang;
    float4 dirl = crow(w, CR_DIR, q), t1w = crow(w, CR_T1, q);
    vec3 dir = xyz(dirl), t1 = xyz(t1w);
    vec3 t2 = cross3(dir, t1);
    float4 imp = crow(w, CR_IMP, q);
    bool is_static = id1 == NO_BODY || id2 == NO_BODY;
    float stf = is_static ? 1.0f : 0.0f;
    float cfm_soft = P.dyn_cfm + stf * (P.static_cfm - P.dyn_cfm);
    float erp = P.dyn_erp + stf * (P.static_erp - P.dyn_erp);
    vec3 lin1 = had(dir, g1.im), lin2 = had(dir, g2.im);

    if (mode == MODE_WARMSTART) {
        float4 acc = crow(w, CR_ACC, q);
        for (int k = 0; k < nc; ++k) {
            float l = getk(imp, k);
            setk(acc, k, getk(acc, k) + l);
            l = l * P.warmstart_coeff;
            setk(imp, k, l);
            vec3 dp1 = xyz(crow(w, CR_DP1 + k, q)), dp2 = xyz(crow(w, CR_DP2 + k, q));
            vec3 itd1 = smul(g1.ii, cross3(dp1, dir)), itd2 = smul(g2.ii, cross3(dp2, -dir));
            v1 = v1 + lin1 * l;
            w1 = w1 + itd1 * l;
            v2 = v2 + lin2 * (-l);
            w2 = w2 + itd2 * l;
        }
        float4 ti = crow(w, CR_TI, q), wi = crow(w, CR_WI, q);
        ti.z = ti.z + ti.x; ti.w = ti.w + ti.y;
        ti.x = ti.x * P.warmstart_coeff; ti.y = ti.y * P.warmstart_coeff;
        wi.y = wi.y + wi.x;
        wi.x = wi.x * P.warmstart_coeff;
        vec3 tdp1 = xyz(crow(w, CR_TDP1, q)), tdp2 = xyz(crow(w, CR_TDP2, q));
        vec3 i10 = smul(g1.ii, cross3(tdp1, t1)), i11 = smul(g1.ii, cross3(tdp1, t2));
        vec3 i20 = smul(g2.ii, cross3(tdp2, -t1)), i21 = smul(g2.ii, cross3(tdp2, -t2));
        v1 = v1 + had(t1 * ti.x + t2 * ti.y, g1.im);
        w1 = w1 + (i10 * ti.x + i11 * ti.y);
        v2 = v2 + had(t1 * (-ti.x) + t2 * (-ti.y), g2.im);
        w2 = w2 + (i20 * ti.x + i21 * ti.y);
        if (nc > 1) {
            w1 = w1 + smul(g1.ii, dir) * wi.x;
            w2 = w2 - smul(g2.ii, dir) * wi.x;
        }
        crow(w, CR_IMP, q) = imp;
        crow(w, CR_ACC, q) = acc;
        crow(w, CR_TI, q) = ti;
        crow(w, CR_WI, q) = wi;
        scatter_vel(bd, id1, v1, w1);
        scatter_vel(bd, id2, v2, w2);
        return;
    }

    if (mode == MODE_RESTITUTION) {  // contact_constraint_element.rs:508-534
        float4 acc = crow(w, CR_ACC, q);
        bool any = false;
        for (int k = 0; k < nc; ++k) any = any || crow(w, CR_LP1 + k, q).w < 0.0f;
        if (!any) return;
        for (int k = 0; k < nc; ++k) {
            float4 d1r = crow(w, CR_DP1 + k, q);
            vec3 dp1 = xyz(d1r), dp2 = xyz(crow(w, CR_DP2 + k, q));
            float seed = crow(w, CR_LP1 + k, q).w;
            vec3 td1 = cross3(dp1, dir), td2 = cross3(dp2, -dir);
            vec3 itd1 = smul(g1.ii, td1), itd2 = smul(g2.ii, td2);
            float l = getk(imp, k);
            float dvel = dot3(dir, v1) + dot3(td1, w1) - dot3(dir, v2) + dot3(td2, w2) + seed;
            bool gate = seed < 0.0f && (getk(acc, k) + l) > 0.0f;
            float nl = max2(l - d1r.w * dvel, 0.0f);
            if (!gate) nl = l;
            float dl = nl - l;
            setk(imp, k, nl);
            v1 = v1 + lin1 * dl;
            w1 = w1 + itd1 * dl;
            v2 = v2 + lin2 * (-dl);
            w2 = w2 + itd2 * dl;
        }
        crow(w, CR_IMP, q) = imp;
        scatter_vel(bd, id1, v1, w1);
        scatter_vel(bd, id2, v2, w2);
        return;
    }

    // normal rows (contact_constraint_element.rs:481-504); rhs / cfm recomputed from the current poses
    const bool relax = mode == MODE_RELAX;
    for (int k = 0; k < nc; ++k) {
        float4 d1r = crow(w, CR_DP1 + k, q), d2r = crow(w, CR_DP2 + k, q);
        vec3 dp1 = xyz(d1r), dp2 = xyz(d2r);
        vec3 p1 = xform(g1.p, xyz(crow(w, CR_LP1 + k, q)));
        vec3 p2 = xform(g2.p, xyz(crow(w, CR_LP2 + k, q)));
        float dist = d2r.w + dot3(p1 - p2, dir);
        float rhs = max2(dist, 0.0f) * P.sub_inv_dt;
        float cfm = 1.0f;
        if (!relax) {
            rhs = rhs + clampf(dist * erp, -P.max_corrective_velocity, 0.0f);
            cfm = dist <= 0.0f ? cfm_soft : 1.0f;
        }
        vec3 td1 = cross3(dp1, dir), td2 = cross3(dp2, -dir);
        vec3 itd1 = smul(g1.ii, td1), itd2 = smul(g2.ii, td2);
        float l = getk(imp, k);
        float dvel = dot3(dir, v1) + dot3(td1, w1) - dot3(dir, v2) + dot3(td2, w2) + rhs;
        float nl = cfm * max2(l - d1r.w * dvel, 0.0f);
        float dl = nl - l;
        setk(imp, k, nl);
        v1 = v1 + lin1 * dl;
        w1 = w1 + itd1 * dl;
        v2 = v2 + lin2 * (-dl);
        w2 = w2 + itd2 * dl;
    }
    crow(w, CR_IMP, q) = imp;

    if (solve_friction) {
        float4 ti = crow(w, CR_TI, q), wi = crow(w, CR_WI, q), twd = crow(w, CR_TWD, q);
        float tlimit = 0.0f, wlimit = 0.0f;
        for (int k = 0; k < nc; ++k) {
            tlimit = tlimit + getk(imp, k);
            wlimit = wlimit + getk(imp, k) * getk(twd, k);
        }
        tlimit = tlimit * dirl.w;
        wlimit = wlimit * dirl.w;
        if (nc > 1) {  // twist first (contact_constraint_element.rs:735-756)
            vec3 i1 = smul(g1.ii, dir), i2 = smul(g2.ii, dir);
            float dvel = dot3(dir, w1 - w2) + 0.0f;
            float nl = clampf(wi.x - t1w.w * dvel, -wlimit, wlimit);
            float dl = nl - wi.x;
            wi.x = nl;
            w1 = w1 + i1 * dl;
            w2 = w2 - i2 * dl;
        }
        float4 tdp1r = crow(w, CR_TDP1, q), tdp2r = crow(w, CR_TDP2, q);
        vec3 tdp1 = xyz(tdp1r), tdp2 = xyz(tdp2r);
        vec3 td10 = cross3(tdp1, t1), td11 = cross3(tdp1, t2);
        vec3 td20 = cross3(tdp2, -t1), td21 = cross3(tdp2, -t2);
        vec3 i10 = smul(g1.ii, td10), i11 = smul(g1.ii, td11), i20 = smul(g2.ii, td20), i21 = smul(g2.ii, td21);
        float rhs0 = 0.0f, rhs1 = 0.0f;  // tangent rhs_wo_bias = tangent_velocity . t = 0 (no hooks)
        if (!relax) {  // update(): bias from the friction-centre drift (contact_with_twist_friction.rs:506-514)
            vec3 p1 = xform(g1.p, xyz(crow(w, CR_LFC1, q)));
            vec3 p2 = xform(g2.p, xyz(crow(w, CR_LFC2, q)));
            rhs0 = 0.0f + dot3(p1 - p2, t1) * P.sub_inv_dt;
            rhs1 = 0.0f + dot3(p1 - p2, t2) * P.sub_inv_dt;
        }
        float dv0 = dot3(t1, v1) + dot3(td10, w1) - dot3(t1, v2) + dot3(td20, w2) + rhs0;
        float dv1 = dot3(t2, v1) + dot3(td11, w1) - dot3(t2, v2) + dot3(td21, w2) + rhs1;
        float k11 = tdp1r.w, k22 = tdp2r.w, k12 = crow(w, CR_LFC1, q).w * 0.5f;
        float inv_det = safe_inv(k11 * k22 - k12 * k12);
        float d0 = (k22 * dv0 - k12 * dv1) * inv_det;
        float d1 = (k11 * dv1 - k12 * dv0) * inv_det;
        float n0 = ti.x - d0, n1 = ti.y - d1;
        float len = sqrtf(n0 * n0 + n1 * n1);
        if (len > tlimit) {
            float s = tlimit / len;
            n0 = n0 * s;
            n1 = n1 * s;
        }
        float dl0 = n0 - ti.x, dl1 = n1 - ti.y;
        ti.x = n0;
        ti.y = n1;
        v1 = v1 + had(t1 * dl0 + t2 * dl1, g1.im);
        w1 = w1 + (i10 * dl0 + i11 * dl1);
        v2 = v2 + had(t1 * (-dl0) + t2 * (-dl1), g2.im);
        w2 = w2 + (i20 * dl0 + i21 * dl1);
        crow(w, CR_TI, q) = ti;
        crow(w, CR_WI, q) = wi;
    }
    scatter_vel(bd, id1, v1, w1);
    scatter_vel(bd, id2, v2, w2);
}

RB_HD float canon0(float x) { return x == 0.0f ? 0.0f : x; }

// S10: contact_with_twist_friction.rs:783-829
RB_HD void cons_writeback(const World& w, int q, int buf) {
    int4 h = w.cons_hdr[q];
    const int p = h.x, nc = h.w;
    float4 dirl = crow(w, CR_DIR, q);
    vec3 dir = xyz(dirl), t1 = xyz(crow(w, CR_T1, q));
    vec3 t2 = cross3(dir, t1);
    float4 imp = crow(w, CR_IMP, q), acc = crow(w, CR_ACC, q), ti = crow(w, CR_TI, q), wi = crow(w, CR_WI, q);
    float a0 = canon0(ti.x), a1 = canon0(ti.y);
    vec3 tw = t1 * a0 + t2 * a1;
    tw = mk3(canon0(tw.x), canon0(tw.y), canon0(tw.z));
    float twist = canon0(wi.x);
    for (int k = 0; k < nc; ++k) {
        int cid = as_int(crow(w, CR_LP2 + k, q).w);
        float4 pd = prow(w, buf, PR_PD + cid, p);
        pd.y = canon0(getk(imp, k));
        pd.x = canon0(getk(acc, k) + getk(imp, k));
        pd.z = twist;
        prow(w, buf, PR_PD + cid, p) = pd;
        prow(w, buf, PR_TW + cid, p) = f4(tw, 0.0f);
    }
}

// rigid_body.rs:2023-2046
RB_HD vec3 gyro_corrected(vec3 angvel, quat axes, vec3 pin, vec3 ipin, float dt) {
    vec3 wl = rotate_inv(axes, angvel);
    vec3 cur = had(pin, wl);
    vec3 eg = (-cross3(wl, cur)) * dt;
    vec3 tot = cur + eg;
    float ts = norm2(tot);
    if (ts != 0.0f) {
        vec3 capped = tot * sqrtf(norm2(cur) / ts);
        return rotate(axes, had(ipin, capped));
    }
    return angvel;
}

// ------------------------------------------------------------------------------------------------
// Joints (locked axes): per-substep rows (joint_constraint_helper.rs:95-164, :411-461, :628-722).
// ------------------------------------------------------------------------------------------------
template <class B>
RB_HD void joint_update(const World& w, const B& bd, int q) {
    int4 h = w.j_sched_ids[q];
    const int j = h.x, id1 = h.y, id2 = h.z;
    int4 ji = w.j_info[j];
    unsigned locked = (unsigned)ji.z;
    BodyState g1 = gather_body(bd, id1), g2 = gather_body(bd, id2);
    // transform_to_solver_body_space (generic_joint.rs:624-636)
    pose lf1 = mkpose(mkq(w.j_f1_q[j]), xyz(w.j_f1_t[j]));
    pose lf2 = mkpose(mkq(w.j_f2_q[j]), xyz(w.j_f2_t[j]));
    if (id1 == NO_BODY) lf1 = pmul(body_pose(w, ji.x), lf1); else lf1.t = lf1.t - xyz(w.b_lcom_im[ji.x]);
    if (id2 == NO_BODY) lf2 = pmul(body_pose(w, ji.y), lf2); else lf2.t = lf2.t - xyz(w.b_lcom_im[ji.y]);
    pose f1 = pmul(g1.p, lf1), f2 = pmul(g2.p, lf2);
    float2 soft = w.j_soft[j];
    float omega = soft.x * 6.283185307179586f;
    float sdt = w.prm.sub_dt;
    float erp_inv_dt = omega / (sdt * omega + 2.0f * soft.y);
    float erpv = sdt * erp_inv_dt;
    float cfm_coeff = 0.0f;
    if (erpv != 0.0f) {
        float e1 = 1.0f / erpv - 1.0f;
        cfm_coeff = e1 * e1 / ((1.0f + e1) * 4.0f * soft.y * soft.y);
    }
    mat3 basis = rotmat(f1.q);
    vec3 bc[3] = {basis.c0, basis.c1, basis.c2};
    vec3 lin_err = f2.t - f1.t;
    vec3 nc1 = f2.t;
    for (int i = 0; i < 3; ++i)
        if (locked & (1u << i)) nc1 = nc1 - bc[i] * dot3(lin_err, bc[i]);
    f1.t = nc1;
    vec3 r1 = f1.t - g1.p.t, r2 = f2.t - g2.p.t;
    float sgn = copysign1(qdot(f1.q, f2.q));
    quat ae = qmul(qconj(f1.q), f2.q);
    vec3 aerr = mk3(ae.x * sgn, ae.y * sgn, ae.z * sgn);
    vec3 a = mk3(f1.q.x, f1.q.y, f1.q.z), b = mk3(f2.q.x, f2.q.y, f2.q.z);
    float wa = f1.q.w, wb = f2.q.w;
    vec3 cv = a * wb + b * wa;
    float ab = dot3(a, b);
    vec3 imsum = g1.im + g2.im;
    vec3 lin[6], aj1[6], aj2[6], ia1[6], ia2[6];
    float rhs[6], rwb[6], cg[6], il[6];
    int dof[6] = {0, 0, 0, 0, 0, 0};
    int len = 0;
    for (int i = 3; i < 6; ++i) {
        if (!(locked & (1u << i))) continue;
        int ax = i - 3;
        // row `ax` of D = 0.5 (a b^T + (wa wb - a.b) I - [cv]x + b a^T)  (rotation_ops.rs:121-137), times sgn
        float av = comp(a, ax), bv = comp(b, ax);
        float dg = wa * wb - ab;
        vec3 cx = ax == 0 ? mk3(0.0f, -cv.z, cv.y) : (ax == 1 ? mk3(cv.z, 0.0f, -cv.x) : mk3(-cv.y, cv.x, 0.0f));
        vec3 row = mk3((av * b.x + (ax == 0 ? dg : 0.0f) - cx.x + bv * a.x) * 0.5f,
                       (av * b.y + (ax == 1 ? dg : 0.0f) - cx.y + bv * a.y) * 0.5f,
                       (av * b.z + (ax == 2 ? dg : 0.0f) - cx.z + bv * a.z) * 0.5f);
        vec3 aj = row * sgn;
        lin[len] = zero3(); aj1[len] = aj; aj2[len] = aj;
        ia1[len] = smul(g1.ii, aj); ia2[len] = smul(g2.ii, aj);
        rwb[len] = 0.0f; rhs[len] = 0.0f + comp(aerr, ax) * erp_inv_dt; cg[len] = 0.0f; il[len] = 0.0f; dof[len] = i;
        ++len;
    }
    for (int i = 0; i < 3; ++i) {
        if (!(locked & (1u << i))) continue;
        lin[len] = bc[i]; aj1[len] = cross3(r1, bc[i]); aj2[len] = cross3(r2, bc[i]);
        ia1[len] = smul(g1.ii, aj1[len]); ia2[len] = smul(g2.ii, aj2[len]);
        rwb[len] = 0.0f; rhs[len] = 0.0f + dot3(bc[i], lin_err) * erp_inv_dt; cg[len] = 0.0f; il[len] = 0.0f; dof[len] = i;
        ++len;
    }
    for (int jx = 0; jx < len; ++jx) {  // finalize_constraints: Gram-Schmidt in the mass metric
        float djj = dot3(lin[jx], had(imsum, lin[jx])) + dot3(ia1[jx], aj1[jx]) + dot3(ia2[jx], aj2[jx]);
        float gain = djj * cfm_coeff + cg[jx];
        float inv_djj = safe_inv(djj);
        il[jx] = safe_inv(djj + gain);
        cg[jx] = gain;
        for (int ix = jx + 1; ix < len; ++ix) {
            float dij = dot3(lin[ix], had(imsum, lin[jx])) + dot3(ia1[ix], aj1[jx]) + dot3(ia2[ix], aj2[jx]);
            float coeff = dij * inv_djj;
            lin[ix] = lin[ix] - lin[jx] * coeff;
            aj1[ix] = aj1[ix] - aj1[jx] * coeff;
            aj2[ix] = aj2[ix] - aj2[jx] * coeff;
            ia1[ix] = ia1[ix] - ia1[jx] * coeff;
            ia2[ix] = ia2[ix] - ia2[jx] * coeff;
            rwb[ix] = rwb[ix] - rwb[jx] * coeff;
            rhs[ix] = rhs[ix] - rhs[jx] * coeff;
        }
    }
    for (int r = 0; r < len; ++r) {
        int s = 6 * q + r;
        jrow(w, JR_LIN, s) = f4(lin[r], 0.0f);   // impulse restarts from 0 (warmstart_joints = false)
        jrow(w, JR_A1, s) = f4(aj1[r], il[r]);
        jrow(w, JR_A2, s) = f4(aj2[r], rhs[r]);
        jrow(w, JR_IA1, s) = f4(ia1[r], rwb[r]);
        jrow(w, JR_IA2, s) = f4(ia2[r], cg[r]);
    }
    h.w = len | (dof[0] << 8) | (len > 1 ? dof[1] << 12 : 0) | (len > 2 ? dof[2] << 16 : 0) | (len > 3 ? dof[3] << 20 : 0) |
          (len > 4 ? dof[4] << 24 : 0) | (len > 5 ? dof[5] << 28 : 0);
    w.j_sched_ids[q] = h;
}

// joint_velocity_constraint.rs:97-124
template <class B>
RB_HD void joint_solve(const World& w, const B& bd, int q, bool wo_bias) {
    int4 h = w.j_sched_ids[q];
    const int id1 = h.y, id2 = h.z, len = h.w & 0xff;
    BodyState g1 = gather_body(bd, id1), g2 = gather_body(bd, id2);
    vec3 v1 = g1.lin, w1 = g1.ang, v2 = g2.lin, w2 = g2.ang;
    for (int r = 0; r < len; ++r) {
        int s = 6 * q + r;
        float4 L = jrow(w, JR_LIN, s), A1 = jrow(w, JR_A1, s), A2 = jrow(w, JR_A2, s), I1 = jrow(w, JR_IA1, s),
               I2 = jrow(w, JR_IA2, s);
        float rhs_c = wo_bias ? I1.w : A2.w;
        float dlin = dot3(xyz(L), v2 - v1);
        float dang = dot3(xyz(A2), w2) - dot3(xyz(A1), w1);
        float rhs = dlin + dang + rhs_c;
        float total = L.w + A1.w * (rhs - I2.w * L.w);
        float delta = total - L.w;
        L.w = total;
        vec3 li = xyz(L) * delta;
        v1 = v1 + had(li, g1.im);
        w1 = w1 + xyz(I1) * delta;
        v2 = v2 - had(li, g2.im);
        w2 = w2 - xyz(I2) * delta;
        jrow(w, JR_LIN, s) = L;
        if (wo_bias) { A2.w = I1.w; jrow(w, JR_A2, s) = A2; }
    }
    scatter_vel(bd, id1, v1, w1);
    scatter_vel(bd, id2, v2, w2);
}

RB_HD void joint_writeback(const World& w, int q) {
    int4 h = w.j_sched_ids[q];
    int len = h.w & 0xff;
    for (int r = 0; r < len; ++r) {
        int dof = (h.w >> (8 + 4 * r)) & 0xf;
        w.j_impulses[h.x * 6 + dof] = jrow(w, JR_LIN, 6 * q + r).w;
    }
}

// rigid_body_components.rs:528-572 for body b at pose p; writes world_com / effective masses.
RB_HD void update_world_mass(const World& w, int b, const pose& p) {
    float4 lc = w.b_lcom_im[b];
    bool dyn = w.b_type[b] == BODY_DYNAMIC;
    unsigned fl = w.b_flags[b];
    w.b_wcom[b] = f4(xform(p, xyz(lc)), 0.0f);
    vec3 im = mk3(lc.w, lc.w, lc.w);
    vec3 d = xyz(w.b_ipi[b]);
    sym3 m = sym_zero();
    if (d.x != 0.0f || d.y != 0.0f || d.z != 0.0f) {
        mat3 r = rotmat(qmul(p.q, mkq(w.b_pframe[b])));
        m.xx = (r.c0.x * d.x) * r.c0.x + (r.c1.x * d.y) * r.c1.x + (r.c2.x * d.z) * r.c2.x;
        m.xy = (r.c0.x * d.x) * r.c0.y + (r.c1.x * d.y) * r.c1.y + (r.c2.x * d.z) * r.c2.y;
        m.xz = (r.c0.x * d.x) * r.c0.z + (r.c1.x * d.y) * r.c1.z + (r.c2.x * d.z) * r.c2.z;
        m.yy = (r.c0.y * d.x) * r.c0.y + (r.c1.y * d.y) * r.c1.y + (r.c2.y * d.z) * r.c2.y;
        m.yz = (r.c0.y * d.x) * r.c0.z + (r.c1.y * d.y) * r.c1.z + (r.c2.y * d.z) * r.c2.z;
        m.zz = (r.c0.z * d.x) * r.c0.z + (r.c1.z * d.y) * r.c1.z + (r.c2.z * d.z) * r.c2.z;
    }
    if (!dyn || (fl & FLAG_LTX)) im.x = 0.0f;
    if (!dyn || (fl & FLAG_LTY)) im.y = 0.0f;
    if (!dyn || (fl & FLAG_LTZ)) im.z = 0.0f;
    if (!dyn || (fl & FLAG_LRX)) { m.xx = 0.0f; m.xy = 0.0f; m.xz = 0.0f; }
    if (!dyn || (fl & FLAG_LRY)) { m.yy = 0.0f; m.xy = 0.0f; m.yz = 0.0f; }
    if (!dyn || (fl & FLAG_LRZ)) { m.zz = 0.0f; m.xz = 0.0f; m.yz = 0.0f; }
    w.b_eim[b] = f4(im, 0.0f);
    w.b_eii0[b] = make_float4(m.xx, m.xy, m.xz, m.yy);
    w.b_eii1[b] = make_float2(m.yz, m.zz);
}

// Executors: how the threads of an item iterate and synchronise.
struct BlockExec {
    const BlockCtx* c;
    RB_HD int tid() const { return c->btid; }
    RB_HD int nth() const { return c->bsize; }
    RB_HD void sync() const { c->block_sync(); }
};
struct GridExec {
    const GridCtx* c;
    RB_HD int tid() const { return c->gtid; }
    RB_HD int nth() const { return c->gsize; }
    RB_HD void sync() const { c->grid_sync(); }
};

// Solve one work item from solver-body init to the final positions of its bodies.
template <class X, class B>
RB_PHASE void solve_item(const X& ex, const World& w, const B& bd, int item, vec3 gravity) {
    const Params& P = w.prm;
    State* st = w.st;
    const int buf = st->cur;
    const bool global_ids = item == 0;
    const int b0 = w.item_body_start[item], b1 = w.item_body_start[item + 1];
    const int c0 = w.item_cons_start[item];
    const int c1 = w.item_cons_start[item + 1] < w.cons_cap ? w.item_cons_start[item + 1] : w.cons_cap;
    const int j0 = w.item_joint_start[item], j1 = w.item_joint_start[item + 1];
    const int* coff = w.item_color_off + (size_t)item * (NUM_COLORS + 1);
    const int* joff = w.item_jcolor_off + (size_t)item * (NUM_COLORS + 1);
    const int ncol = st->nused_colors, njcol = st->njused_colors;
    const int ovf = w.color_pos[COLOR_OVERFLOW], jovf = w.jcolor_pos[COLOR_OVERFLOW];
    const int tid = ex.tid(), nth = ex.nth();

    if (tid == 0) w.item_flags[item] = 0;   // bit 0: some contact of this item holds a restitution seed
    // a7 + S1: forces, solver bodies, per-substep increments
    for (int l = b0 + tid; l < b1; l += nth) {
        int b = w.item_bodies[l];
        int id = global_ids ? b : l - b0;
        float4 lc = w.b_lcom_im[b];
        vec3 eim = xyz(w.b_eim[b]);
        sym3 eii = load_ii(w, b);
        float4 misc = w.b_misc[b];
        vec3 emass = mk3(inv_exact0(eim.x), inv_exact0(eim.y), inv_exact0(eim.z));
        vec3 force = xyz(w.b_uforce[b]) + had(gravity, emass) * misc.z;
        vec3 torque = xyz(w.b_utorque[b]);
        bd.set_vel(id, xyz(w.b_linvel[b]), xyz(w.b_angvel[b]));
        bd.set_xf(id, prepend_translation(body_pose(w, b), xyz(lc)));
        bd.set_mass(id, eii, eim);
        bd.set_incr(id, had(force, eim) * P.sub_dt, smul(eii, torque) * P.sub_dt);
    }
    ex.sync();
    // S2 generate
    for (int q = c0 + tid; q < c1; q += nth) cons_generate(w, bd, q, buf, item);
    ex.sync();

    for (int sub = 0; sub < P.num_substeps; ++sub) {
        // S3 increments + gyroscopic correction
        for (int l = b0 + tid; l < b1; l += nth) {
            int b = w.item_bodies[l];
            int id = global_ids ? b : l - b0;
            vec3 lin = bd.lin(id) + bd.incr_lin(id);
            vec3 ang = bd.ang(id) + bd.incr_ang(id);
            if (w.b_flags[b] & FLAG_GYRO) {
                quat axes = qmul(bd.xf(id).q, mkq(w.b_pframe[b]));
                ang = gyro_corrected(ang, axes, xyz(w.b_pi[b]), xyz(w.b_ipi[b]), P.sub_dt);
            }
            bd.set_vel(id, lin, ang);
        }
        ex.sync();
        // S4 joint rows from the current poses
        if (j1 > j0) {
            for (int q = j0 + tid; q < j1; q += nth) joint_update(w, bd, q);
            ex.sync();
        }
        // S5 update + warmstart, colour by colour
        if (P.warmstart_coeff != 0.0f) {
            for (int c = 0; c < ncol; ++c) {
                int a = c0 + coff[c], e = c0 + coff[c + 1];
                if (e > c1) e = c1;
                if (a >= e) continue;
                if (c == ovf) {
                    if (tid == 0) for (int q = a; q < e; ++q) cons_sweep(w, bd, q, MODE_WARMSTART, false);
                } else {
                    for (int q = a + tid; q < e; q += nth) cons_sweep(w, bd, q, MODE_WARMSTART, false);
                }
                ex.sync();
            }
        } else {
            // warmstart_coefficient == 0: update only banks and zeroes the impulses (no velocity change)
            for (int q = c0 + tid; q < c1; q += nth) {
                float4 imp = crow(w, CR_IMP, q), acc = crow(w, CR_ACC, q), ti = crow(w, CR_TI, q), wi = crow(w, CR_WI, q);
                acc.x = acc.x + imp.x; acc.y = acc.y + imp.y; acc.z = acc.z + imp.z; acc.w = acc.w + imp.w;
                imp.x = imp.x * 0.0f; imp.y = imp.y * 0.0f; imp.z = imp.z * 0.0f; imp.w = imp.w * 0.0f;
                ti.z = ti.z + ti.x; ti.w = ti.w + ti.y; ti.x = ti.x * 0.0f; ti.y = ti.y * 0.0f;
                wi.y = wi.y + wi.x; wi.x = wi.x * 0.0f;
                crow(w, CR_IMP, q) = imp; crow(w, CR_ACC, q) = acc; crow(w, CR_TI, q) = ti; crow(w, CR_WI, q) = wi;
            }
            ex.sync();
        }
        for (int pass = 0; pass < 2; ++pass) {
            const bool relax = pass == 1;
            const int iters = relax ? P.num_relax : P.num_pgs;
            const bool fric = relax || P.friction_in_bias || P.num_relax == 0;
            for (int it = 0; it < iters; ++it) {
                // joints first (solve.rs:89-92), then contacts
                for (int c = 0; c < njcol; ++c) {
                    int a = j0 + joff[c], e = j0 + joff[c + 1];
                    if (a >= e) continue;
                    if (c == jovf) {
                        if (tid == 0) for (int q = a; q < e; ++q) joint_solve(w, bd, q, relax);
                    } else {
                        for (int q = a + tid; q < e; q += nth) joint_solve(w, bd, q, relax);
                    }
                    ex.sync();
                }
                for (int c = 0; c < ncol; ++c) {
                    int a = c0 + coff[c], e = c0 + coff[c + 1];
                    if (e > c1) e = c1;
                    if (a >= e) continue;
                    if (c == ovf) {
                        if (tid == 0) for (int q = a; q < e; ++q) cons_sweep(w, bd, q, relax ? MODE_RELAX : MODE_BIASED, fric);
                    } else {
                        for (int q = a + tid; q < e; q += nth) cons_sweep(w, bd, q, relax ? MODE_RELAX : MODE_BIASED, fric);
                    }
                    ex.sync();
                }
            }
            if (!relax) {
                // S7 integrate (speed caps + linearised quaternion update)
                for (int l = b0 + tid; l < b1; l += nth) {
                    int b = w.item_bodies[l];
                    int id = global_ids ? b : l - b0;
                    vec3 lin = bd.lin(id), ang = bd.ang(id);
                    if (P.max_lin_vel != FMAX32) {
                        float n = norm(lin);
                        if (n > P.max_lin_vel) lin = lin * (P.max_lin_vel / n);
                    }
                    if (!(w.b_flags[b] & FLAG_FAST_ROT)) {
                        float n = norm(ang);
                        if (n > P.max_ang_vel) ang = ang * (P.max_ang_vel / n);
                    }
                    bd.set_vel(id, lin, ang);
                    pose p = bd.xf(id);
                    vec3 hang = ang * (P.sub_dt * 0.5f);
                    quat dq; dq.x = hang.x; dq.y = hang.y; dq.z = hang.z; dq.w = 1.0f;
                    p.q = qnormalize(qmul(dq, p.q));
                    p.t = p.t + lin * P.sub_dt;
                    bd.set_xf(id, p);
                }
                ex.sync();
            }
        }
    }
    // S9 restitution
    if (w.item_flags[item]) {
        for (int c = 0; c < ncol; ++c) {
            int a = c0 + coff[c], e = c0 + coff[c + 1];
            if (e > c1) e = c1;
            if (a >= e) continue;
            if (c == ovf) {
                if (tid == 0) for (int q = a; q < e; ++q) cons_sweep(w, bd, q, MODE_RESTITUTION, false);
            } else {
                for (int q = a + tid; q < e; q += nth) cons_sweep(w, bd, q, MODE_RESTITUTION, false);
            }
            ex.sync();
        }
    }
    // S10 impulse writeback
    for (int q = c0 + tid; q < c1; q += nth) cons_writeback(w, q, buf);
    for (int q = j0 + tid; q < j1; q += nth) joint_writeback(w, q);
    // S11 body writeback + advance_to_final_positions
    for (int l = b0 + tid; l < b1; l += nth) {
        int b = w.item_bodies[l];
        int id = global_ids ? b : l - b0;
        float4 misc = w.b_misc[b];
        vec3 lin = bd.lin(id) * (1.0f / (1.0f + P.dt * misc.x));
        vec3 ang = bd.ang(id) * (1.0f / (1.0f + P.dt * misc.y));
        pose np = prepend_translation(bd.xf(id), -xyz(w.b_lcom_im[b]));
        w.b_linvel[b] = f4(lin, 0.0f);
        w.b_angvel[b] = f4(ang, 0.0f);
        w.b_pos_t[b] = f4(np.t, 0.0f);
        w.b_pos_q[b] = f4(np.q);
        update_world_mass(w, b, np);
        float* s = w.state13 + (size_t)b * 13;
        s[0] = np.t.x; s[1] = np.t.y; s[2] = np.t.z; s[3] = np.q.x; s[4] = np.q.y; s[5] = np.q.z; s[6] = np.q.w;
        s[7] = lin.x; s[8] = lin.y; s[9] = lin.z; s[10] = ang.x; s[11] = ang.y; s[12] = ang.z;
    }
}

}  // namespace rb
