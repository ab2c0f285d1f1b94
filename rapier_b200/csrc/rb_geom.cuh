// rb_geom.cuh -- device contact-manifold generation for cuboid / ball pairs and shape AABBs.
//
// Replaces the reference's call into parry3d 0.30.2 at
// src/geometry/narrow_phase/pair_update.rs:323-330 (`query_dispatcher.contact_manifolds`) and
// src/geometry/collider.rs:553-556 (`Shape::compute_aabb`) for the shapes in scope.  One thread
// computes one manifold: SAT over 6 face axes + 9 edge axes, reference/incident face selection,
// then quad/quad clipping in the plane orthogonal to the separating axis (vertex-in-face tests +
// edge crossings), at most 8 points.  Ball pairs reduce to closed forms.
#pragma once
#include "rb_world.cuh"

namespace rb {

struct RawPt { vec3 p1, p2; float dist; uint32_t fid1, fid2; };
struct RawManifold { int n; RawPt pt[MAX_RAW]; vec3 n1, n2; };

constexpr float EPS32 = 1.1920929e-7f;
constexpr float FMAX32 = 3.4028235e38f;
#define RB_INF (as_float(0x7f800000u))
constexpr uint32_t FID_VERTEX = 0x10000000u, FID_FACE = 0x20000000u, FID_EDGE = 0x30000000u;

RB_HD vec3 box_support(vec3 he, vec3 d) { return mk3(copysignf(he.x, d.x), copysignf(he.y, d.y), copysignf(he.z, d.z)); }

RB_HD void sat_faces(vec3 he1, vec3 he2, const pose& p12, float& best, vec3& dir) {
    best = -FMAX32;
    dir = zero3();
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        float sign = copysign1(comp(p12.t, i));
        vec3 axis1 = with_comp(zero3(), i, sign);
        vec3 axis2 = rotate_inv(p12.q, -axis1);
        vec3 lp2 = box_support(he2, axis2);
        vec3 pt2 = xform(p12, lp2);
        float sep = comp(pt2, i) * sign - comp(he1, i);
        if (sep > best) { best = sep; dir = axis1; }
    }
}

RB_HD void sat_edges(vec3 he1, vec3 he2, const pose& p12, float& best, vec3& dir) {
    vec3 e2[3];
    e2[0] = rotate(p12.q, mk3(1.f, 0.f, 0.f));
    e2[1] = rotate(p12.q, mk3(0.f, 1.f, 0.f));
    e2[2] = rotate(p12.q, mk3(0.f, 0.f, 1.f));
    best = -FMAX32;
    dir = zero3();
    for (int k = 0; k < 9; ++k) {
        vec3 u = e2[k / 3];
        int a = k % 3;
        vec3 ax = a == 0 ? mk3(0.0f, -u.z, u.y) : (a == 1 ? mk3(u.z, 0.0f, -u.x) : mk3(-u.y, u.x, 0.0f));
        float n = norm(ax);
        if (!(n > EPS32)) continue;
        vec3 axis1 = ax * (1.0f / n);
        float sg = copysign1(dot3(p12.t, axis1));
        axis1 = axis1 * sg;
        vec3 axis2 = rotate_inv(p12.q, -axis1);
        vec3 lp1 = box_support(he1, axis1);
        vec3 lp2 = box_support(he2, axis2);
        vec3 pt2 = xform(p12, lp2);
        float sep = dot3(pt2 - lp1, axis1);
        if (sep > best) { best = sep; dir = axis1; }
    }
}

struct QuadFace { vec3 v[4]; uint32_t vid[4], eid[4], fid; };

RB_HD uint32_t box_vertex_id(vec3 p) {
    return FID_VERTEX | ((p.x < 0.0f) ? 1u : 0u) | ((p.y < 0.0f) ? 2u : 0u) | ((p.z < 0.0f) ? 4u : 0u);
}

RB_HD void box_support_face(vec3 he, vec3 d, QuadFace& f) {
    float ax = fabsf(d.x), ay = fabsf(d.y), az = fabsf(d.z);
    int im = 0;
    float am = ax;
    if (ay > am) { im = 1; am = ay; }
    if (az > am) { im = 2; am = az; }
    float s = copysign1(comp(d, im));
    if (im == 0) {
        f.v[0] = mk3(he.x * s, he.y, he.z); f.v[1] = mk3(he.x * s, -he.y, he.z);
        f.v[2] = mk3(he.x * s, -he.y, -he.z); f.v[3] = mk3(he.x * s, he.y, -he.z);
    } else if (im == 1) {
        f.v[0] = mk3(he.x, he.y * s, he.z); f.v[1] = mk3(-he.x, he.y * s, he.z);
        f.v[2] = mk3(-he.x, he.y * s, -he.z); f.v[3] = mk3(he.x, he.y * s, -he.z);
    } else {
        f.v[0] = mk3(he.x, he.y, he.z * s); f.v[1] = mk3(he.x, -he.y, he.z * s);
        f.v[2] = mk3(-he.x, -he.y, he.z * s); f.v[3] = mk3(-he.x, he.y, he.z * s);
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) f.vid[i] = box_vertex_id(f.v[i]);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        uint32_t a = f.vid[i] & 7u, b = f.vid[(i + 1) & 3] & 7u;
        uint32_t lo = a < b ? a : b, hi = a < b ? b : a;
        f.eid[i] = FID_EDGE | (lo << 4) | hi;
    }
    f.fid = FID_FACE | (uint32_t)(im + (s < 0.0f ? 3 : 0));
}

struct pt2 { float x, y; };
RB_HD float perp2(pt2 a, pt2 b) { return a.x * b.y - a.y * b.x; }
RB_HD pt2 sub2(pt2 a, pt2 b) { pt2 r; r.x = a.x - b.x; r.y = a.y - b.y; return r; }

RB_HD bool ulps_close(float a, float b) {
    float d = a - b;
    if (d < 0.0f) d = -d;
    if (d <= EPS32) return true;
    if ((a < 0.0f) != (b < 0.0f)) return false;
    int ia = as_int(a), ib = as_int(b);
    int df = ia > ib ? ia - ib : ib - ia;
    return df <= 4;
}

RB_HD bool line_line_2d(pt2 a0, pt2 a1, pt2 b0, pt2 b1, float& s, float& t) {
    pt2 d1 = sub2(a1, a0), d2 = sub2(b1, b0), r = sub2(a0, b0);
    float a = d1.x * d1.x + d1.y * d1.y;
    float e = d2.x * d2.x + d2.y * d2.y;
    float f = d2.x * r.x + d2.y * r.y;
    if (a <= EPS32 && e <= EPS32) { s = 0.0f; t = 0.0f; return true; }
    if (a <= EPS32) { s = 0.0f; t = f / e; return true; }
    float c = d1.x * r.x + d1.y * r.y;
    if (e <= EPS32) { s = -c / a; t = 0.0f; return true; }
    float b = d1.x * d2.x + d1.y * d2.y;
    float ae = a * e, bb = b * b, den = ae - bb;
    if (den <= EPS32 || ulps_close(ae, bb)) return false;
    s = (b * f - c * e) / den;
    t = (b * s + f) / e;
    return true;
}

RB_HD void raw_push(RawManifold& m, vec3 p1, vec3 p2, uint32_t f1, uint32_t f2, float d) {
    if (m.n >= MAX_RAW) return;
    RawPt& q = m.pt[m.n++];
    q.p1 = p1; q.p2 = p2; q.fid1 = f1; q.fid2 = f2; q.dist = d == 0.0f ? 0.0f : d;   // canonical zero (a signed zero carries no meaning here)
}

RB_HD void clip_faces(const pose& p12, const QuadFace& f1, vec3 axis, const QuadFace& f2, RawManifold& m) {
    vec3 b0, b1;
    ortho_basis(axis, b0, b1);
    pt2 q1[4], q2[4];
    vec3 w2[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        q1[i].x = dot3(f1.v[i], b0); q1[i].y = dot3(f1.v[i], b1);
        w2[i] = xform(p12, f2.v[i]);
        q2[i].x = dot3(w2[i], b0); q2[i].y = dot3(w2[i], b1);
    }
    {
        vec3 nrm2 = cross3(w2[2] - w2[1], w2[0] - w2[1]);
        float den = dot3(nrm2, axis);
        if (fabsf(den) > EPS32) {
            for (int i = 0; i < 4; ++i) {
                pt2 p = q1[i];
                float sg = perp2(sub2(q2[0], q2[3]), sub2(p, q2[3]));
                bool out = false;
                for (int j = 0; j < 3; ++j) {
                    float ns = perp2(sub2(q2[j + 1], q2[j]), sub2(p, q2[j]));
                    if (ns * sg < 0.0f) { out = true; break; }
                }
                if (out) continue;
                float d = dot3(w2[0] - f1.v[i], nrm2) / den;
                vec3 lp1 = f1.v[i];
                vec3 lp2 = xform_inv(p12, f1.v[i] + axis * d);
                raw_push(m, lp1, lp2, f1.vid[i], f2.fid, d);
            }
        }
    }
    {
        vec3 nrm1 = cross3(f1.v[2] - f1.v[1], f1.v[0] - f1.v[1]);
        float den = -dot3(nrm1, axis);
        if (fabsf(den) > EPS32) {
            for (int i = 0; i < 4; ++i) {
                pt2 p = q2[i];
                float sg = perp2(sub2(q1[0], q1[3]), sub2(p, q1[3]));
                bool out = false;
                for (int j = 0; j < 3; ++j) {
                    float ns = perp2(sub2(q1[j + 1], q1[j]), sub2(p, q1[j]));
                    if (ns * sg < 0.0f) { out = true; break; }
                }
                if (out) continue;
                float d = dot3(f1.v[0] - w2[i], nrm1) / den;
                vec3 lp2 = f2.v[i];
                vec3 lp1 = w2[i] - axis * d;
                raw_push(m, lp1, lp2, f1.fid, f2.vid[i], d);
            }
        }
    }
    for (int j = 0; j < 4; ++j) {
        pt2 c0 = q2[j], c1 = q2[(j + 1) & 3];
        for (int i = 0; i < 4; ++i) {
            pt2 a0 = q1[i], a1 = q1[(i + 1) & 3];
            float s, t;
            if (!line_line_2d(a0, a1, c0, c1, s, t)) continue;
            if (s > 0.0f && s < 1.0f && t > 0.0f && t < 1.0f) {
                vec3 lp1 = f1.v[i] * (1.0f - s) + f1.v[(i + 1) & 3] * s;
                vec3 lp21 = w2[j] * (1.0f - t) + w2[(j + 1) & 3] * t;
                float d = dot3(lp21 - lp1, axis);
                vec3 lp2 = xform_inv(p12, lp21);
                raw_push(m, lp1, lp2, f1.eid[i], f2.eid[j], d);
            }
        }
    }
}

RB_HD void manifold_box_box(vec3 he1, vec3 he2, const pose& p12, float prediction, RawManifold& m) {
    m.n = 0; m.n1 = zero3(); m.n2 = zero3();
    pose p21 = pinverse(p12);
    float s1, s2, s3;
    vec3 d1, d2, d3;
    sat_faces(he1, he2, p12, s1, d1);
    if (s1 > prediction) return;
    sat_faces(he2, he1, p21, s2, d2);
    if (s2 > prediction) return;
    sat_edges(he1, he2, p12, s3, d3);
    if (s3 > prediction) return;
    vec3 best = d1;
    if (s2 > s1 && s2 > s3) best = rotate(p12.q, -d2);
    else if (s3 > s1) best = d3;
    vec3 n2 = rotate(p21.q, -best);
    QuadFace f1, f2;
    box_support_face(he1, best, f1);
    box_support_face(he2, n2, f2);
    clip_faces(p12, f1, best, f2, m);
    m.n1 = best;
    m.n2 = n2;
}

RB_HD void manifold_ball_ball(float r1, float r2, const pose& p12, float prediction, RawManifold& m) {
    m.n = 0; m.n1 = zero3(); m.n2 = zero3();
    vec3 dc = p12.t;
    float cd = norm(dc);
    float d = cd - r1 - r2;
    if (!(d < prediction)) return;
    vec3 n1 = cd != 0.0f ? dc * (1.0f / cd) : mk3(0.f, 1.f, 0.f);
    vec3 n2 = rotate_inv(p12.q, -n1);
    raw_push(m, n1 * r1, n2 * r2, FID_FACE, FID_FACE, d);
    m.n1 = n1;
    m.n2 = n2;
}

// Ball (pose `pb` in the box frame) against a box: closest point / least-penetration face.
RB_HD bool box_ball(vec3 he, float r, const pose& pb, float prediction, vec3& p_box, vec3& p_ball, vec3& n_box,
                    vec3& n_ball, float& dist, uint32_t& fid) {
    vec3 c = pb.t;
    vec3 lo = mk3(-he.x - c.x, -he.y - c.y, -he.z - c.z);
    vec3 hi = mk3(c.x - he.x, c.y - he.y, c.z - he.z);
    vec3 sh = mk3(max2(lo.x, 0.0f) - max2(hi.x, 0.0f), max2(lo.y, 0.0f) - max2(hi.y, 0.0f), max2(lo.z, 0.0f) - max2(hi.z, 0.0f));
    bool inside = sh.x == 0.0f && sh.y == 0.0f && sh.z == 0.0f;
    vec3 proj;
    if (!inside) {
        proj = c + sh;
        fid = FID_FACE;
    } else {
        float best = -FMAX32;
        int bi = 0;
        float bs = 1.0f;
        for (int i = 0; i < 3; ++i) {
            if (comp(hi, i) > best) { best = comp(hi, i); bi = i; bs = 1.0f; }
            if (comp(lo, i) > best) { best = comp(lo, i); bi = i; bs = -1.0f; }
        }
        proj = with_comp(c, bi, bs * comp(he, bi));
        fid = FID_FACE | (uint32_t)(bi + (bs < 0.0f ? 3 : 0));
    }
    vec3 dp = c - proj;
    float d = norm(dp);
    if (!(d > 0.0f)) return false;
    vec3 n1 = dp * (1.0f / d);
    if (inside) { n1 = -n1; d = -d; }
    if (!(d <= r + prediction)) return false;
    vec3 n2 = rotate_inv(pb.q, -n1);
    p_box = proj; p_ball = n2 * r; n_box = n1; n_ball = n2; dist = d - r;
    return true;
}


// ------------------------------------------------------------------------------------------------
// Capsules (parry Capsule: a segment + a radius; ColliderBuilder::capsule_{x,y,z}).  he = (half height, radius, axis).
// Restated from first principles like the cuboid manifolds (parry's contact_manifold_capsule_capsule /
// _cuboid_capsule / _ball_convex are not in the tree): closest features of the segment cores, contact points on the
// surfaces, `dist` = surface distance.  Capsule-capsule: two points when the axes are parallel and overlap, else one;
// cuboid-capsule: the segment clipped against the reference face (up to two points), or the closest points of the
// segment and the supporting box edge; ball-capsule: one point.
// ------------------------------------------------------------------------------------------------
RB_HD vec3 capsule_dir(vec3 he) { return he.z == 0.0f ? mk3(1.f, 0.f, 0.f) : (he.z == 1.0f ? mk3(0.f, 1.f, 0.f) : mk3(0.f, 0.f, 1.f)); }

// closest points of two segments p1 + s d1, p2 + t d2 (s, t in [0, 1]); Ericson, Real-Time Collision Detection, 5.1.9
RB_HD void seg_seg_params(vec3 p1, vec3 d1, vec3 p2, vec3 d2, float& s, float& t) {
    const vec3 r = p1 - p2;
    const float a = dot3(d1, d1), e = dot3(d2, d2), f = dot3(d2, r);
    if (a <= EPS32 && e <= EPS32) { s = 0.0f; t = 0.0f; return; }
    if (a <= EPS32) { s = 0.0f; t = clampf(f / e, 0.0f, 1.0f); return; }
    const float c = dot3(d1, r);
    if (e <= EPS32) { t = 0.0f; s = clampf(-c / a, 0.0f, 1.0f); return; }
    const float b = dot3(d1, d2);
    const float denom = a * e - b * b;
    s = denom > 1.0e-6f * a * e ? clampf((b * f - c * e) / denom, 0.0f, 1.0f) : 0.0f;
    t = (b * s + f) / e;
    if (t < 0.0f) { t = 0.0f; s = clampf(-c / a, 0.0f, 1.0f); }
    else if (t > 1.0f) { t = 1.0f; s = clampf((b - c) / a, 0.0f, 1.0f); }
}

RB_HD void manifold_capsule_capsule(vec3 he1, vec3 he2, const pose& p12, float prediction, RawManifold& m) {
    m.n = 0; m.n1 = zero3(); m.n2 = zero3();
    const float hh1 = he1.x, r1 = he1.y, hh2 = he2.x, r2 = he2.y;
    const vec3 u1 = capsule_dir(he1), u2 = rotate(p12.q, capsule_dir(he2));
    const vec3 a1 = u1 * (-hh1), d1 = u1 * (2.0f * hh1);
    const vec3 a2 = p12.t - u2 * hh2, d2 = u2 * (2.0f * hh2);
    const vec3 cr = cross3(d1, d2);
    const float l1 = norm2(d1), l2 = norm2(d2);
    if (l1 > EPS32 && l2 > EPS32 && norm2(cr) <= 1.0e-6f * l1 * l2) {   // parallel axes: the shared interval, if any, gives two contacts
        const float inv = 1.0f / l1;
        const float ta = dot3(a2 - a1, d1) * inv, tb = dot3((a2 + d2) - a1, d1) * inv;
        const float lo = max2(min2(ta, tb), 0.0f), hi = min2(max2(ta, tb), 1.0f);
        if (hi > lo) {
            const vec3 w0 = a2 - a1;
            const vec3 wv = w0 - d1 * (dot3(w0, d1) * inv);   // offset between the two lines
            const float wl = norm(wv);
            const vec3 n1 = wl > EPS32 ? wv * (1.0f / wl) : ortho_vector(u1);
            const float dist = wl - r1 - r2;
            if (!(dist < prediction)) return;
            const float ts[2] = {lo, hi};
#pragma unroll
            for (int k = 0; k < 2; ++k) {
                const vec3 q1 = a1 + d1 * ts[k];
                const vec3 q2 = q1 + n1 * wl;
                raw_push(m, q1 + n1 * r1, xform_inv(p12, q2 - n1 * r2), FID_VERTEX | (uint32_t)k, FID_VERTEX | (uint32_t)k, dist);
            }
            m.n1 = n1;
            m.n2 = rotate_inv(p12.q, -n1);
            return;
        }
    }
    float s, t;
    seg_seg_params(a1, d1, a2, d2, s, t);
    const vec3 q1 = a1 + d1 * s, q2 = a2 + d2 * t;
    const vec3 dl = q2 - q1;
    const float len = norm(dl);
    vec3 n1;
    if (len > EPS32) n1 = dl * (1.0f / len);
    else { const float cl = norm(cr); n1 = cl > EPS32 ? cr * (1.0f / cl) : ortho_vector(u1); }
    const float dist = len - r1 - r2;
    if (!(dist < prediction)) return;
    raw_push(m, q1 + n1 * r1, xform_inv(p12, q2 - n1 * r2), FID_EDGE | 2u, FID_EDGE | 2u, dist);
    m.n1 = n1;
    m.n2 = rotate_inv(p12.q, -n1);
}

// Ball (pose `pb` in the capsule's frame) against a capsule.
RB_HD bool capsule_ball(vec3 hec, float rb, const pose& pb, float prediction, vec3& p_cap, vec3& p_ball, vec3& n_cap, vec3& n_ball, float& dist) {
    const float hh = hec.x, rc = hec.y;
    const vec3 u = capsule_dir(hec);
    const vec3 a = u * (-hh), d = u * (2.0f * hh);
    const float l2 = norm2(d);
    const float t = l2 > EPS32 ? clampf(dot3(pb.t - a, d) / l2, 0.0f, 1.0f) : 0.0f;
    const vec3 q = a + d * t;
    const vec3 dl = pb.t - q;
    const float len = norm(dl);
    const vec3 n = len > EPS32 ? dl * (1.0f / len) : ortho_vector(u);
    dist = len - rc - rb;
    if (!(dist < prediction)) return false;
    n_cap = n; n_ball = rotate_inv(pb.q, -n);
    p_cap = q + n * rc; p_ball = n_ball * rb;
    return true;
}

// Capsule (pose `pc` in the box's frame) against a box: points in the box's / the capsule's frame, normal from the box.
RB_HD void box_capsule(vec3 he, vec3 hec, const pose& pc, float prediction, RawManifold& m) {
    m.n = 0; m.n1 = zero3(); m.n2 = zero3();
    const float hh = hec.x, r = hec.y;
    const vec3 u = rotate(pc.q, capsule_dir(hec));
    const vec3 A = pc.t - u * hh, B = pc.t + u * hh;
    float best = -FMAX32;
    vec3 n = mk3(0.f, 1.f, 0.f);
    int kind = 0, bi = 0;   // kind 0 = box face bi (sign in n), 1 = box edge parallel to axis bi
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        const float ai = comp(A, i), bb = comp(B, i), h = comp(he, i);
        const float sp = min2(ai, bb) - h, sm = -max2(ai, bb) - h;
        if (sp > best) { best = sp; n = with_comp(zero3(), i, 1.0f); kind = 0; bi = i; }
        if (sm > best) { best = sm; n = with_comp(zero3(), i, -1.0f); kind = 0; bi = i; }
    }
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        const vec3 c = cross3(with_comp(zero3(), i, 1.0f), u);
        const float l2 = norm2(c);
        if (!(l2 > 1.0e-6f)) continue;
        vec3 nn = c * (1.0f / sqrtf(l2));
        float s0 = dot3(nn, A);
        if (s0 < 0.0f) { nn = -nn; s0 = -s0; }
        const float sep = s0 - fma_(he.z, fabsf(nn.z), fma_(he.y, fabsf(nn.y), he.x * fabsf(nn.x)));
        if (sep > best) { best = sep; n = nn; kind = 1; bi = i; }
    }
    if (!(best - r < prediction)) return;
    const vec3 D = B - A;
    if (kind == 0) {
        const float sg = comp(n, bi);
        float t0 = 0.0f, t1 = 1.0f;
        bool miss = false;
#pragma unroll
        for (int j = 0; j < 3; ++j) {   // Liang-Barsky against the side planes of the face
            if (j == bi) continue;
            const float aj = comp(A, j), dj = comp(D, j), h = comp(he, j);
            if (fabsf(dj) <= EPS32) { if (fabsf(aj) > h) miss = true; continue; }
            float ta = (-h - aj) / dj, tb = (h - aj) / dj;
            if (ta > tb) { const float tmp = ta; ta = tb; tb = tmp; }
            t0 = max2(t0, ta); t1 = min2(t1, tb);
        }
        if (!miss && t0 <= t1) {
            const float ts[2] = {t0, t1};
            const int np = (t1 - t0 > 1.0e-5f) ? 2 : 1;
            for (int k = 0; k < np; ++k) {
                const vec3 q = A + D * ts[k];
                const float dp = sg * comp(q, bi) - comp(he, bi);
                const float dist = dp - r;
                if (!(dist < prediction)) continue;
                raw_push(m, q - n * dp, xform_inv(pc, q - n * r), FID_FACE | (uint32_t)(bi + (sg < 0.0f ? 3 : 0)), FID_VERTEX | (uint32_t)k, dist);
            }
            m.n1 = n;
            m.n2 = rotate_inv(pc.q, -n);
            return;
        }
        // the segment passes beside the face: closest point of the box to the nearer end of the segment
        const vec3 q = (sg * comp(A, bi) <= sg * comp(B, bi)) ? A : B;
        const vec3 pb = mk3(clampf(q.x, -he.x, he.x), clampf(q.y, -he.y, he.y), clampf(q.z, -he.z, he.z));
        const vec3 dl = q - pb;
        const float len = norm(dl);
        const vec3 nn = len > EPS32 ? dl * (1.0f / len) : n;
        const float dist = len - r;
        if (!(dist < prediction)) return;
        raw_push(m, pb, xform_inv(pc, q - nn * r), box_vertex_id(pb), FID_VERTEX | 3u, dist);
        m.n1 = nn;
        m.n2 = rotate_inv(pc.q, -nn);
        return;
    }
    // box edge parallel to axis bi at the corner the normal points to, against the capsule's axis
    vec3 e0 = box_support(he, n), ed = zero3();
    e0 = with_comp(e0, bi, -comp(he, bi));
    ed = with_comp(ed, bi, 2.0f * comp(he, bi));
    float s, t;
    seg_seg_params(e0, ed, A, D, s, t);
    const vec3 pe = e0 + ed * s, q = A + D * t;
    const float dist = dot3(q - pe, n) - r;
    if (!(dist < prediction)) return;
    raw_push(m, pe, xform_inv(pc, q - n * r), FID_EDGE | ((uint32_t)bi << 4) | (box_vertex_id(e0) & 7u), FID_EDGE | 2u, dist);
    m.n1 = n;
    m.n2 = rotate_inv(pc.q, -n);
}

RB_HD void manifold_flip(const RawManifold& a, RawManifold& m) {   // the same contacts seen from the other shape
    m.n = a.n; m.n1 = a.n2; m.n2 = a.n1;
    for (int i = 0; i < a.n; ++i) { m.pt[i].p1 = a.pt[i].p2; m.pt[i].p2 = a.pt[i].p1; m.pt[i].dist = a.pt[i].dist; m.pt[i].fid1 = a.pt[i].fid2; m.pt[i].fid2 = a.pt[i].fid1; }
}
RB_HD void contact_manifold_capsules(int sh1, vec3 he1, int sh2, vec3 he2, const pose& p12, float prediction, RawManifold& m) {
    m.n = 0; m.n1 = zero3(); m.n2 = zero3();
    if (sh1 == SHAPE_CAPSULE && sh2 == SHAPE_CAPSULE) { manifold_capsule_capsule(he1, he2, p12, prediction, m); return; }
    if (sh1 == SHAPE_CUBOID) { box_capsule(he1, he2, p12, prediction, m); return; }
    if (sh2 == SHAPE_CUBOID) {
        RawManifold t;
        box_capsule(he2, he1, pinverse(p12), prediction, t);
        manifold_flip(t, m);
        return;
    }
    vec3 pc, pb, nc, nb; float d;
    if (sh1 == SHAPE_CAPSULE) {   // capsule - ball
        if (capsule_ball(he1, he2.x, p12, prediction, pc, pb, nc, nb, d)) { raw_push(m, pc, pb, FID_EDGE | 2u, FID_FACE, d); m.n1 = nc; m.n2 = nb; }
    } else {                      // ball - capsule
        if (capsule_ball(he2, he1.x, pinverse(p12), prediction, pc, pb, nc, nb, d)) { raw_push(m, pb, pc, FID_FACE, FID_EDGE | 2u, d); m.n1 = nb; m.n2 = nc; }
    }
}

}  // namespace rb
#include "rb_poly.cuh"
namespace rb {

// SHAPES = 1: worlds with capsules or convex polyhedra (a compile-time variant of the collision kernel, so the ball / cuboid one is unchanged)
template <int SHAPES = 0>
RB_HD void contact_manifold(const HullTables& hulls, int sh1, vec3 he1, int sh2, vec3 he2, const pose& p12, float prediction, RawManifold& m) {
    m.n = 0; m.n1 = zero3(); m.n2 = zero3();
    if (SHAPES && (sh1 == SHAPE_CONVEX || sh2 == SHAPE_CONVEX)) { contact_manifold_convex(hulls, sh1, he1, sh2, he2, p12, prediction, m); return; }
    if (SHAPES && (sh1 == SHAPE_CAPSULE || sh2 == SHAPE_CAPSULE)) { contact_manifold_capsules(sh1, he1, sh2, he2, p12, prediction, m); return; }
    if (sh1 == SHAPE_CUBOID && sh2 == SHAPE_CUBOID) {
        manifold_box_box(he1, he2, p12, prediction, m);
    } else if (sh1 == SHAPE_BALL && sh2 == SHAPE_BALL) {
        manifold_ball_ball(he1.x, he2.x, p12, prediction, m);
    } else if (sh1 == SHAPE_CUBOID) {
        vec3 pc, pb, nc, nb; float d; uint32_t fid;
        if (box_ball(he1, he2.x, p12, prediction, pc, pb, nc, nb, d, fid)) {
            raw_push(m, pc, pb, fid, FID_FACE, d);
            m.n1 = nc; m.n2 = nb;
        }
    } else {
        pose p21 = pinverse(p12);
        vec3 pc, pb, nc, nb; float d; uint32_t fid;
        if (box_ball(he2, he1.x, p21, prediction, pc, pb, nc, nb, d, fid)) {
            raw_push(m, pb, pc, FID_FACE, fid, d);
            m.n1 = nb; m.n2 = nc;
        }
    }
}

RB_HD void shape_aabb(int shape, vec3 he, const pose& p, vec3& lo, vec3& hi) {
    vec3 ws;
    if (shape == SHAPE_BALL) {
        ws = mk3(he.x, he.x, he.x);
    } else if (shape == SHAPE_CAPSULE) {   // the segment's box loosened by the radius
        const vec3 u = rotate(p.q, capsule_dir(he));
        ws = mk3(fabsf(u.x) * he.x + he.y, fabsf(u.y) * he.x + he.y, fabsf(u.z) * he.x + he.y);
    } else {
        mat3 r = rotmat(p.q);
        ws = mk3(fabsf(r.c0.x) * he.x + fabsf(r.c1.x) * he.y + fabsf(r.c2.x) * he.z,
                 fabsf(r.c0.y) * he.x + fabsf(r.c1.y) * he.y + fabsf(r.c2.y) * he.z,
                 fabsf(r.c0.z) * he.x + fabsf(r.c1.z) * he.y + fabsf(r.c2.z) * he.z);
    }
    lo = p.t - ws;
    hi = p.t + ws;
}

}  // namespace rb
