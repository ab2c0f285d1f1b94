// rb_poly.cuh -- contact manifolds of convex polyhedra (ColliderBuilder::convex_hull / convex_mesh / round_convex_hull).
//
// Replaces the reference's call into parry3d 0.30.2 (`query_dispatcher.contact_manifolds`,
// src/geometry/narrow_phase/pair_update.rs:323-330) for pairs that involve a ConvexPolyhedron.  parry's own routine
// (GJK/EPA + PolygonalFeatureMap clipping) is not in the tree; this is a first-principles restatement with the same
// contract as the cuboid path: separating-axis search (face normals of both shapes, then the edge pairs that form a face of the Minkowski
// difference -- pruned on the Gauss map; for a capsule's segment: supporting edge pairs),
// reference / incident feature selection, Sutherland-Hodgman clipping of the incident face against the side planes of
// the reference face, points within `prediction` kept, `dist` measured along the manifold normal, stable feature ids.
// A cuboid meets a polyhedron as the unit-cube topology (hull 0) scaled by its half extents, a capsule as a segment
// (two vertices, one edge, no face) with a radius; a border radius ("round" shapes) is handled like a capsule's.
// Only the SHAPES = 1 collision kernel compiles this.
#pragma once
#include "rb_world.cuh"

namespace rb {

struct Poly {   // one convex polyhedron in its own frame (nf == 0: a segment)
    const float4* v;    // vertices
    const float4* pl;   // face planes: outward normal, offset
    const int2* fc;     // faces: start in lp, vertex count (counter-clockwise seen from outside)
    const int* lp;      // face loops: vertex indices
    const int4* ed;     // edges: v0, v1, face on which the edge runs v0 -> v1, the other face
    int nv, nf, ne;
};
struct PolyLocal { float4 v[8]; float4 pl[6]; int4 seg; };   // scratch of a cuboid / capsule view

RB_HD Poly poly_of_hull(const HullTables& h, int id) {
    const int4 d = h.desc[id], d2 = h.desc2[id];
    Poly p;
    p.v = h.verts + d.x; p.nv = d.y; p.pl = h.planes + d.z; p.fc = h.faces + d.z; p.nf = d.w;
    p.ed = h.edges + d2.x; p.ne = d2.y; p.lp = h.loops + d2.z;
    return p;
}
// (shape, he) of a collider as a polyhedron + radius.  Cuboid: hull 0 (the unit cube) scaled; capsule: its segment.
RB_HD Poly poly_of_shape(const HullTables& h, int shape, vec3 he, PolyLocal& st, float& radius) {
    if (shape == SHAPE_CONVEX) { radius = he.y; return poly_of_hull(h, (int)he.x); }
    Poly p = poly_of_hull(h, 0);
    radius = 0.0f;
    if (shape == SHAPE_CUBOID) {
        for (int i = 0; i < 8; ++i) { const float4 u = p.v[i]; st.v[i] = make_float4(u.x * he.x, u.y * he.y, u.z * he.z, 0.0f); }
        for (int f = 0; f < 6; ++f) {
            const float4 q = p.pl[f];
            st.pl[f] = make_float4(q.x, q.y, q.z, fma_(he.z, fabsf(q.z), fma_(he.y, fabsf(q.y), he.x * fabsf(q.x))));
        }
        p.v = st.v; p.pl = st.pl;
        return p;
    }
    // capsule: he = (half height, radius, axis)
    const vec3 u = he.z == 0.0f ? mk3(1.f, 0.f, 0.f) : (he.z == 1.0f ? mk3(0.f, 1.f, 0.f) : mk3(0.f, 0.f, 1.f));
    st.v[0] = f4(u * -he.x, 0.0f); st.v[1] = f4(u * he.x, 0.0f);
    st.seg = make_int4(0, 1, -1, -1);
    p.v = st.v; p.nv = 2; p.nf = 0; p.ed = &st.seg; p.ne = 1;
    radius = he.y;
    return p;
}

struct ClipPt { vec3 p; uint32_t id; int eout; };   // eout: the incident edge the segment to the NEXT point lies on, or -1 - j: on side plane j
constexpr uint32_t PID_VERT = 0x100u, PID_EDGE = 0x200u, PID_CORNER = 0x400u;
constexpr int POLY_CLIP_MAX = 2 * HULL_MAX_FACE_VERTS;

// Clips the polygon (or, closed == false, the polyline) `in` against the half space dot(sn, x) <= sd (side plane j of the reference face).
RB_HD int poly_clip_plane(const ClipPt* in, int n, bool closed, vec3 sn, float sd, int j, ClipPt* out) {
    int m = 0;
    const int segs = closed ? n : n - 1;
    for (int i = 0; i < segs; ++i) {
        const ClipPt& P = in[i];
        const ClipPt& Q = in[i + 1 < n ? i + 1 : 0];
        const float dp = dot3(sn, P.p) - sd, dq = dot3(sn, Q.p) - sd;
        const bool pin = dp <= 0.0f, qin = dq <= 0.0f;
        if (pin && m < POLY_CLIP_MAX) out[m++] = P;
        if (pin != qin && m < POLY_CLIP_MAX) {
            const float t = dp / (dp - dq);
            ClipPt X;
            X.p = P.p + (Q.p - P.p) * t;
            X.id = P.eout >= 0 ? (PID_EDGE | ((uint32_t)P.eout << 4) | (uint32_t)j) : (PID_CORNER | ((uint32_t)(-1 - P.eout) << 4) | (uint32_t)j);
            X.eout = pin ? -1 - j : P.eout;
            out[m++] = X;
        }
    }
    if (!closed && n > 0) {   // the last point of a polyline has no outgoing segment
        const ClipPt& P = in[n - 1];
        if (dot3(sn, P.p) - sd <= 0.0f && m < POLY_CLIP_MAX) out[m++] = P;
    }
    return m;
}

// Face `rf` of the reference shape (vertices rv, outward normal n, offset d, all in ONE frame) against the incident
// feature of the other shape (vertices iv in the same frame): its face most opposed to n, or the segment itself.
// `inormals`: the incident shape's face normals in that frame, or null = its own planes.  Emits the clipped points.
RB_HD int poly_clip_incident(const Poly& R, const vec3* rv, int rf, vec3 n, const Poly& I, const vec3* iv, const vec3* inormals,
                             ClipPt* out) {
    ClipPt a[POLY_CLIP_MAX], b[POLY_CLIP_MAX];
    int cnt = 0;
    bool closed = true;
    if (I.nf == 0) {
        closed = false;
        for (int k = 0; k < 2; ++k) { a[k].p = iv[k]; a[k].id = PID_VERT | (uint32_t)k; a[k].eout = 0; }
        cnt = 2;
    } else {
        int inc = 0;
        float most = FMAX32;
        for (int f = 0; f < I.nf; ++f) {
            const float c = dot3(inormals ? inormals[f] : xyz(I.pl[f]), n);
            if (c < most) { most = c; inc = f; }
        }
        const int2 fc = I.fc[inc];
        for (int k = 0; k < fc.y; ++k) {
            const int vi = I.lp[fc.x + k];
            a[k].p = iv[vi]; a[k].id = (PID_VERT | (uint32_t)vi) | ((uint32_t)inc << 16); a[k].eout = k;   // (the incident face is part of the feature id)
        }
        cnt = fc.y;
    }
    const int2 rfc = R.fc[rf];
    ClipPt* src = a;
    ClipPt* dst = b;
    for (int j = 0; j < rfc.y && cnt > 0; ++j) {
        const vec3 e0 = rv[R.lp[rfc.x + j]], e1 = rv[R.lp[rfc.x + (j + 1 < rfc.y ? j + 1 : 0)]];
        const vec3 sn = cross3(e1 - e0, n);   // outward side normal (the loop runs counter-clockwise about n)
        cnt = poly_clip_plane(src, cnt, closed, sn, dot3(sn, e0), j, dst);
        if (cnt < 3) closed = false;   // (a polygon clipped down to a sliver goes on as a polyline)
        ClipPt* t = src; src = dst; dst = t;
    }
    for (int k = 0; k < cnt; ++k) out[k] = src[k];
    return cnt;
}

// Warp helpers of the cooperative search (one lane, no-ops, in the host emulation).
#if RB_DEVICE_BUILD
RB_D void poly_warp_sync() { __syncwarp(); }
RB_D bool poly_warp_any(bool p) { return __any_sync(0xffffffffu, p); }
RB_D void poly_warp_argmax(float& s, int& idx) {   // largest s, the SMALLEST index among equals (= the first one a serial scan meets)
    for (int o = 16; o > 0; o >>= 1) {
        const float s2 = __shfl_xor_sync(0xffffffffu, s, o);
        const int i2 = __shfl_xor_sync(0xffffffffu, idx, o);
        if (s2 > s || (s2 == s && i2 < idx)) { s = s2; idx = i2; }
    }
}
#else
inline void poly_warp_sync() {}
inline bool poly_warp_any(bool p) { return p; }
inline void poly_warp_argmax(float&, int&) {}
#endif

// Separation of the edge pair (ea of A, eb of B) along the normalised cross product of the edges, or false when the
// pair is no candidate: with faces on both sides it must be a face of the Minkowski difference (the arcs of the two
// edges cross on the Gauss map; Gregorius, GDC 2013), else -- a capsule's segment -- a SUPPORTING pair along the axis.
RB_HD bool poly_edge_axis(bool gauss, const Poly& A, const Poly& B, const vec3* va, const vec3* vb, const vec3* nb, vec3 ca, int ea, int eb,
                          vec3& n, float& s) {
    const int4 e1 = A.ed[ea], e2 = B.ed[eb];
    const vec3 da = va[e1.y] - va[e1.x], db = vb[e2.y] - vb[e2.x];
    if (gauss) {
        const vec3 u1 = xyz(A.pl[e1.z]), v1 = xyz(A.pl[e1.w]);
        const float du2 = dot3(nb[e2.z], da), dv2 = dot3(nb[e2.w], da), du1 = dot3(u1, db), dv1 = dot3(v1, db);
        if (!(du2 * dv2 < 0.0f && du1 * dv1 < 0.0f && du2 * dv1 < 0.0f)) return false;
    }
    const vec3 c = cross3(da, db);
    const float l2 = norm2(c);
    if (!(l2 > 1.0e-8f * norm2(da) * norm2(db))) return false;
    n = c * (1.0f / sqrtf(l2));
    if (gauss) {
        if (dot3(n, va[e1.x] - ca) < 0.0f) n = -n;   // away from A
        s = dot3(n, vb[e2.x] - va[e1.x]);
        return true;
    }
    float pa = dot3(n, va[e1.x]), pb = dot3(n, vb[e2.x]);
    if (pb < pa) { n = -n; pa = -pa; pb = -pb; }   // from A to B
    const float tol = 1.0e-5f * (1.0f + fabsf(pa) + fabsf(pb));
    bool support = true;
    for (int i = 0; i < A.nv && support; ++i) support = dot3(n, va[i]) <= pa + tol;
    for (int i = 0; i < B.nv && support; ++i) support = dot3(n, vb[i]) >= pb - tol;
    if (!support) return false;
    s = pb - pa;
    return true;
}

// Polyhedron A (radius ra, own frame) against polyhedron B (radius rb, pose p12 in A's frame), by the `nl` lanes of a
// warp together (lane = this one; nl = 1: one thread alone): the lanes share the candidate axes and agree on the best
// one -- largest separation, lowest candidate index among equals, which is what a serial scan in index order finds --
// then every lane derives the (small) manifold from it redundantly.  sva / svb / snb: 3 x 32 vec3 of scratch shared by
// the lanes (shared memory on the device).
RB_HD void manifold_poly_poly(int lane, int nl, const Poly& A, float ra, const Poly& B, float rb, const pose& p12, float prediction,
                              vec3* va, vec3* vb, vec3* nb, RawManifold& m) {
    m.n = 0; m.n1 = zero3(); m.n2 = zero3();
    const float eff = prediction + ra + rb;
    poly_warp_sync();   // (the scratch may still be read by a lane finishing the previous pair)
    for (int i = lane; i < A.nv; i += nl) va[i] = xyz(A.v[i]);
    for (int i = lane; i < B.nv; i += nl) vb[i] = xform(p12, xyz(B.v[i]));
    for (int f = lane; f < B.nf; f += nl) nb[f] = rotate(p12.q, xyz(B.pl[f]));
    poly_warp_sync();
    bool separated = false;
    float best = -FMAX32;
    int bidx = 0x7fffffff;
    for (int c = lane; c < A.nf + B.nf; c += nl) {   // face normals of A, then of B (in A's frame)
        float s = FMAX32;
        if (c < A.nf) {
            const vec3 n = xyz(A.pl[c]);
            for (int i = 0; i < B.nv; ++i) s = min2(s, dot3(n, vb[i]));
            s = s - A.pl[c].w;
        } else {
            const int f = c - A.nf;
            const vec3 n = nb[f];
            const float d = B.pl[f].w + dot3(n, p12.t);
            for (int i = 0; i < A.nv; ++i) s = min2(s, dot3(n, va[i]));
            s = s - d;
        }
        if (s > eff) separated = true;
        if (s > best) { best = s; bidx = c; }
    }
    if (poly_warp_any(separated)) return;
    poly_warp_argmax(best, bidx);
    const bool gauss = A.nf > 0 && B.nf > 0;
    vec3 ca = zero3();
    for (int i = 0; i < A.nv; ++i) ca = ca + va[i];
    ca = ca * (1.0f / (float)A.nv);
    float ebest = -FMAX32;
    int eidx = 0x7fffffff;
    const int npairs = A.ne * B.ne;
    for (int c = lane; c < npairs; c += nl) {
        vec3 n;
        float s;
        if (!poly_edge_axis(gauss, A, B, va, vb, nb, ca, c / B.ne, c % B.ne, n, s)) continue;
        if (s > eff) separated = true;
        if (s > ebest) { ebest = s; eidx = c; }
    }
    if (poly_warp_any(separated)) return;
    poly_warp_argmax(ebest, eidx);
    int kind, bi, bj = 0;
    vec3 bn;
    if (bidx == 0x7fffffff || ebest > best + 1.0e-4f) {
        if (eidx == 0x7fffffff) return;
        kind = 2; bi = eidx / B.ne; bj = eidx % B.ne;
        float s;
        poly_edge_axis(gauss, A, B, va, vb, nb, ca, bi, bj, bn, s);
    } else if (bidx < A.nf) {
        kind = 0; bi = bidx; bn = xyz(A.pl[bi]);
    } else {
        kind = 1; bi = bidx - A.nf; bn = -nb[bi];
    }
    const vec3 n2 = rotate_inv(p12.q, -bn);
    if (kind == 2) {
        const int4 e1 = A.ed[bi], e2 = B.ed[bj];
        float s, t;
        seg_seg_params(va[e1.x], va[e1.y] - va[e1.x], vb[e2.x], vb[e2.y] - vb[e2.x], s, t);
        const vec3 qa = va[e1.x] + (va[e1.y] - va[e1.x]) * s, qb = vb[e2.x] + (vb[e2.y] - vb[e2.x]) * t;
        const float dist = dot3(qb - qa, bn) - ra - rb;
        if (!(dist < prediction)) return;
        raw_push(m, qa + bn * ra, xform_inv(p12, qb - bn * rb), FID_EDGE | (uint32_t)bi, FID_EDGE | (uint32_t)bj, dist);
        m.n1 = bn; m.n2 = n2;
        return;
    }
    ClipPt pts[POLY_CLIP_MAX];
    int cnt;
    vec3 rn;   // outward normal of the reference face, in A's frame
    float rd;
    if (kind == 0) {
        rn = bn; rd = A.pl[bi].w;
        cnt = poly_clip_incident(A, va, bi, rn, B, vb, nb, pts);
    } else {
        rn = -bn; rd = B.pl[bi].w + dot3(rn, p12.t);
        cnt = poly_clip_incident(B, vb, bi, rn, A, va, nullptr, pts);
    }
    int keep[POLY_CLIP_MAX], nk = 0;
    for (int k = 0; k < cnt; ++k)
        if (dot3(rn, pts[k].p) - rd - ra - rb < prediction) keep[nk++] = k;
    const int take = nk < MAX_RAW ? nk : MAX_RAW;
    for (int k = 0; k < take; ++k) {
        const ClipPt& c = pts[keep[nk <= MAX_RAW ? k : (k * nk) / MAX_RAW]];   // (more than MAX_RAW: evenly spaced around the polygon)
        const float dc = dot3(rn, c.p) - rd;   // signed distance of the incident point to the reference plane
        const vec3 on_ref = c.p - rn * dc;
        if (kind == 0) raw_push(m, on_ref + bn * ra, xform_inv(p12, c.p - bn * rb), FID_FACE | (uint32_t)bi, c.id, dc - ra - rb);
        else raw_push(m, c.p + bn * ra, xform_inv(p12, on_ref - bn * rb), c.id, FID_FACE | (uint32_t)bi, dc - ra - rb);
    }
    m.n1 = bn; m.n2 = n2;
}

// Raw manifolds of polyhedron pairs travel from the cooperative pass to the per-pair pass through a global scratch table.
constexpr int POLY_RAW_STRIDE = 8 + MAX_RAW * 9;   // n, n1, n2, pad, then p1 p2 dist fid1 fid2 per point
RB_HD void poly_raw_store(float* d, const RawManifold& m) {
    d[0] = as_float_i(m.n); d[1] = m.n1.x; d[2] = m.n1.y; d[3] = m.n1.z; d[4] = m.n2.x; d[5] = m.n2.y; d[6] = m.n2.z;
    for (int k = 0; k < m.n; ++k) {
        float* q = d + 8 + 9 * k;
        const RawPt& p = m.pt[k];
        q[0] = p.p1.x; q[1] = p.p1.y; q[2] = p.p1.z; q[3] = p.p2.x; q[4] = p.p2.y; q[5] = p.p2.z; q[6] = p.dist;
        q[7] = as_float(p.fid1); q[8] = as_float(p.fid2);
    }
}
RB_HD void poly_raw_load(const float* d, RawManifold& m) {
    m.n = as_int(d[0]); m.n1 = mk3(d[1], d[2], d[3]); m.n2 = mk3(d[4], d[5], d[6]);
    for (int k = 0; k < m.n; ++k) {
        const float* q = d + 8 + 9 * k;
        RawPt& p = m.pt[k];
        p.p1 = mk3(q[0], q[1], q[2]); p.p2 = mk3(q[3], q[4], q[5]); p.dist = q[6]; p.fid1 = as_uint(q[7]); p.fid2 = as_uint(q[8]);
    }
}
RB_HD bool shape_is_poly(int sh) { return sh == SHAPE_CONVEX || sh == SHAPE_CUBOID || sh == SHAPE_CAPSULE; }
// pairs the cooperative pass computes: a convex polyhedron against a polyhedron, a cuboid or a capsule
RB_HD bool pair_is_poly_poly(int sh1, int sh2) { return (sh1 == SHAPE_CONVEX || sh2 == SHAPE_CONVEX) && shape_is_poly(sh1) && shape_is_poly(sh2); }

// Ball (centre c in the polyhedron's frame, radius r) against polyhedron P (border radius rp).
RB_HD bool poly_ball(const Poly& P, float rp, vec3 c, float r, float prediction, vec3& p_poly, vec3& n_poly, float& dist, uint32_t& fid) {
    float smax = -FMAX32;
    int fm = 0;
    for (int f = 0; f < P.nf; ++f) {
        const float s = dot3(xyz(P.pl[f]), c) - P.pl[f].w;
        if (s > smax) { smax = s; fm = f; }
    }
    if (smax > prediction + r + rp) return false;
    if (smax <= 0.0f) {   // the centre is inside: least-penetration face
        n_poly = xyz(P.pl[fm]);
        p_poly = c - n_poly * smax + n_poly * rp;
        dist = smax - r - rp;
        fid = FID_FACE | (uint32_t)fm;
        return true;
    }
    float bestd = FMAX32;
    vec3 bp = c;
    fid = FID_FACE;
    for (int f = 0; f < P.nf; ++f) {   // the closest point lies on a face turned towards the centre
        const vec3 n = xyz(P.pl[f]);
        const float s = dot3(n, c) - P.pl[f].w;
        if (!(s > 0.0f)) continue;
        const vec3 q = c - n * s;
        const int2 fc = P.fc[f];
        bool inside = true;
        for (int k = 0; k < fc.y; ++k) {
            const vec3 e0 = xyz(P.v[P.lp[fc.x + k]]), e1 = xyz(P.v[P.lp[fc.x + (k + 1 < fc.y ? k + 1 : 0)]]);
            const vec3 ed = e1 - e0;
            if (dot3(cross3(ed, n), q - e0) > 0.0f) {
                inside = false;
                const float l2 = norm2(ed);
                const float t = l2 > 0.0f ? clampf(dot3(c - e0, ed) / l2, 0.0f, 1.0f) : 0.0f;
                const vec3 x = e0 + ed * t;
                const float d2 = norm2(c - x);
                if (d2 < bestd) { bestd = d2; bp = x; fid = FID_EDGE | ((uint32_t)f << 8) | (uint32_t)k; }
            }
        }
        if (inside && s * s < bestd) { bestd = s * s; bp = q; fid = FID_FACE | (uint32_t)f; }
    }
    const vec3 dl = c - bp;
    const float len = norm(dl);
    if (!(len > 0.0f)) return false;
    if (!(len - r - rp < prediction)) return false;
    n_poly = dl * (1.0f / len);
    p_poly = bp + n_poly * rp;
    dist = len - r - rp;
    return true;
}

// Dispatch for a ball against a convex polyhedron (the other pairs with a polyhedron: phase_convex_manifolds, rb_collide.cuh).
RB_HD void contact_manifold_convex(const HullTables& h, int sh1, vec3 he1, int sh2, vec3 he2, const pose& p12, float prediction, RawManifold& m) {
    m.n = 0; m.n1 = zero3(); m.n2 = zero3();
    PolyLocal s1, s2;
    float r1, r2;
    if (sh2 == SHAPE_BALL) {
        const Poly P = poly_of_shape(h, sh1, he1, s1, r1);
        vec3 pp, np; float d; uint32_t fid;
        if (poly_ball(P, r1, p12.t, he2.x, prediction, pp, np, d, fid)) {
            const vec3 nb = rotate_inv(p12.q, -np);
            raw_push(m, pp, nb * he2.x, fid, FID_FACE, d);
            m.n1 = np; m.n2 = nb;
        }
        return;
    }
    if (sh1 == SHAPE_BALL) {
        const pose p21 = pinverse(p12);
        const Poly P = poly_of_shape(h, sh2, he2, s2, r2);
        vec3 pp, np; float d; uint32_t fid;
        if (poly_ball(P, r2, p21.t, he1.x, prediction, pp, np, d, fid)) {
            const vec3 nb = rotate_inv(p21.q, -np);
            raw_push(m, nb * he1.x, pp, FID_FACE, fid, d);
            m.n1 = nb; m.n2 = np;
        }
        return;
    }
    // (polyhedron against polyhedron / cuboid / capsule: phase_convex_manifolds computed it, the caller reads the scratch table)
}

RB_HD void convex_aabb(const HullTables& h, vec3 he, const pose& p, vec3& lo, vec3& hi) {
    const Poly P = poly_of_hull(h, (int)he.x);
    lo = mk3(FMAX32, FMAX32, FMAX32); hi = mk3(-FMAX32, -FMAX32, -FMAX32);
    for (int i = 0; i < P.nv; ++i) {
        const vec3 x = rotate(p.q, xyz(P.v[i]));
        lo = mk3(min2(lo.x, x.x), min2(lo.y, x.y), min2(lo.z, x.z));
        hi = mk3(max2(hi.x, x.x), max2(hi.y, x.y), max2(hi.z, x.z));
    }
    lo = mk3(lo.x - he.y, lo.y - he.y, lo.z - he.y) + p.t;
    hi = mk3(hi.x + he.y, hi.y + he.y, hi.z + he.y) + p.t;
}

}  // namespace rb
