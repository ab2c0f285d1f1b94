// TEST INFRASTRUCTURE ONLY -- contact manifolds of convex polyhedra for the CPU oracle (see oracle.h header note).
//
// Stands in for parry3d 0.30.2's contact_manifolds on pairs with a ConvexPolyhedron (pair_update.rs:323-330); parry is
// not in the tree.  Contract kept from the cuboid routine: the manifold normal is the axis of largest
// separation among the face normals of both shapes and the cross products of the edge pairs that form a face of the
// Minkowski difference (Gauss-map test; supporting edge pairs when one shape is a capsule's segment); a face axis gives
// the incident face of the other shape clipped (Sutherland-Hodgman) against the side planes of the reference face, an
// edge axis the closest points of the two edges; points closer than `prediction`; `dist` along the normal.  Cuboids
// enter as the scaled unit cube (hull 0), capsules as their segment with a radius; round polyhedra carry a radius too.
#include <cmath>

#include "oracle_internal.h"

namespace orc {

namespace {
const float F32_MAX = 3.4028235e38f;
const uint32_t ID_VERT = 0x100u, ID_EDGE = 0x200u, ID_CORNER = 0x400u, FEAT_FACE = 0x20000000u, FEAT_EDGE = 0x30000000u;
const int CLIP_MAX = 2 * HULL_MAX_FACE_VERTS;

struct Solid {   // a convex polyhedron in its own frame; no face = a segment
    std::vector<V3> v, n;
    std::vector<float> d;
    const Hull* topo = nullptr;   // faces / loops / edges (hull 0 for a cuboid)
    bool segment = false;
    float radius = 0.0f;
    int nfaces() const { return segment ? 0 : (int)topo->face_count.size(); }
    int nedges() const { return segment ? 1 : (int)topo->edges.size(); }
    HullEdge edge(int e) const { return segment ? HullEdge{0, 1, -1, -1} : topo->edges[e]; }
};

Solid solid_of(const std::vector<Hull>& hulls, int shape, V3 he) {
    Solid s;
    if (shape == RB_SHAPE_CONVEX) {
        const Hull& h = hulls[(int)he.x];
        s.v = h.verts; s.n = h.normals; s.d = h.offsets; s.topo = &h; s.radius = he.y;
        return s;
    }
    const Hull& cube = hulls[0];
    if (shape == RB_SHAPE_CUBOID) {
        for (const V3& u : cube.verts) s.v.push_back(V3{u.x * he.x, u.y * he.y, u.z * he.z});
        for (const V3& q : cube.normals) {
            s.n.push_back(q);
            s.d.push_back(fma_(he.z, fabsf(q.z), fma_(he.y, fabsf(q.y), he.x * fabsf(q.x))));
        }
        s.topo = &cube;
        return s;
    }
    V3 u = he.z == 0.0f ? V3{1.f, 0.f, 0.f} : (he.z == 1.0f ? V3{0.f, 1.f, 0.f} : V3{0.f, 0.f, 1.f});
    s.v.push_back(u * -he.x);
    s.v.push_back(u * he.x);
    s.segment = true;
    s.radius = he.y;
    return s;
}

struct ClipPoint { V3 p; uint32_t id; int eout; };

int clip_by_plane(const ClipPoint* in, int n, bool closed, V3 sn, float sd, int j, ClipPoint* out) {
    int m = 0;
    int segs = closed ? n : n - 1;
    for (int i = 0; i < segs; ++i) {
        const ClipPoint& P = in[i];
        const ClipPoint& Q = in[i + 1 < n ? i + 1 : 0];
        float dp = dot(sn, P.p) - sd, dq = dot(sn, Q.p) - sd;
        bool pin = dp <= 0.0f, qin = dq <= 0.0f;
        if (pin && m < CLIP_MAX) out[m++] = P;
        if (pin != qin && m < CLIP_MAX) {
            float t = dp / (dp - dq);
            ClipPoint X;
            X.p = P.p + (Q.p - P.p) * t;
            X.id = P.eout >= 0 ? (ID_EDGE | ((uint32_t)P.eout << 4) | (uint32_t)j) : (ID_CORNER | ((uint32_t)(-1 - P.eout) << 4) | (uint32_t)j);
            X.eout = pin ? -1 - j : P.eout;
            out[m++] = X;
        }
    }
    if (!closed && n > 0) {
        const ClipPoint& P = in[n - 1];
        if (dot(sn, P.p) - sd <= 0.0f && m < CLIP_MAX) out[m++] = P;
    }
    return m;
}

// reference face rf of R (vertices rv, outward normal n) against the incident feature of I (vertices iv, face normals in)
int clip_incident(const Solid& R, const std::vector<V3>& rv, int rf, V3 n, const Solid& I, const std::vector<V3>& iv,
                  const std::vector<V3>& in, ClipPoint* out) {
    ClipPoint a[CLIP_MAX], b[CLIP_MAX];
    int cnt = 0;
    bool closed = true;
    if (I.segment) {
        closed = false;
        for (int k = 0; k < 2; ++k) a[k] = ClipPoint{iv[k], ID_VERT | (uint32_t)k, 0};
        cnt = 2;
    } else {
        int inc = 0;
        float most = F32_MAX;
        for (int f = 0; f < I.nfaces(); ++f) {
            float c = dot(in[f], n);
            if (c < most) { most = c; inc = f; }
        }
        int s = I.topo->face_start[inc];
        cnt = I.topo->face_count[inc];
        for (int k = 0; k < cnt; ++k) {
            int vi = I.topo->loops[s + k];
            a[k] = ClipPoint{iv[vi], (ID_VERT | (uint32_t)vi) | ((uint32_t)inc << 16), k};
        }
    }
    int rs = R.topo->face_start[rf], rn = R.topo->face_count[rf];
    ClipPoint* src = a;
    ClipPoint* dst = b;
    for (int j = 0; j < rn && cnt > 0; ++j) {
        V3 e0 = rv[R.topo->loops[rs + j]], e1 = rv[R.topo->loops[rs + (j + 1 < rn ? j + 1 : 0)]];
        V3 sn = cross(e1 - e0, n);
        cnt = clip_by_plane(src, cnt, closed, sn, dot(sn, e0), j, dst);
        if (cnt < 3) closed = false;
        ClipPoint* t = src; src = dst; dst = t;
    }
    for (int k = 0; k < cnt; ++k) out[k] = src[k];
    return cnt;
}

inline void push(RawManifold& m, V3 p1, V3 p2, uint32_t f1, uint32_t f2, float dist) {
    if (m.n >= MAX_RAW_POINTS) return;
    RawPoint& q = m.pts[m.n++];
    q.local_p1 = p1; q.local_p2 = p2; q.fid1 = f1; q.fid2 = f2; q.dist = dist == 0.0f ? 0.0f : dist;
}

// Ericson 5.1.9, as in the capsule routines (segments p1 + s d1, p2 + t d2, s and t in [0, 1])
void closest_on_segments(V3 p1, V3 d1, V3 p2, V3 d2, float& s, float& t) {
    const float EPS = 1.1920929e-7f;
    V3 r = p1 - p2;
    float a = dot(d1, d1), e = dot(d2, d2), f = dot(d2, r);
    if (a <= EPS && e <= EPS) { s = 0.0f; t = 0.0f; return; }
    if (a <= EPS) { s = 0.0f; t = fclamp(f / e, 0.0f, 1.0f); return; }
    float c = dot(d1, r);
    if (e <= EPS) { t = 0.0f; s = fclamp(-c / a, 0.0f, 1.0f); return; }
    float b = dot(d1, d2);
    float denom = a * e - b * b;
    s = denom > 1.0e-6f * a * e ? fclamp((b * f - c * e) / denom, 0.0f, 1.0f) : 0.0f;
    t = (b * s + f) / e;
    if (t < 0.0f) { t = 0.0f; s = fclamp(-c / a, 0.0f, 1.0f); }
    else if (t > 1.0f) { t = 1.0f; s = fclamp((b - c) / a, 0.0f, 1.0f); }
}

void solid_solid(const Solid& A, const Solid& B, const Pose& p12, float prediction, RawManifold& m) {
    const float ra = A.radius, rb = B.radius, eff = prediction + ra + rb;
    const std::vector<V3>& va = A.v;
    std::vector<V3> vb, nb;
    for (const V3& x : B.v) vb.push_back(pose_point(p12, x));
    float best = -F32_MAX;
    V3 bn{0.f, 1.f, 0.f};
    int kind = -1, bi = 0, bj = 0;
    for (int f = 0; f < A.nfaces(); ++f) {
        float s = F32_MAX;
        for (const V3& x : vb) s = fmin2(s, dot(A.n[f], x));
        s = s - A.d[f];
        if (s > eff) return;
        if (s > best) { best = s; bn = A.n[f]; kind = 0; bi = f; }
    }
    for (int f = 0; f < B.nfaces(); ++f) {
        V3 n = qrot(p12.q, B.n[f]);
        nb.push_back(n);
        float d = B.d[f] + dot(n, p12.t);
        float s = F32_MAX;
        for (const V3& x : va) s = fmin2(s, dot(n, x));
        s = s - d;
        if (s > eff) return;
        if (s > best) { best = s; bn = -n; kind = 1; bi = f; }
    }
    float ebest = -F32_MAX;
    V3 en = bn;
    int ei = 0, ej = 0;
    const bool gauss = A.nfaces() > 0 && B.nfaces() > 0;   // Gauss-map pruning of the edge pairs (Gregorius, GDC 2013)
    V3 ca = vzero();
    for (const V3& x : va) ca = ca + x;
    ca = ca * (1.0f / (float)va.size());
    for (int ea = 0; ea < A.nedges(); ++ea) {
        HullEdge e1 = A.edge(ea);
        V3 da = va[e1.v1] - va[e1.v0];
        float la = length_sq(da);
        V3 u1 = vzero(), v1 = vzero();
        if (gauss) { u1 = A.n[e1.f0]; v1 = A.n[e1.f1]; }
        for (int eb = 0; eb < B.nedges(); ++eb) {
            HullEdge e2 = B.edge(eb);
            V3 db = vb[e2.v1] - vb[e2.v0];
            if (gauss) {
                float du2 = dot(nb[e2.f0], da), dv2 = dot(nb[e2.f1], da), du1 = dot(u1, db), dv1 = dot(v1, db);
                if (!(du2 * dv2 < 0.0f && du1 * dv1 < 0.0f && du2 * dv1 < 0.0f)) continue;
            }
            V3 c = cross(da, db);
            float l2 = length_sq(c);
            if (!(l2 > 1.0e-8f * la * length_sq(db))) continue;
            V3 n = c * (1.0f / sqrtf(l2));
            float s;
            if (gauss) {
                if (dot(n, va[e1.v0] - ca) < 0.0f) n = -n;
                s = dot(n, vb[e2.v0] - va[e1.v0]);
            } else {
                float pa = dot(n, va[e1.v0]), pb = dot(n, vb[e2.v0]);
                if (pb < pa) { n = -n; pa = -pa; pb = -pb; }
                float tol = 1.0e-5f * (1.0f + fabsf(pa) + fabsf(pb));
                bool support = true;
                for (size_t i = 0; i < va.size() && support; ++i) support = dot(n, va[i]) <= pa + tol;
                for (size_t i = 0; i < vb.size() && support; ++i) support = dot(n, vb[i]) >= pb - tol;
                if (!support) continue;
                s = pb - pa;
            }
            if (s > eff) return;
            if (s > ebest) { ebest = s; en = n; ei = ea; ej = eb; }
        }
    }
    if (kind < 0 || ebest > best + 1.0e-4f) {
        if (ebest == -F32_MAX) return;
        best = ebest; bn = en; kind = 2; bi = ei; bj = ej;
    }
    V3 n2 = qrot_inv(p12.q, -bn);
    if (kind == 2) {
        HullEdge e1 = A.edge(bi), e2 = B.edge(bj);
        float s, t;
        closest_on_segments(va[e1.v0], va[e1.v1] - va[e1.v0], vb[e2.v0], vb[e2.v1] - vb[e2.v0], s, t);
        V3 qa = va[e1.v0] + (va[e1.v1] - va[e1.v0]) * s, qb = vb[e2.v0] + (vb[e2.v1] - vb[e2.v0]) * t;
        float dist = dot(qb - qa, bn) - ra - rb;
        if (!(dist < prediction)) return;
        push(m, qa + bn * ra, pose_inv_point(p12, qb - bn * rb), FEAT_EDGE | (uint32_t)bi, FEAT_EDGE | (uint32_t)bj, dist);
        m.local_n1 = bn; m.local_n2 = n2;
        return;
    }
    ClipPoint pts[CLIP_MAX];
    int cnt;
    V3 rn;
    float rd;
    if (kind == 0) {
        rn = bn; rd = A.d[bi];
        cnt = clip_incident(A, va, bi, rn, B, vb, nb, pts);
    } else {
        rn = -bn; rd = B.d[bi] + dot(rn, p12.t);
        cnt = clip_incident(B, vb, bi, rn, A, va, A.n, pts);
    }
    int keep[CLIP_MAX], nk = 0;
    for (int k = 0; k < cnt; ++k)
        if (dot(rn, pts[k].p) - rd - ra - rb < prediction) keep[nk++] = k;
    int take = nk < MAX_RAW_POINTS ? nk : MAX_RAW_POINTS;
    for (int k = 0; k < take; ++k) {
        const ClipPoint& c = pts[keep[nk <= MAX_RAW_POINTS ? k : (k * nk) / MAX_RAW_POINTS]];
        float dc = dot(rn, c.p) - rd;
        V3 on_ref = c.p - rn * dc;
        if (kind == 0) push(m, on_ref + bn * ra, pose_inv_point(p12, c.p - bn * rb), FEAT_FACE | (uint32_t)bi, c.id, dc - ra - rb);
        else push(m, c.p + bn * ra, pose_inv_point(p12, on_ref - bn * rb), c.id, FEAT_FACE | (uint32_t)bi, dc - ra - rb);
    }
    m.local_n1 = bn; m.local_n2 = n2;
}

bool solid_ball(const Solid& P, V3 c, float r, float prediction, V3& p_poly, V3& n_poly, float& dist, uint32_t& fid) {
    const float rp = P.radius;
    float smax = -F32_MAX;
    int fm = 0;
    for (int f = 0; f < P.nfaces(); ++f) {
        float s = dot(P.n[f], c) - P.d[f];
        if (s > smax) { smax = s; fm = f; }
    }
    if (smax > prediction + r + rp) return false;
    if (smax <= 0.0f) {
        n_poly = P.n[fm];
        p_poly = c - n_poly * smax + n_poly * rp;
        dist = smax - r - rp;
        fid = FEAT_FACE | (uint32_t)fm;
        return true;
    }
    float bestd = F32_MAX;
    V3 bp = c;
    fid = FEAT_FACE;
    for (int f = 0; f < P.nfaces(); ++f) {
        V3 n = P.n[f];
        float s = dot(n, c) - P.d[f];
        if (!(s > 0.0f)) continue;
        V3 q = c - n * s;
        int fs = P.topo->face_start[f], fn = P.topo->face_count[f];
        bool inside = true;
        for (int k = 0; k < fn; ++k) {
            V3 e0 = P.v[P.topo->loops[fs + k]], e1 = P.v[P.topo->loops[fs + (k + 1 < fn ? k + 1 : 0)]];
            V3 ed = e1 - e0;
            if (dot(cross(ed, n), q - e0) > 0.0f) {
                inside = false;
                float l2 = length_sq(ed);
                float t = l2 > 0.0f ? fclamp(dot(c - e0, ed) / l2, 0.0f, 1.0f) : 0.0f;
                V3 x = e0 + ed * t;
                float d2 = length_sq(c - x);
                if (d2 < bestd) { bestd = d2; bp = x; fid = FEAT_EDGE | ((uint32_t)f << 8) | (uint32_t)k; }
            }
        }
        if (inside && s * s < bestd) { bestd = s * s; bp = q; fid = FEAT_FACE | (uint32_t)f; }
    }
    V3 dl = c - bp;
    float len = length(dl);
    if (!(len > 0.0f)) return false;
    if (!(len - r - rp < prediction)) return false;
    n_poly = dl * (1.0f / len);
    p_poly = bp + n_poly * rp;
    dist = len - r - rp;
    return true;
}
}  // namespace

void contact_manifold_convex(const std::vector<Hull>& hulls, int sh1, V3 he1, int sh2, V3 he2, const Pose& p12, float prediction,
                             RawManifold& m) {
    m.n = 0;
    m.local_n1 = vzero();
    m.local_n2 = vzero();
    if (sh2 == RB_SHAPE_BALL || sh1 == RB_SHAPE_BALL) {
        const bool ball2 = sh2 == RB_SHAPE_BALL;
        const Pose rel = ball2 ? p12 : pose_inverse(p12);   // pose of the ball in the polyhedron's frame
        const Solid P = ball2 ? solid_of(hulls, sh1, he1) : solid_of(hulls, sh2, he2);
        const float r = ball2 ? he2.x : he1.x;
        V3 pp, np;
        float d;
        uint32_t fid;
        if (!solid_ball(P, rel.t, r, prediction, pp, np, d, fid)) return;
        V3 nb = qrot_inv(rel.q, -np);
        if (ball2) { push(m, pp, nb * r, fid, FEAT_FACE, d); m.local_n1 = np; m.local_n2 = nb; }
        else { push(m, nb * r, pp, FEAT_FACE, fid, d); m.local_n1 = nb; m.local_n2 = np; }
        return;
    }
    solid_solid(solid_of(hulls, sh1, he1), solid_of(hulls, sh2, he2), p12, prediction, m);
}

Aabb convex_aabb(const std::vector<Hull>& hulls, V3 he, const Pose& pos) {
    const Hull& h = hulls[(int)he.x];
    V3 lo{F32_MAX, F32_MAX, F32_MAX}, hi{-F32_MAX, -F32_MAX, -F32_MAX};
    for (const V3& v : h.verts) {
        V3 x = qrot(pos.q, v);
        lo = V3{fmin2(lo.x, x.x), fmin2(lo.y, x.y), fmin2(lo.z, x.z)};
        hi = V3{fmax2(hi.x, x.x), fmax2(hi.y, x.y), fmax2(hi.z, x.z)};
    }
    return Aabb{V3{lo.x - he.y, lo.y - he.y, lo.z - he.y} + pos.t, V3{hi.x + he.y, hi.y + he.y, hi.z + he.y} + pos.t};
}

}  // namespace orc
