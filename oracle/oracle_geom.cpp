// TEST INFRASTRUCTURE ONLY -- CPU oracle, geometric queries (see oracle.h header note).
//
// The reference calls `query_dispatcher.contact_manifolds(pos12, shape1, shape2, prediction, ..)`
// (src/geometry/narrow_phase/pair_update.rs:323-330), `Shape::compute_aabb`
// (src/geometry/collider.rs:553-556) and parry's mass properties.  Those live in the un-vendored
// dependency parry3d 0.30.2 (Cargo.toml:72-75, no lockfile) whose source is not under
// /root/reference.  The functions below restate parry's PUBLISHED algorithms for the shapes in
// scope (cuboid, ball):
//   * cuboid-cuboid: SAT over the 3+3 face normals and 9 edge cross products
//     (parry `sat::cuboid_cuboid_find_local_separating_normal_oneway`,
//     `sat::cuboid_cuboid_find_local_separating_edge_twoway`), reference/incident
//     `Cuboid::support_face`, then `PolygonalFeature::contacts` face-face clipping in the plane
//     orthogonal to the separating axis (vertices of one face inside the other + edge/edge crossings);
//   * ball-ball, ball-cuboid (`contact_manifold_ball_ball`, `contact_manifold_convex_ball`).
// PARITY UNPINNED at this level: the reference's tests hold no numeric manifold fixture.
#include "oracle_internal.h"

namespace orc {

static const float F32_EPS = 1.1920929e-7f;

static inline V3 support_point(V3 he, V3 dir) {
    return V3{copysignf(he.x, dir.x), copysignf(he.y, dir.y), copysignf(he.z, dir.z)};
}

// parry sat::cuboid_cuboid_find_local_separating_normal_oneway
static void sat_normal_oneway(V3 he1, V3 he2, const Pose& pos12, float& best_sep, V3& best_dir) {
    best_sep = -3.4028235e38f;
    best_dir = vzero();
    for (int i = 0; i < 3; ++i) {
        float sign = copysignf(1.0f, vget(pos12.t, i));
        V3 axis1 = vzero();
        vset(axis1, i, sign);
        V3 axis2 = qrot_inv(pos12.q, -axis1);
        V3 local_pt2 = support_point(he2, axis2);
        V3 pt2 = pose_point(pos12, local_pt2);
        float separation = vget(pt2, i) * sign - vget(he1, i);
        if (separation > best_sep) {
            best_sep = separation;
            best_dir = axis1;
        }
    }
}

// parry sat::cuboid_support_map_compute_separation_wrt_local_line + ..._find_local_separating_edge_twoway
static void sat_edge_twoway(V3 he1, V3 he2, const Pose& pos12, float& best_sep, V3& best_dir) {
    V3 x2 = qrot(pos12.q, V3{1, 0, 0});
    V3 y2 = qrot(pos12.q, V3{0, 1, 0});
    V3 z2 = qrot(pos12.q, V3{0, 0, 1});
    V3 axes[9] = {
        V3{0.0f, -x2.z, x2.y}, V3{x2.z, 0.0f, -x2.x}, V3{-x2.y, x2.x, 0.0f},
        V3{0.0f, -y2.z, y2.y}, V3{y2.z, 0.0f, -y2.x}, V3{-y2.y, y2.x, 0.0f},
        V3{0.0f, -z2.z, z2.y}, V3{z2.z, 0.0f, -z2.x}, V3{-z2.y, z2.x, 0.0f},
    };
    best_sep = -3.4028235e38f;
    best_dir = vzero();
    for (int k = 0; k < 9; ++k) {
        float n = length(axes[k]);
        if (!(n > F32_EPS)) continue;  // try_normalize(eps)
        V3 axis1 = axes[k] * (1.0f / n);
        float signum = copysignf(1.0f, dot(pos12.t, axis1));
        axis1 = axis1 * signum;
        V3 axis2 = qrot_inv(pos12.q, -axis1);
        V3 local_pt1 = support_point(he1, axis1);
        V3 local_pt2 = support_point(he2, axis2);
        V3 pt2 = pose_point(pos12, local_pt2);
        float separation = dot(pt2 - local_pt1, axis1);
        if (separation > best_sep) {
            best_sep = separation;
            best_dir = axis1;
        }
    }
}

struct Face {
    V3 v[4];
    uint32_t vids[4], eids[4], fid;
};

static inline uint32_t vertex_id(V3 p) {
    return 0x10000000u | ((p.x < 0.0f) ? 1u : 0u) | ((p.y < 0.0f) ? 2u : 0u) | ((p.z < 0.0f) ? 4u : 0u);
}

// parry Cuboid::support_face: the face whose outward normal is the dominant axis of `dir`.
static void support_face(V3 he, V3 dir, Face& f) {
    float ax = fabsf(dir.x), ay = fabsf(dir.y), az = fabsf(dir.z);
    int iamax = 0;
    float amax = ax;
    if (ay > amax) { iamax = 1; amax = ay; }
    if (az > amax) { iamax = 2; amax = az; }
    float sign = copysignf(1.0f, vget(dir, iamax));
    if (iamax == 0) {
        f.v[0] = V3{he.x * sign, he.y, he.z};
        f.v[1] = V3{he.x * sign, -he.y, he.z};
        f.v[2] = V3{he.x * sign, -he.y, -he.z};
        f.v[3] = V3{he.x * sign, he.y, -he.z};
    } else if (iamax == 1) {
        f.v[0] = V3{he.x, he.y * sign, he.z};
        f.v[1] = V3{-he.x, he.y * sign, he.z};
        f.v[2] = V3{-he.x, he.y * sign, -he.z};
        f.v[3] = V3{he.x, he.y * sign, -he.z};
    } else {
        f.v[0] = V3{he.x, he.y, he.z * sign};
        f.v[1] = V3{he.x, -he.y, he.z * sign};
        f.v[2] = V3{-he.x, -he.y, he.z * sign};
        f.v[3] = V3{-he.x, he.y, he.z * sign};
    }
    for (int i = 0; i < 4; ++i) f.vids[i] = vertex_id(f.v[i]);
    for (int i = 0; i < 4; ++i) {
        uint32_t a = f.vids[i] & 7u, b = f.vids[(i + 1) & 3] & 7u;
        uint32_t lo = a < b ? a : b, hi = a < b ? b : a;
        f.eids[i] = 0x30000000u | (lo << 4) | hi;
    }
    f.fid = 0x20000000u | (uint32_t)(iamax + (sign < 0.0f ? 3 : 0));
}

struct P2 {
    float x, y;
};
static inline float perp(P2 a, P2 b) { return a.x * b.y - a.y * b.x; }
static inline P2 sub2(P2 a, P2 b) { return P2{a.x - b.x, a.y - b.y}; }

static bool ulps_eq(float a, float b) {
    float d = a - b;
    if (d < 0) d = -d;
    if (d <= F32_EPS) return true;
    if ((a < 0) != (b < 0)) return false;
    int32_t ia, ib;
    memcpy(&ia, &a, 4);
    memcpy(&ib, &b, 4);
    int32_t diff = ia > ib ? ia - ib : ib - ia;
    return diff <= 4;
}

// parry utils closest_points_line2d (Ericson, Real-Time Collision Detection 5.1.9, 2-D lines).
static bool closest_points_line2d(P2 e1a, P2 e1b, P2 e2a, P2 e2b, float& s, float& t) {
    P2 dir1 = sub2(e1b, e1a), dir2 = sub2(e2b, e2a), r = sub2(e1a, e2a);
    float a = dir1.x * dir1.x + dir1.y * dir1.y;
    float e = dir2.x * dir2.x + dir2.y * dir2.y;
    float f = dir2.x * r.x + dir2.y * r.y;
    if (a <= F32_EPS && e <= F32_EPS) { s = 0; t = 0; return true; }
    if (a <= F32_EPS) { s = 0; t = f / e; return true; }
    float c = dir1.x * r.x + dir1.y * r.y;
    if (e <= F32_EPS) { s = -c / a; t = 0; return true; }
    float b = dir1.x * dir2.x + dir1.y * dir2.y;
    float ae = a * e, bb = b * b, denom = ae - bb;
    bool parallel = denom <= F32_EPS || ulps_eq(ae, bb);
    if (parallel) return false;
    s = (b * f - c * e) / denom;
    t = (b * s + f) / e;
    return true;
}

static inline void push_point(RawManifold& m, V3 p1, V3 p2, uint32_t f1, uint32_t f2, float dist) {
    if (m.n >= MAX_RAW_POINTS) return;  // two convex quads: <= 8 (cap documented in DESIGN.md)
    RawPoint& p = m.pts[m.n++];
    p.local_p1 = p1; p.local_p2 = p2; p.fid1 = f1; p.fid2 = f2; p.dist = dist == 0.0f ? 0.0f : dist;  // canonical zero
}

// parry PolygonalFeature::contacts_face_face (3-D, both features are 4-vertex faces).
static void contacts_face_face(const Pose& pos12, const Face& face1, V3 sep_axis1, const Face& face2,
                               RawManifold& m) {
    V3 b0, b1;
    orthonormal_basis(sep_axis1, b0, b1);
    P2 pf1[4], pf2[4];
    V3 v2_1[4];
    for (int i = 0; i < 4; ++i) {
        pf1[i] = P2{dot(face1.v[i], b0), dot(face1.v[i], b1)};
        v2_1[i] = pose_point(pos12, face2.v[i]);
        pf2[i] = P2{dot(v2_1[i], b0), dot(v2_1[i], b1)};
    }
    // Vertices of face1 inside the projection of face2.
    {
        V3 normal2_1 = cross(v2_1[2] - v2_1[1], v2_1[0] - v2_1[1]);
        float denom = dot(normal2_1, sep_axis1);
        if (fabsf(denom) > F32_EPS) {
            for (int i = 0; i < 4; ++i) {
                P2 p1 = pf1[i];
                float sign = perp(sub2(pf2[0], pf2[3]), sub2(p1, pf2[3]));
                bool outside = false;
                for (int j = 0; j < 3; ++j) {
                    float new_sign = perp(sub2(pf2[j + 1], pf2[j]), sub2(p1, pf2[j]));
                    if (new_sign * sign < 0.0f) { outside = true; break; }
                }
                if (outside) continue;
                float dist = dot(v2_1[0] - face1.v[i], normal2_1) / denom;
                V3 local_p1 = face1.v[i];
                V3 local_p2 = pose_inv_point(pos12, face1.v[i] + sep_axis1 * dist);
                push_point(m, local_p1, local_p2, face1.vids[i], face2.fid, dist);
            }
        }
    }
    // Vertices of face2 inside the projection of face1.
    {
        V3 normal1 = cross(face1.v[2] - face1.v[1], face1.v[0] - face1.v[1]);
        float denom = -dot(normal1, sep_axis1);
        if (fabsf(denom) > F32_EPS) {
            for (int i = 0; i < 4; ++i) {
                P2 p2 = pf2[i];
                float sign = perp(sub2(pf1[0], pf1[3]), sub2(p2, pf1[3]));
                bool outside = false;
                for (int j = 0; j < 3; ++j) {
                    float new_sign = perp(sub2(pf1[j + 1], pf1[j]), sub2(p2, pf1[j]));
                    if (new_sign * sign < 0.0f) { outside = true; break; }
                }
                if (outside) continue;
                float dist = dot(face1.v[0] - v2_1[i], normal1) / denom;
                V3 local_p2 = face2.v[i];
                V3 local_p1 = v2_1[i] - sep_axis1 * dist;
                push_point(m, local_p1, local_p2, face1.fid, face2.vids[i], dist);
            }
        }
    }
    // Edge/edge crossings.
    for (int j = 0; j < 4; ++j) {
        P2 e2a = pf2[j], e2b = pf2[(j + 1) & 3];
        for (int i = 0; i < 4; ++i) {
            P2 e1a = pf1[i], e1b = pf1[(i + 1) & 3];
            float s, t;
            if (!closest_points_line2d(e1a, e1b, e2a, e2b, s, t)) continue;
            if (s > 0.0f && s < 1.0f && t > 0.0f && t < 1.0f) {
                V3 a0 = face1.v[i], a1 = face1.v[(i + 1) & 3];
                V3 c0 = v2_1[j], c1 = v2_1[(j + 1) & 3];
                V3 local_p1 = a0 * (1.0f - s) + a1 * s;
                V3 local_p2_1 = c0 * (1.0f - t) + c1 * t;
                float dist = dot(local_p2_1 - local_p1, sep_axis1);
                V3 local_p2 = pose_inv_point(pos12, local_p2_1);
                push_point(m, local_p1, local_p2, face1.eids[i], face2.eids[j], dist);
            }
        }
    }
}

// parry contact_manifold_cuboid_cuboid (without the `try_update_contacts` spatial-coherence
// shortcut: rapier's own contact recycling (pair_update.rs:111-171) already sits in front of it;
// see DESIGN.md "deviations").
static void manifold_cuboid_cuboid(V3 he1, V3 he2, const Pose& pos12, float prediction, RawManifold& m) {
    m.n = 0;
    m.local_n1 = vzero();
    m.local_n2 = vzero();
    Pose pos21 = pose_inverse(pos12);
    float sep1, sep2, sep3;
    V3 dir1, dir2, dir3;
    sat_normal_oneway(he1, he2, pos12, sep1, dir1);
    if (sep1 > prediction) return;
    sat_normal_oneway(he2, he1, pos21, sep2, dir2);
    if (sep2 > prediction) return;
    sat_edge_twoway(he1, he2, pos12, sep3, dir3);
    if (sep3 > prediction) return;
    float best_sep = sep1;
    V3 best_dir = dir1;
    if (sep2 > sep1 && sep2 > sep3) {
        best_sep = sep2;
        best_dir = qrot(pos12.q, -dir2);
    } else if (sep3 > sep1) {
        best_sep = sep3;
        best_dir = dir3;
    }
    (void)best_sep;
    V3 local_n2 = qrot(pos21.q, -best_dir);
    Face f1, f2;
    support_face(he1, best_dir, f1);
    support_face(he2, local_n2, f2);
    contacts_face_face(pos12, f1, best_dir, f2, m);
    m.local_n1 = best_dir;
    m.local_n2 = local_n2;
}

// parry contact_manifold_ball_ball
static void manifold_ball_ball(float r1, float r2, const Pose& pos12, float prediction, RawManifold& m) {
    m.n = 0;
    m.local_n1 = vzero();
    m.local_n2 = vzero();
    V3 dcenter = pos12.t;
    float center_dist = length(dcenter);
    float dist = center_dist - r1 - r2;
    if (!(dist < prediction)) return;
    V3 n1 = center_dist != 0.0f ? dcenter * (1.0f / center_dist) : V3{0, 1, 0};
    V3 n2 = qrot_inv(pos12.q, -n1);
    push_point(m, n1 * r1, n2 * r2, 0x20000000u, 0x20000000u, dist);
    m.local_n1 = n1;
    m.local_n2 = n2;
}

// parry Cuboid::project_local_point_and_get_feature + contact_manifold_convex_ball.
// `posb_c`: pose of the ball in the cuboid frame.  Outputs in (cuboid, ball) order.
static bool cuboid_ball(V3 he, float r, const Pose& posb_c, float prediction, V3& p_cuboid, V3& p_ball,
                        V3& n_cuboid, V3& n_ball, float& dist_out, uint32_t& fid) {
    V3 c = posb_c.t;
    V3 mins_pt = V3{-he.x - c.x, -he.y - c.y, -he.z - c.z};
    V3 pt_maxs = V3{c.x - he.x, c.y - he.y, c.z - he.z};
    V3 shift = V3{fmax2(mins_pt.x, 0.0f) - fmax2(pt_maxs.x, 0.0f), fmax2(mins_pt.y, 0.0f) - fmax2(pt_maxs.y, 0.0f),
                  fmax2(mins_pt.z, 0.0f) - fmax2(pt_maxs.z, 0.0f)};
    bool inside = shift.x == 0.0f && shift.y == 0.0f && shift.z == 0.0f;
    V3 proj;
    if (!inside) {
        proj = c + shift;
        fid = 0x20000000u;
    } else {
        // closest face: the largest (least negative) of mins_pt / pt_maxs.
        float best = -3.4028235e38f;
        int bi = 0;
        float bs = 1.0f;
        for (int i = 0; i < 3; ++i) {
            if (vget(pt_maxs, i) > best) { best = vget(pt_maxs, i); bi = i; bs = 1.0f; }
            if (vget(mins_pt, i) > best) { best = vget(mins_pt, i); bi = i; bs = -1.0f; }
        }
        proj = c;
        vset(proj, bi, bs * vget(he, bi));
        fid = 0x20000000u | (uint32_t)(bi + (bs < 0 ? 3 : 0));
    }
    V3 dpos = c - proj;
    float d = length(dpos);
    if (!(d > 0.0f)) return false;  // Unit::try_new_and_get(dpos, 0.0)
    V3 n1 = dpos * (1.0f / d);
    if (inside) { n1 = -n1; d = -d; }
    if (!(d <= r + prediction)) return false;
    V3 n2 = qrot_inv(posb_c.q, -n1);
    p_cuboid = proj;
    p_ball = n2 * r;
    n_cuboid = n1;
    n_ball = n2;
    dist_out = d - r;
    return true;
}


// ---------------------------------------------------------------------------------------------
// Capsules (parry Capsule: segment + radius; he = (half height, radius, axis)).  PARITY UNPINNED like the cuboid
// manifolds: parry's contact_manifold_capsule_capsule / _cuboid_capsule / _ball_convex are not in the tree, so the
// manifolds are written from first principles (closest features of the segment cores, points on the surfaces).
// ---------------------------------------------------------------------------------------------
static inline V3 capsule_dir(V3 he) { return he.z == 0.0f ? V3{1.f, 0.f, 0.f} : (he.z == 1.0f ? V3{0.f, 1.f, 0.f} : V3{0.f, 0.f, 1.f}); }
static inline V3 vwith(V3 v, int i, float x) { vset(v, i, x); return v; }

// closest points of two segments (Ericson, Real-Time Collision Detection, 5.1.9)
static void seg_seg_params(V3 p1, V3 d1, V3 p2, V3 d2, float& s, float& t) {
    const V3 r = p1 - p2;
    const float a = dot(d1, d1), e = dot(d2, d2), f = dot(d2, r);
    if (a <= F32_EPS && e <= F32_EPS) { s = 0.0f; t = 0.0f; return; }
    if (a <= F32_EPS) { s = 0.0f; t = fclamp(f / e, 0.0f, 1.0f); return; }
    const float c = dot(d1, r);
    if (e <= F32_EPS) { t = 0.0f; s = fclamp(-c / a, 0.0f, 1.0f); return; }
    const float b = dot(d1, d2);
    const float denom = a * e - b * b;
    s = denom > 1.0e-6f * a * e ? fclamp((b * f - c * e) / denom, 0.0f, 1.0f) : 0.0f;
    t = (b * s + f) / e;
    if (t < 0.0f) { t = 0.0f; s = fclamp(-c / a, 0.0f, 1.0f); }
    else if (t > 1.0f) { t = 1.0f; s = fclamp((b - c) / a, 0.0f, 1.0f); }
}

static void manifold_capsule_capsule(V3 he1, V3 he2, const Pose& p12, float prediction, RawManifold& m) {
    m.n = 0; m.local_n1 = vzero(); m.local_n2 = vzero();
    const float hh1 = he1.x, r1 = he1.y, hh2 = he2.x, r2 = he2.y;
    const V3 u1 = capsule_dir(he1), u2 = qrot(p12.q, capsule_dir(he2));
    const V3 a1 = u1 * (-hh1), d1 = u1 * (2.0f * hh1);
    const V3 a2 = p12.t - u2 * hh2, d2 = u2 * (2.0f * hh2);
    const V3 cr = cross(d1, d2);
    const float l1 = length_sq(d1), l2 = length_sq(d2);
    if (l1 > F32_EPS && l2 > F32_EPS && length_sq(cr) <= 1.0e-6f * l1 * l2) {
        const float inv = 1.0f / l1;
        const float ta = dot(a2 - a1, d1) * inv, tb = dot((a2 + d2) - a1, d1) * inv;
        const float lo = fmax2(fmin2(ta, tb), 0.0f), hi = fmin2(fmax2(ta, tb), 1.0f);
        if (hi > lo) {
            const V3 w0 = a2 - a1;
            const V3 wv = w0 - d1 * (dot(w0, d1) * inv);
            const float wl = length(wv);
            const V3 n1 = wl > F32_EPS ? wv * (1.0f / wl) : orthonormal_vector(u1);
            const float dist = wl - r1 - r2;
            if (!(dist < prediction)) return;
            const float ts[2] = {lo, hi};
            for (int k = 0; k < 2; ++k) {
                const V3 q1 = a1 + d1 * ts[k];
                const V3 q2 = q1 + n1 * wl;
                push_point(m, q1 + n1 * r1, pose_inv_point(p12, q2 - n1 * r2), 0x10000000u | (uint32_t)k, 0x10000000u | (uint32_t)k, dist);
            }
            m.local_n1 = n1;
            m.local_n2 = qrot_inv(p12.q, -n1);
            return;
        }
    }
    float s, t;
    seg_seg_params(a1, d1, a2, d2, s, t);
    const V3 q1 = a1 + d1 * s, q2 = a2 + d2 * t;
    const V3 dl = q2 - q1;
    const float len = length(dl);
    V3 n1;
    if (len > F32_EPS) n1 = dl * (1.0f / len);
    else { const float cl = length(cr); n1 = cl > F32_EPS ? cr * (1.0f / cl) : orthonormal_vector(u1); }
    const float dist = len - r1 - r2;
    if (!(dist < prediction)) return;
    push_point(m, q1 + n1 * r1, pose_inv_point(p12, q2 - n1 * r2), 0x30000000u | 2u, 0x30000000u | 2u, dist);
    m.local_n1 = n1;
    m.local_n2 = qrot_inv(p12.q, -n1);
}

static bool capsule_ball(V3 hec, float rb, const Pose& pb, float prediction, V3& p_cap, V3& p_ball, V3& n_cap, V3& n_ball, float& dist) {
    const float hh = hec.x, rc = hec.y;
    const V3 u = capsule_dir(hec);
    const V3 a = u * (-hh), d = u * (2.0f * hh);
    const float l2 = length_sq(d);
    const float t = l2 > F32_EPS ? fclamp(dot(pb.t - a, d) / l2, 0.0f, 1.0f) : 0.0f;
    const V3 q = a + d * t;
    const V3 dl = pb.t - q;
    const float len = length(dl);
    const V3 n = len > F32_EPS ? dl * (1.0f / len) : orthonormal_vector(u);
    dist = len - rc - rb;
    if (!(dist < prediction)) return false;
    n_cap = n; n_ball = qrot_inv(pb.q, -n);
    p_cap = q + n * rc; p_ball = n_ball * rb;
    return true;
}

static void cuboid_capsule(V3 he, V3 hec, const Pose& pc, float prediction, RawManifold& m) {
    m.n = 0; m.local_n1 = vzero(); m.local_n2 = vzero();
    const float hh = hec.x, r = hec.y;
    const V3 u = qrot(pc.q, capsule_dir(hec));
    const V3 A = pc.t - u * hh, B = pc.t + u * hh;
    float best = -3.4028235e38f;
    V3 n = V3{0.f, 1.f, 0.f};
    int kind = 0, bi = 0;
    for (int i = 0; i < 3; ++i) {
        const float ai = vget(A, i), bb = vget(B, i), h = vget(he, i);
        const float sp = fmin2(ai, bb) - h, sm = -fmax2(ai, bb) - h;
        if (sp > best) { best = sp; n = vwith(vzero(), i, 1.0f); kind = 0; bi = i; }
        if (sm > best) { best = sm; n = vwith(vzero(), i, -1.0f); kind = 0; bi = i; }
    }
    for (int i = 0; i < 3; ++i) {
        const V3 c = cross(vwith(vzero(), i, 1.0f), u);
        const float l2 = length_sq(c);
        if (!(l2 > 1.0e-6f)) continue;
        V3 nn = c * (1.0f / sqrtf(l2));
        float s0 = dot(nn, A);
        if (s0 < 0.0f) { nn = -nn; s0 = -s0; }
        const float sep = s0 - fma_(he.z, fabsf(nn.z), fma_(he.y, fabsf(nn.y), he.x * fabsf(nn.x)));
        if (sep > best) { best = sep; n = nn; kind = 1; bi = i; }
    }
    if (!(best - r < prediction)) return;
    const V3 D = B - A;
    if (kind == 0) {
        const float sg = vget(n, bi);
        float t0 = 0.0f, t1 = 1.0f;
        bool miss = false;
        for (int j = 0; j < 3; ++j) {
            if (j == bi) continue;
            const float aj = vget(A, j), dj = vget(D, j), h = vget(he, j);
            if (fabsf(dj) <= F32_EPS) { if (fabsf(aj) > h) miss = true; continue; }
            float ta = (-h - aj) / dj, tb = (h - aj) / dj;
            if (ta > tb) { const float tmp = ta; ta = tb; tb = tmp; }
            t0 = fmax2(t0, ta); t1 = fmin2(t1, tb);
        }
        if (!miss && t0 <= t1) {
            const float ts[2] = {t0, t1};
            const int np = (t1 - t0 > 1.0e-5f) ? 2 : 1;
            for (int k = 0; k < np; ++k) {
                const V3 q = A + D * ts[k];
                const float dp = sg * vget(q, bi) - vget(he, bi);
                const float dist = dp - r;
                if (!(dist < prediction)) continue;
                push_point(m, q - n * dp, pose_inv_point(pc, q - n * r), 0x20000000u | (uint32_t)(bi + (sg < 0.0f ? 3 : 0)), 0x10000000u | (uint32_t)k, dist);
            }
            m.local_n1 = n;
            m.local_n2 = qrot_inv(pc.q, -n);
            return;
        }
        const V3 q = (sg * vget(A, bi) <= sg * vget(B, bi)) ? A : B;
        const V3 pb = V3{fclamp(q.x, -he.x, he.x), fclamp(q.y, -he.y, he.y), fclamp(q.z, -he.z, he.z)};
        const V3 dl = q - pb;
        const float len = length(dl);
        const V3 nn = len > F32_EPS ? dl * (1.0f / len) : n;
        const float dist = len - r;
        if (!(dist < prediction)) return;
        push_point(m, pb, pose_inv_point(pc, q - nn * r), vertex_id(pb), 0x10000000u | 3u, dist);
        m.local_n1 = nn;
        m.local_n2 = qrot_inv(pc.q, -nn);
        return;
    }
    V3 e0 = support_point(he, n), ed = vzero();
    vset(e0, bi, -vget(he, bi));
    vset(ed, bi, 2.0f * vget(he, bi));
    float s, t;
    seg_seg_params(e0, ed, A, D, s, t);
    const V3 pe = e0 + ed * s, q = A + D * t;
    const float dist = dot(q - pe, n) - r;
    if (!(dist < prediction)) return;
    push_point(m, pe, pose_inv_point(pc, q - n * r), 0x30000000u | ((uint32_t)bi << 4) | (vertex_id(e0) & 7u), 0x30000000u | 2u, dist);
    m.local_n1 = n;
    m.local_n2 = qrot_inv(pc.q, -n);
}

static void manifold_flip(const RawManifold& a, RawManifold& m) {
    m.n = a.n; m.local_n1 = a.local_n2; m.local_n2 = a.local_n1;
    for (int i = 0; i < a.n; ++i) {
        m.pts[i].local_p1 = a.pts[i].local_p2; m.pts[i].local_p2 = a.pts[i].local_p1; m.pts[i].dist = a.pts[i].dist;
        m.pts[i].fid1 = a.pts[i].fid2; m.pts[i].fid2 = a.pts[i].fid1;
    }
}
static void contact_manifold_capsules(int sh1, V3 he1, int sh2, V3 he2, const Pose& p12, float prediction, RawManifold& m) {
    m.n = 0; m.local_n1 = vzero(); m.local_n2 = vzero();
    if (sh1 == RB_SHAPE_CAPSULE && sh2 == RB_SHAPE_CAPSULE) { manifold_capsule_capsule(he1, he2, p12, prediction, m); return; }
    if (sh1 == RB_SHAPE_CUBOID) { cuboid_capsule(he1, he2, p12, prediction, m); return; }
    if (sh2 == RB_SHAPE_CUBOID) {
        RawManifold t;
        cuboid_capsule(he2, he1, pose_inverse(p12), prediction, t);
        manifold_flip(t, m);
        return;
    }
    V3 pc, pb, nc, nb;
    float d;
    if (sh1 == RB_SHAPE_CAPSULE) {
        if (capsule_ball(he1, he2.x, p12, prediction, pc, pb, nc, nb, d)) { push_point(m, pc, pb, 0x30000000u | 2u, 0x20000000u, d); m.local_n1 = nc; m.local_n2 = nb; }
    } else {
        if (capsule_ball(he2, he1.x, pose_inverse(p12), prediction, pc, pb, nc, nb, d)) { push_point(m, pb, pc, 0x20000000u, 0x30000000u | 2u, d); m.local_n1 = nb; m.local_n2 = nc; }
    }
}

void contact_manifold(int shape1, V3 he1, int shape2, V3 he2, const Pose& pos12, float prediction,
                      RawManifold& m) {
    m.n = 0;
    m.local_n1 = vzero();
    m.local_n2 = vzero();
    if (shape1 == RB_SHAPE_CAPSULE || shape2 == RB_SHAPE_CAPSULE) { contact_manifold_capsules(shape1, he1, shape2, he2, pos12, prediction, m); return; }
    if (shape1 == RB_SHAPE_CUBOID && shape2 == RB_SHAPE_CUBOID) {
        manifold_cuboid_cuboid(he1, he2, pos12, prediction, m);
    } else if (shape1 == RB_SHAPE_BALL && shape2 == RB_SHAPE_BALL) {
        manifold_ball_ball(he1.x, he2.x, pos12, prediction, m);
    } else if (shape1 == RB_SHAPE_CUBOID && shape2 == RB_SHAPE_BALL) {
        V3 pc, pb, nc, nb;
        float d;
        uint32_t fid;
        if (cuboid_ball(he1, he2.x, pos12, prediction, pc, pb, nc, nb, d, fid)) {
            push_point(m, pc, pb, fid, 0x20000000u, d);
            m.local_n1 = nc;
            m.local_n2 = nb;
        }
    } else {  // ball (1) vs cuboid (2): flipped
        Pose pos21 = pose_inverse(pos12);
        V3 pc, pb, nc, nb;
        float d;
        uint32_t fid;
        if (cuboid_ball(he2, he1.x, pos21, prediction, pc, pb, nc, nb, d, fid)) {
            push_point(m, pb, pc, 0x20000000u, fid, d);
            m.local_n1 = nb;
            m.local_n2 = nc;
        }
    }
}

// parry Cuboid::aabb / Ball::aabb (Shape::compute_aabb, collider.rs:553-556).
Aabb shape_aabb(int shape, V3 he, const Pose& pos) {
    V3 ws;
    if (shape == RB_SHAPE_BALL) {
        ws = V3{he.x, he.x, he.x};
    } else if (shape == RB_SHAPE_CAPSULE) {
        const V3 u = qrot(pos.q, capsule_dir(he));
        ws = V3{fabsf(u.x) * he.x + he.y, fabsf(u.y) * he.x + he.y, fabsf(u.z) * he.x + he.y};
    } else {
        M3 r = qto_mat(pos.q);
        ws = V3{fabsf(r.c0.x) * he.x + fabsf(r.c1.x) * he.y + fabsf(r.c2.x) * he.z,
                fabsf(r.c0.y) * he.x + fabsf(r.c1.y) * he.y + fabsf(r.c2.y) * he.z,
                fabsf(r.c0.z) * he.x + fabsf(r.c1.z) * he.y + fabsf(r.c2.z) * he.z};
    }
    return Aabb{pos.t - ws, pos.t + ws};
}

}  // namespace orc
