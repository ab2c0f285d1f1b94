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

void contact_manifold(int shape1, V3 he1, int shape2, V3 he2, const Pose& pos12, float prediction,
                      RawManifold& m) {
    m.n = 0;
    m.local_n1 = vzero();
    m.local_n2 = vzero();
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
    } else {
        M3 r = qto_mat(pos.q);
        ws = V3{fabsf(r.c0.x) * he.x + fabsf(r.c1.x) * he.y + fabsf(r.c2.x) * he.z,
                fabsf(r.c0.y) * he.x + fabsf(r.c1.y) * he.y + fabsf(r.c2.y) * he.z,
                fabsf(r.c0.z) * he.x + fabsf(r.c1.z) * he.y + fabsf(r.c2.z) * he.z};
    }
    return Aabb{pos.t - ws, pos.t + ws};
}

}  // namespace orc
