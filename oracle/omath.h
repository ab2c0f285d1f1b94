// TEST INFRASTRUCTURE ONLY -- CPU oracle math helpers (scalar f32).
//
// This header is part of the parity oracle under oracle/.  Nothing in the product
// (rapier_b200/) includes or links it.  It restates, in plain scalar C++, the f32
// vector / quaternion / pose / symmetric-matrix arithmetic the reference gets from
// glamx 0.3 (scalar Vec3/Quat/Pose, src/lib.rs "math" aliases) and parry's SdpMatrix3
// (src/lib.rs:222).  Those crates are NOT vendored in /root/reference, so the formulas
// below are their published algorithms, written so that every expression has a fixed
// left-to-right evaluation order (the CUDA kernels are compiled with -fmad=false and
// this file with -ffp-contract=off, so both sides round identically).
#pragma once
#include <cmath>
#include <cstdint>
#include <cstring>

namespace orc {

struct V3 {
    float x, y, z;
};
struct Q4 {
    float x, y, z, w;
};
struct Pose {
    Q4 q;
    V3 t;
};
// Symmetric 3x3 (parry SdpMatrix3: m11 m12 m13 m22 m23 m33).
struct Sdp3 {
    float m11, m12, m13, m22, m23, m33;
};

static inline V3 v3(float x, float y, float z) { return V3{x, y, z}; }
static inline V3 vzero() { return V3{0.f, 0.f, 0.f}; }
static inline V3 operator+(V3 a, V3 b) { return V3{a.x + b.x, a.y + b.y, a.z + b.z}; }
static inline V3 operator-(V3 a, V3 b) { return V3{a.x - b.x, a.y - b.y, a.z - b.z}; }
static inline V3 operator-(V3 a) { return V3{-a.x, -a.y, -a.z}; }
static inline V3 operator*(V3 a, float s) { return V3{a.x * s, a.y * s, a.z * s}; }
static inline V3 cmul(V3 a, V3 b) { return V3{a.x * b.x, a.y * b.y, a.z * b.z}; }
// Explicit fused multiply-adds (IEEE correctly rounded, identical on CPU and GPU; the file is built
// with -ffp-contract=off so nothing else is ever contracted).  The placement below defines the
// arithmetic of this engine and is mirrored, expression for expression, by the CUDA kernels.
static inline float fma_(float a, float b, float c) { return fmaf(a, b, c); }
static inline float dot(V3 a, V3 b) { return fma_(a.z, b.z, fma_(a.y, b.y, a.x * b.x)); }
static inline V3 cross(V3 a, V3 b) {
    return V3{fma_(a.y, b.z, -(a.z * b.y)), fma_(a.z, b.x, -(a.x * b.z)), fma_(a.x, b.y, -(a.y * b.x))};
}
static inline V3 madd(V3 v, V3 a, float s) { return V3{fma_(a.x, s, v.x), fma_(a.y, s, v.y), fma_(a.z, s, v.z)}; }        // v + a*s
static inline V3 maddv(V3 v, V3 a, V3 b) { return V3{fma_(a.x, b.x, v.x), fma_(a.y, b.y, v.y), fma_(a.z, b.z, v.z)}; }   // v + a.*b
static inline float length_sq(V3 a) { return dot(a, a); }
static inline float length(V3 a) { return sqrtf(dot(a, a)); }
static inline float vget(V3 a, int i) { return i == 0 ? a.x : (i == 1 ? a.y : a.z); }
static inline void vset(V3& a, int i, float v) {
    if (i == 0) a.x = v; else if (i == 1) a.y = v; else a.z = v;
}
static inline bool veq(V3 a, V3 b) { return a.x == b.x && a.y == b.y && a.z == b.z; }

// src/utils/mod.rs:131,143-146 -- inv(x) = 0 when |x| < 1e-20.
static inline float inv_or_zero(float x) {
    return (x >= -1.0e-20f && x <= 1.0e-20f) ? 0.0f : 1.0f / x;
}

static inline Q4 qidentity() { return Q4{0.f, 0.f, 0.f, 1.f}; }
static inline Q4 qconj(Q4 q) { return Q4{-q.x, -q.y, -q.z, q.w}; }
static inline float qdot(Q4 a, Q4 b) { return fma_(a.w, b.w, fma_(a.z, b.z, fma_(a.y, b.y, a.x * b.x))); }
// Hamilton product a*b (glam Quat::mul_quat, scalar path).
static inline Q4 qmul(Q4 a, Q4 b) {
    return Q4{fma_(-a.z, b.y, fma_(a.y, b.z, fma_(a.x, b.w, a.w * b.x))),
              fma_(a.z, b.x, fma_(a.y, b.w, fma_(-a.x, b.z, a.w * b.y))),
              fma_(a.z, b.w, fma_(-a.y, b.x, fma_(a.x, b.y, a.w * b.z))),
              fma_(-a.z, b.z, fma_(-a.y, b.y, fma_(-a.x, b.x, a.w * b.w)))};
}
static inline Q4 qnormalize(Q4 q) {
    float inv = 1.0f / sqrtf(qdot(q, q));
    return Q4{q.x * inv, q.y * inv, q.z * inv, q.w * inv};
}
// glam Quat::mul_vec3 (scalar path): v*(w^2 - b.b) + b*(2 v.b) + (b x v)*(2w).
static inline V3 qrot(Q4 q, V3 v) {
    V3 b = V3{q.x, q.y, q.z};
    float b2 = dot(b, b);
    return madd(madd(v * fma_(q.w, q.w, -b2), b, dot(v, b) * 2.0f), cross(b, v), q.w * 2.0f);
}
static inline V3 qrot_inv(Q4 q, V3 v) { return qrot(qconj(q), v); }

static inline Pose pose_identity() { return Pose{qidentity(), vzero()}; }
static inline V3 pose_point(const Pose& p, V3 v) { return qrot(p.q, v) + p.t; }
static inline V3 pose_inv_point(const Pose& p, V3 v) { return qrot_inv(p.q, v - p.t); }
static inline Pose pose_mul(const Pose& a, const Pose& b) {
    return Pose{qmul(a.q, b.q), qrot(a.q, b.t) + a.t};
}
// a^-1 * b  (Pose::inv_mul)
static inline Pose pose_inv_mul(const Pose& a, const Pose& b) {
    Q4 ai = qconj(a.q);
    return Pose{qmul(ai, b.q), qrot(ai, b.t - a.t)};
}
static inline Pose pose_inverse(const Pose& a) {
    Q4 ai = qconj(a.q);
    return Pose{ai, qrot(ai, -a.t)};
}
// prepend_translation(v): pose * Translation(v)
static inline Pose pose_prepend_translation(const Pose& p, V3 v) {
    return Pose{p.q, qrot(p.q, v) + p.t};
}

static inline Sdp3 sdp_zero() { return Sdp3{0, 0, 0, 0, 0, 0}; }
static inline V3 sdp_mul(const Sdp3& m, V3 v) {
    return V3{fma_(m.m13, v.z, fma_(m.m12, v.y, m.m11 * v.x)),
              fma_(m.m23, v.z, fma_(m.m22, v.y, m.m12 * v.x)),
              fma_(m.m33, v.z, fma_(m.m23, v.y, m.m13 * v.x))};
}

// Rotation matrix columns of a unit quaternion (c0,c1,c2 = images of x,y,z).
struct M3 {
    V3 c0, c1, c2;
};
static inline M3 qto_mat(Q4 q) {
    float x2 = q.x + q.x, y2 = q.y + q.y, z2 = q.z + q.z;
    float xx = q.x * x2, xy = q.x * y2, xz = q.x * z2;
    float yy = q.y * y2, yz = q.y * z2, zz = q.z * z2;
    float wx = q.w * x2, wy = q.w * y2, wz = q.w * z2;
    M3 m;
    m.c0 = V3{1.0f - (yy + zz), xy + wz, xz - wy};
    m.c1 = V3{xy - wz, 1.0f - (xx + zz), yz + wx};
    m.c2 = V3{xz + wy, yz - wx, 1.0f - (xx + yy)};
    return m;
}

// src/utils/orthonormal_basis.rs:76-93 (Pixar branchless basis).
static inline V3 orthonormal_vector(V3 v) {
    float sign = copysignf(1.0f, v.z);
    float a = -1.0f / (sign + v.z);
    float b = v.x * v.y * a;
    return V3{b, sign + v.y * v.y * a, -v.y};
}
static inline void orthonormal_basis(V3 v, V3& b0, V3& b1) {
    float sign = copysignf(1.0f, v.z);
    float a = -1.0f / (sign + v.z);
    float b = v.x * v.y * a;
    b0 = V3{1.0f + sign * v.x * v.x * a, sign * b, -sign * v.x};
    b1 = V3{b, sign + v.y * v.y * a, -v.y};
}

static inline float fclamp(float x, float lo, float hi) { return x < lo ? lo : (x > hi ? hi : x); }
static inline float fmax2(float a, float b) { return a > b ? a : b; }
static inline float fmin2(float a, float b) { return a < b ? a : b; }

}  // namespace orc
