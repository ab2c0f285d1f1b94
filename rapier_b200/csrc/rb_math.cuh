// rb_math.cuh -- f32 vector / quaternion / pose / symmetric 3x3 helpers for the kernels.
//
// Every expression has a fixed left-to-right association and the library is compiled with
// -fmad=false, IEEE division and square root, so results are reproducible bit for bit (and equal
// to the scalar oracle's, which is compiled with -ffp-contract=off).  The formulas are the
// published ones of the reference's math layer (glam scalar Quat/Vec3 as aliased in
// src/lib.rs "math"; parry SdpMatrix3, src/lib.rs:222; Pixar orthonormal basis,
// src/utils/orthonormal_basis.rs:76-93; simd_inv, src/utils/mod.rs:143-146).
#pragma once
#include "rb_common.cuh"

namespace rb {

struct vec3 { float x, y, z; };
struct quat { float x, y, z, w; };
struct pose { quat q; vec3 t; };
struct sym3 { float xx, xy, xz, yy, yz, zz; };

RB_HD vec3 mk3(float x, float y, float z) { vec3 v; v.x = x; v.y = y; v.z = z; return v; }
RB_HD vec3 xyz(float4 f) { return mk3(f.x, f.y, f.z); }
RB_HD quat mkq(float4 f) { quat q; q.x = f.x; q.y = f.y; q.z = f.z; q.w = f.w; return q; }
RB_HD float4 f4(vec3 v, float w) { return make_float4(v.x, v.y, v.z, w); }
RB_HD float4 f4(quat q) { return make_float4(q.x, q.y, q.z, q.w); }
RB_HD vec3 zero3() { return mk3(0.f, 0.f, 0.f); }
RB_HD vec3 operator+(vec3 a, vec3 b) { return mk3(a.x + b.x, a.y + b.y, a.z + b.z); }
RB_HD vec3 operator-(vec3 a, vec3 b) { return mk3(a.x - b.x, a.y - b.y, a.z - b.z); }
RB_HD vec3 operator-(vec3 a) { return mk3(-a.x, -a.y, -a.z); }
RB_HD vec3 operator*(vec3 a, float s) { return mk3(a.x * s, a.y * s, a.z * s); }
RB_HD vec3 had(vec3 a, vec3 b) { return mk3(a.x * b.x, a.y * b.y, a.z * b.z); }
// Fused multiply-adds are written out explicitly (the library is compiled with -fmad=false, so the
// compiler never contracts on its own): fmaf is correctly rounded on the GPU and on the host, which
// keeps the kernels and the CPU oracle bit-identical while halving the FP instruction count.
RB_HD float fma_(float a, float b, float c) { return fmaf(a, b, c); }
RB_HD float dot3(vec3 a, vec3 b) { return fma_(a.z, b.z, fma_(a.y, b.y, a.x * b.x)); }
RB_HD vec3 cross3(vec3 a, vec3 b) {
    return mk3(fma_(a.y, b.z, -(a.z * b.y)), fma_(a.z, b.x, -(a.x * b.z)), fma_(a.x, b.y, -(a.y * b.x)));
}
RB_HD vec3 madd3(vec3 v, vec3 a, float s) { return mk3(fma_(a.x, s, v.x), fma_(a.y, s, v.y), fma_(a.z, s, v.z)); }          // v + a*s
RB_HD vec3 madd3v(vec3 v, vec3 a, vec3 b) { return mk3(fma_(a.x, b.x, v.x), fma_(a.y, b.y, v.y), fma_(a.z, b.z, v.z)); }    // v + a.*b
RB_HD float norm2(vec3 a) { return dot3(a, a); }
RB_HD float norm(vec3 a) { return sqrtf(dot3(a, a)); }
RB_HD float comp(vec3 a, int i) { return i == 0 ? a.x : (i == 1 ? a.y : a.z); }
RB_HD vec3 with_comp(vec3 a, int i, float v) { if (i == 0) a.x = v; else if (i == 1) a.y = v; else a.z = v; return a; }

RB_HD bool finite3(vec3 a) { return isfinite(a.x) && isfinite(a.y) && isfinite(a.z); }
RB_HD float safe_inv(float x) { return (x >= -1.0e-20f && x <= 1.0e-20f) ? 0.0f : 1.0f / x; }
RB_HD float inv_exact0(float x) { return x == 0.0f ? 0.0f : 1.0f / x; }
RB_HD float clampf(float x, float lo, float hi) { return x < lo ? lo : (x > hi ? hi : x); }
RB_HD int min2i(int a, int b) { return a < b ? a : b; }
RB_HD int max2i(int a, int b) { return a > b ? a : b; }
RB_HD float max2(float a, float b) { return a > b ? a : b; }
RB_HD float min2(float a, float b) { return a < b ? a : b; }
RB_HD float copysign1(float s) { return copysignf(1.0f, s); }

RB_HD quat qident() { quat q; q.x = 0.f; q.y = 0.f; q.z = 0.f; q.w = 1.f; return q; }
RB_HD quat qconj(quat q) { quat r; r.x = -q.x; r.y = -q.y; r.z = -q.z; r.w = q.w; return r; }
RB_HD float qdot(quat a, quat b) { return fma_(a.w, b.w, fma_(a.z, b.z, fma_(a.y, b.y, a.x * b.x))); }
RB_HD quat qmul(quat a, quat b) {
    quat r;
    r.x = fma_(-a.z, b.y, fma_(a.y, b.z, fma_(a.x, b.w, a.w * b.x)));
    r.y = fma_(a.z, b.x, fma_(a.y, b.w, fma_(-a.x, b.z, a.w * b.y)));
    r.z = fma_(a.z, b.w, fma_(-a.y, b.x, fma_(a.x, b.y, a.w * b.z)));
    r.w = fma_(-a.z, b.z, fma_(-a.y, b.y, fma_(-a.x, b.x, a.w * b.w)));
    return r;
}
RB_HD quat qnormalize(quat q) {
    float inv = 1.0f / sqrtf(qdot(q, q));
    quat r; r.x = q.x * inv; r.y = q.y * inv; r.z = q.z * inv; r.w = q.w * inv;
    return r;
}
RB_HD vec3 rotate(quat q, vec3 v) {
    vec3 b = mk3(q.x, q.y, q.z);
    float b2 = dot3(b, b);
    return madd3(madd3(v * fma_(q.w, q.w, -b2), b, dot3(v, b) * 2.0f), cross3(b, v), q.w * 2.0f);
}
RB_HD vec3 rotate_inv(quat q, vec3 v) { return rotate(qconj(q), v); }

RB_HD pose pident() { pose p; p.q = qident(); p.t = zero3(); return p; }
RB_HD pose mkpose(quat q, vec3 t) { pose p; p.q = q; p.t = t; return p; }
RB_HD vec3 xform(const pose& p, vec3 v) { return rotate(p.q, v) + p.t; }
RB_HD vec3 xform_inv(const pose& p, vec3 v) { return rotate_inv(p.q, v - p.t); }
RB_HD pose pmul(const pose& a, const pose& b) { return mkpose(qmul(a.q, b.q), rotate(a.q, b.t) + a.t); }
RB_HD pose pinv_mul(const pose& a, const pose& b) {
    quat ai = qconj(a.q);
    return mkpose(qmul(ai, b.q), rotate(ai, b.t - a.t));
}
RB_HD pose pinverse(const pose& a) {
    quat ai = qconj(a.q);
    return mkpose(ai, rotate(ai, -a.t));
}
RB_HD pose prepend_translation(const pose& p, vec3 v) { return mkpose(p.q, rotate(p.q, v) + p.t); }

RB_HD sym3 sym_zero() { sym3 m; m.xx = m.xy = m.xz = m.yy = m.yz = m.zz = 0.f; return m; }
RB_HD vec3 smul(const sym3& m, vec3 v) {
    return mk3(fma_(m.xz, v.z, fma_(m.xy, v.y, m.xx * v.x)), fma_(m.yz, v.z, fma_(m.yy, v.y, m.xy * v.x)),
               fma_(m.zz, v.z, fma_(m.yz, v.y, m.xz * v.x)));
}

struct mat3 { vec3 c0, c1, c2; };
RB_HD mat3 rotmat(quat q) {
    float x2 = q.x + q.x, y2 = q.y + q.y, z2 = q.z + q.z;
    float xx = q.x * x2, xy = q.x * y2, xz = q.x * z2;
    float yy = q.y * y2, yz = q.y * z2, zz = q.z * z2;
    float wx = q.w * x2, wy = q.w * y2, wz = q.w * z2;
    mat3 m;
    m.c0 = mk3(1.0f - (yy + zz), xy + wz, xz - wy);
    m.c1 = mk3(xy - wz, 1.0f - (xx + zz), yz + wx);
    m.c2 = mk3(xz + wy, yz - wx, 1.0f - (xx + yy));
    return m;
}

RB_HD vec3 ortho_vector(vec3 v) {
    float sign = copysign1(v.z);
    float a = -1.0f / (sign + v.z);
    float b = v.x * v.y * a;
    return mk3(b, sign + v.y * v.y * a, -v.y);
}
RB_HD void ortho_basis(vec3 v, vec3& b0, vec3& b1) {
    float sign = copysign1(v.z);
    float a = -1.0f / (sign + v.z);
    float b = v.x * v.y * a;
    b0 = mk3(1.0f + sign * v.x * v.x * a, sign * b, -sign * v.x);
    b1 = mk3(b, sign + v.y * v.y * a, -v.y);
}

}  // namespace rb
