// HALO / A-NeRF per-bone point embedding (utils/fields.py:22-52, 142-148) and its derivatives.
// Per (point, joint):  q = R x + t - T,  v = |q|,  r = q / v,  h = 1 - sigmoid(200 (v - cutoff)),
//   F(q) = h * [v, sin(2^k v), cos(2^k v) (k<10), r (3), sin(2^k r_a), cos(2^k r_a) (k<7, a<3)]   (66 values)
// For a cotangent c[66]: s(q) = <c, F(q)> = h(v) (A(v) + B(r));  the kernels need
//   grad  g  = ds/dq, the JVP F'(q) w, and the Hessian-vector product d(g.w)/dq
// (formulas validated against autograd in oracle/analytic.py: halo_grad_hvp).
#pragma once
#include "common.cuh"

namespace hn {

constexpr int HALO_J = 21, HALO_F = 66, HALO_DIM = HALO_J * HALO_F;   // 1386
constexpr int HALO_LV = 10, HALO_LR = 7;
constexpr float HALO_TAU = 200.0f;

__constant__ float c_halo_cutoff[HALO_J] = {0.08f, 0.03f, 0.03f, 0.02f, 0.02f, 0.03f, 0.02f, 0.02f, 0.02f, 0.03f, 0.02f,
                                            0.02f, 0.02f, 0.03f, 0.02f, 0.02f, 0.02f, 0.03f, 0.02f, 0.02f, 0.02f};

// sin / cos of the encoding arguments (|x| < ~800) for the DERIVATIVE kernels: two-term Cody-Waite reduction to [-pi, pi],
// then the SFU approximations (abs error ~4e-7; the feature itself keeps sincosf, its values enter the SDF directly)
__device__ __forceinline__ void halo_sincos_fast(float x, float* s, float* c) {
    const float n = rintf(x * 0.15915494309189535f);
    float r = fmaf(-n, 6.2831854820251465f, x);
    r = fmaf(-n, -1.7484556000744883e-07f, r);
    *s = __sinf(r);
    *c = __cosf(r);
}

struct HaloBase {
    float q[3], r[3], v, h, h1, h2;
    bool dead;      // h == h' == h'' == 0 exactly (sigmoid saturated): every output is exactly zero
};

// M = bt_inv[f, j] row-major 4x4, Tp = T_pose_21[f, j]
__device__ __forceinline__ HaloBase halo_base(const float* __restrict__ M, const float* __restrict__ Tp,
                                              const float x[3], int j) {
    HaloBase b;
#pragma unroll
    for (int a = 0; a < 3; ++a)
        b.q[a] = ((M[a * 4 + 0] * x[0] + M[a * 4 + 1] * x[1]) + M[a * 4 + 2] * x[2] + M[a * 4 + 3]) - Tp[a];
    b.v = sqrtf(b.q[0] * b.q[0] + b.q[1] * b.q[1] + b.q[2] * b.q[2]);
#pragma unroll
    for (int a = 0; a < 3; ++a) b.r[a] = b.q[a] / b.v;      // no epsilon, as the reference (SURVEY D-6)
    float sg = sigmoidf_(HALO_TAU * (b.v - c_halo_cutoff[j]));
    b.h = 1.0f - sg;
    b.h1 = -HALO_TAU * sg * (1.0f - sg);
    b.h2 = -HALO_TAU * HALO_TAU * sg * (1.0f - sg) * (1.0f - 2.0f * sg);
    b.dead = (b.h == 0.0f) && (b.h1 == 0.0f);
    return b;
}

// out[66] = F(q).  FAST: SFU sin / cos after the Cody-Waite reduction (abs error ~6e-7 on values that already carry ~1e-5 from the
// fp32 cancellation in q = R x + t - T; used by the HN_TC_MIXED16 path, whose tiles round the feature to 22 bits anyway)
template <bool FAST = false>
__device__ __forceinline__ void halo_feature(const HaloBase& b, float* out) {
    if (b.dead) {
#pragma unroll 6
        for (int i = 0; i < HALO_F; ++i) out[i] = 0.0f * (i == 0 ? b.v : 1.0f);   // keeps NaN of v == 0 out of dead joints
        return;
    }
    out[0] = b.v * b.h;
    float f = 1.0f;
    for (int k = 0; k < HALO_LV; ++k) {
        float s, c;
        if (FAST) halo_sincos_fast(b.v * f, &s, &c); else sincosf(b.v * f, &s, &c);
        out[1 + k] = s * b.h;
        out[1 + HALO_LV + k] = c * b.h;
        f *= 2.0f;
    }
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        out[21 + a] = b.r[a] * b.h;
        f = 1.0f;
        for (int k = 0; k < HALO_LR; ++k) {
            float s, c;
            if (FAST) halo_sincos_fast(b.r[a] * f, &s, &c); else sincosf(b.r[a] * f, &s, &c);
            out[24 + a * 14 + k] = s * b.h;
            out[24 + a * 14 + HALO_LR + k] = c * b.h;
            f *= 2.0f;
        }
    }
}

// out[66] = F'(q) w
__device__ __forceinline__ void halo_jvp(const HaloBase& b, const float w[3], float* out) {
    if (b.dead) {
#pragma unroll 6
        for (int i = 0; i < HALO_F; ++i) out[i] = 0.0f;
        return;
    }
    const float dv = b.r[0] * w[0] + b.r[1] * w[1] + b.r[2] * w[2];
    float dr[3];
#pragma unroll
    for (int a = 0; a < 3; ++a) dr[a] = (w[a] - b.r[a] * dv) / b.v;
    const float hd = b.h1 * dv;
    out[0] = dv * b.h + b.v * hd;
    float f = 1.0f;
    for (int k = 0; k < HALO_LV; ++k) {
        float s, c;
        halo_sincos_fast(b.v * f, &s, &c);
        out[1 + k] = (f * c * dv) * b.h + s * hd;
        out[1 + HALO_LV + k] = (-f * s * dv) * b.h + c * hd;
        f *= 2.0f;
    }
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        out[21 + a] = dr[a] * b.h + b.r[a] * hd;
        f = 1.0f;
        for (int k = 0; k < HALO_LR; ++k) {
            float s, c;
            halo_sincos_fast(b.r[a] * f, &s, &c);
            out[24 + a * 14 + k] = (f * c * dr[a]) * b.h + s * hd;
            out[24 + a * 14 + HALO_LR + k] = (-f * s * dr[a]) * b.h + c * hd;
            f *= 2.0f;
        }
    }
}

// g = d<c,F>/dq ; when HVP: hv = d(g.w)/dq.  c points at 66 floats (shared memory), CS floats apart.
template <bool HVP, int CS = 1>
__device__ __forceinline__ void halo_grad_hvp(const HaloBase& b, const float* __restrict__ c, const float w[3],
                                              float g[3], float hv[3]) {
    if (b.dead) {
        g[0] = g[1] = g[2] = 0.0f;
        if (HVP) hv[0] = hv[1] = hv[2] = 0.0f;
        return;
    }
    float A = c[0] * b.v, A1 = c[0], A2 = 0.0f;
    float f = 1.0f;
    for (int k = 0; k < HALO_LV; ++k) {
        float s, co;
        halo_sincos_fast(b.v * f, &s, &co);
        const float cs = c[(1 + k) * CS], cc = c[(1 + HALO_LV + k) * CS];
        A += cs * s + cc * co;
        A1 += f * (cs * co - cc * s);
        A2 -= f * f * (cs * s + cc * co);
        f *= 2.0f;
    }
    float B = 0.0f, b1[3], b2[3];
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        B += c[(21 + a) * CS] * b.r[a];
        b1[a] = c[(21 + a) * CS];
        b2[a] = 0.0f;
        f = 1.0f;
        for (int k = 0; k < HALO_LR; ++k) {
            float s, co;
            halo_sincos_fast(b.r[a] * f, &s, &co);
            const float cs = c[(24 + a * 14 + k) * CS], cc = c[(24 + a * 14 + HALO_LR + k) * CS];
            B += cs * s + cc * co;
            b1[a] += f * (cs * co - cc * s);
            b2[a] -= f * f * (cs * s + cc * co);
            f *= 2.0f;
        }
    }
    const float G = A + B;
    const float rb1 = b.r[0] * b1[0] + b.r[1] * b1[1] + b.r[2] * b1[2];
    float Pb1[3];
#pragma unroll
    for (int a = 0; a < 3; ++a) Pb1[a] = b1[a] - b.r[a] * rb1;
    const float radial = b.h1 * G + b.h * A1;
    const float hov = b.h / b.v;
#pragma unroll
    for (int a = 0; a < 3; ++a) g[a] = radial * b.r[a] + hov * Pb1[a];
    if (!HVP) return;
    const float dv = b.r[0] * w[0] + b.r[1] * w[1] + b.r[2] * w[2];
    float dr[3], db1[3];
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        dr[a] = (w[a] - b.r[a] * dv) / b.v;
        db1[a] = b2[a] * dr[a];
    }
    const float dG = A1 * dv + (b1[0] * dr[0] + b1[1] * dr[1] + b1[2] * dr[2]);
    const float drb1 = dr[0] * b1[0] + dr[1] * b1[1] + dr[2] * b1[2];
    const float rdb1 = b.r[0] * db1[0] + b.r[1] * db1[1] + b.r[2] * db1[2];
    const float dradial = b.h2 * dv * G + b.h1 * dG + b.h1 * dv * A1 + b.h * A2 * dv;
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        const float dPb1 = -dr[a] * rb1 - b.r[a] * drb1 + (db1[a] - b.r[a] * rdb1);
        hv[a] = dradial * b.r[a] + radial * dr[a] + (b.h1 * dv / b.v) * Pb1[a] + hov * dPb1 -
                (b.h * dv / (b.v * b.v)) * Pb1[a];
    }
}

}  // namespace hn
