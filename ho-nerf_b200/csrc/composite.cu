// NeuS s-density alpha, transmittance scan and colour compositing, forward and backward:
// replaces utils/renderer.py:144-169 (render_core) and the autograd graph behind it.
// One warp per ray; every per-sample buffer is read/written once, coalesced along the ray.
// Algorithmic traffic per sample: fwd 40 B, fwd+bwd 104 B.
#include <algorithm>
#include <initializer_list>

#include "common.cuh"

namespace hn {

constexpr int CMP_WARPS = 4;
constexpr int CMP_MAX_CHUNKS = 8;   // n <= 256 samples per ray

struct SampleEval {
    float c, nx, alpha_raw, alpha, true_cos, est_prev, est_next;
};

__device__ __forceinline__ SampleEval eval_sample(float sdf, float nx_, float ny_, float nz_, float dx,
                                                  float dy, float dz, float dist, float inv_s) {
    SampleEval e;
    e.true_cos = dx * nx_ + dy * ny_ + dz * nz_;
    float ic = fminf(e.true_cos, 0.0f);                 // -relu(-true_cos), anneal ratio 1.0
    float half = ic * dist * 0.5f;
    e.est_next = sdf + half;
    e.est_prev = sdf - half;
    e.c = sigmoidf_(e.est_prev * inv_s);
    e.nx = sigmoidf_(e.est_next * inv_s);
    e.alpha_raw = (e.c - e.nx + 1e-5f) / (e.c + 1e-5f);
    e.alpha = fminf(fmaxf(e.alpha_raw, 0.0f), 1.0f);
    return e;
}

// The same quantities for the vectorised BACKWARD with SFU exp / reciprocal (rel. error ~1e-6, against 1e-2 on gradients): the
// IEEE divisions and expf of the exact version are a third of that kernel's instructions (FCHK + slow-path branches), and it
// is issue-bound at 54 % with 16 resident warps.  The forward keeps the exact functions (its weights are parity-tested).
__device__ __forceinline__ float rcp_fast(float x) { return __fdividef(1.0f, x); }
__device__ __forceinline__ float sigmoid_fast(float x) { return rcp_fast(1.0f + __expf(-x)); }
__device__ __forceinline__ SampleEval eval_sample_fast(float sdf, float nx_, float ny_, float nz_, float dx, float dy, float dz,
                                                       float dist, float inv_s) {
    SampleEval e;
    e.true_cos = dx * nx_ + dy * ny_ + dz * nz_;
    const float half = fminf(e.true_cos, 0.0f) * dist * 0.5f;
    e.est_next = sdf + half;
    e.est_prev = sdf - half;
    e.c = sigmoid_fast(e.est_prev * inv_s);
    e.nx = sigmoid_fast(e.est_next * inv_s);
    e.alpha_raw = (e.c - e.nx + 1e-5f) * rcp_fast(e.c + 1e-5f);
    e.alpha = fminf(fmaxf(e.alpha_raw, 0.0f), 1.0f);
    return e;
}

__device__ __forceinline__ float inv_s_of(const float* variance) {
    return fminf(fmaxf(expf(variance[0] * 10.0f), 1e-6f), 1e6f);
}

// exclusive product scan over the warp: returns prod of f over lanes < lane; total in *total
__device__ __forceinline__ float warp_excl_prod(float f, int lane, float* total) {
    float inc = f;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        float t = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc *= t;
    }
    *total = __shfl_sync(0xffffffffu, inc, 31);
    float ex = __shfl_up_sync(0xffffffffu, inc, 1);
    return lane == 0 ? 1.0f : ex;
}
// inclusive suffix sum over the warp: sum of v over lanes >= lane
__device__ __forceinline__ float warp_suffix_sum(float v, int lane, float* total) {
    float inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        float t = __shfl_down_sync(0xffffffffu, inc, o);
        if (lane + o < 32) inc += t;
    }
    *total = __shfl_sync(0xffffffffu, inc, 0);
    return inc;
}

__global__ void __launch_bounds__(CMP_WARPS * 32) neus_composite_fwd_kernel(
    const float* __restrict__ sdf, const float* __restrict__ normal, const float* __restrict__ rgb,
    const float* __restrict__ dists, const float* __restrict__ rays_d, const float* __restrict__ variance,
    int64_t n_rays, int n, int seed_c0, float* __restrict__ weights, float* __restrict__ cdf,
    float* __restrict__ alpha_out, float* __restrict__ color, float* __restrict__ wsum_out,
    float* __restrict__ wmax_out, float* __restrict__ eik_out) {
    const int lane = threadIdx.x & 31;
    int64_t ray = (int64_t)blockIdx.x * CMP_WARPS + (threadIdx.x >> 5);
    if (ray >= n_rays) return;
    const float inv_s = inv_s_of(variance);
    const float dx = rays_d[ray * 3], dy = rays_d[ray * 3 + 1], dz = rays_d[ray * 3 + 2];
    float T = 1.0f;          // running transmittance at the start of the current chunk
    float cr = 0.f, cg = 0.f, cb = 0.f, ws = 0.f, wm = 0.f, ek = 0.f;
    for (int base = 0; base < n; base += 32) {
        int i = base + lane;
        bool ok = i < n;
        int64_t s = ray * n + (ok ? i : 0);
        float nxv = ok ? normal[s * 3] : 0.f, nyv = ok ? normal[s * 3 + 1] : 0.f, nzv = ok ? normal[s * 3 + 2] : 0.f;
        SampleEval e = eval_sample(ok ? sdf[s] : 0.f, nxv, nyv, nzv, dx, dy, dz, ok ? dists[s] : 0.f, inv_s);
        if (base == 0 && seed_c0) T = __shfl_sync(0xffffffffu, e.c, 0);
        float f = ok ? (1.0f - e.alpha + 1e-7f) : 1.0f;
        float tot;
        float Ti = T * warp_excl_prod(f, lane, &tot);
        T *= tot;
        if (ok) {
            float w = e.alpha * Ti;
            weights[s] = w;
            cdf[s] = e.c;
            if (alpha_out) alpha_out[s] = e.alpha;
            cr += w * rgb[s * 3]; cg += w * rgb[s * 3 + 1]; cb += w * rgb[s * 3 + 2];
            ws += w;
            wm = fmaxf(wm, w);
            float nn = sqrtf(nxv * nxv + nyv * nyv + nzv * nzv) - 1.0f;
            ek += nn * nn;
        }
    }
    cr = warp_sum(cr); cg = warp_sum(cg); cb = warp_sum(cb);
    ws = warp_sum(ws); ek = warp_sum(ek); wm = warp_max(wm);
    if (lane == 0) {
        color[ray * 3] = cr; color[ray * 3 + 1] = cg; color[ray * 3 + 2] = cb;
        wsum_out[ray] = ws;
        wmax_out[ray] = wm;
        eik_out[ray] = ek;
    }
}

__global__ void __launch_bounds__(CMP_WARPS * 32) neus_composite_bwd_kernel(
    const float* __restrict__ sdf, const float* __restrict__ normal, const float* __restrict__ rgb,
    const float* __restrict__ dists, const float* __restrict__ rays_d, const float* __restrict__ variance,
    const float* __restrict__ weights, int64_t n_rays, int n, int seed_c0,
    const float* __restrict__ d_color, const float* __restrict__ d_wsum, const float* __restrict__ d_weights,
    const float* __restrict__ d_eik, float* __restrict__ d_sdf, float* __restrict__ d_normal,
    float* __restrict__ d_rgb, float* __restrict__ d_rays_d, float* __restrict__ d_variance) {
    const int lane = threadIdx.x & 31;
    int64_t ray = (int64_t)blockIdx.x * CMP_WARPS + (threadIdx.x >> 5);
    if (ray >= n_rays) return;
    const float var = variance[0];
    const float inv_s_raw = expf(var * 10.0f);
    const float inv_s = fminf(fmaxf(inv_s_raw, 1e-6f), 1e6f);
    const bool s_live = inv_s_raw >= 1e-6f && inv_s_raw <= 1e6f;
    const float dx = rays_d[ray * 3], dy = rays_d[ray * 3 + 1], dz = rays_d[ray * 3 + 2];
    const float gcr = d_color[ray * 3], gcg = d_color[ray * 3 + 1], gcb = d_color[ray * 3 + 2];
    const float gws = d_wsum ? d_wsum[ray] : 0.0f;
    const float gek = d_eik ? d_eik[ray] : 0.0f;
    const int chunks = (n + 31) / 32;

    // pass 1 (forward order): recompute alpha, transmittance; keep per-lane values in registers
    SampleEval ev[CMP_MAX_CHUNKS];
    float Ti[CMP_MAX_CHUNKS], gw[CMP_MAX_CHUNKS];
    float T = 1.0f, c0 = 1.0f;
#pragma unroll
    for (int t = 0; t < CMP_MAX_CHUNKS; ++t) {
        if (t < chunks) {
            int i = t * 32 + lane;
            bool ok = i < n;
            int64_t s = ray * n + (ok ? i : 0);
            float nxv = ok ? normal[s * 3] : 0.f, nyv = ok ? normal[s * 3 + 1] : 0.f, nzv = ok ? normal[s * 3 + 2] : 0.f;
            ev[t] = eval_sample(ok ? sdf[s] : 0.f, nxv, nyv, nzv, dx, dy, dz, ok ? dists[s] : 0.f, inv_s);
            if (t == 0) {
                c0 = __shfl_sync(0xffffffffu, ev[0].c, 0);
                if (seed_c0) T = c0;
            }
            float f = ok ? (1.0f - ev[t].alpha + 1e-7f) : 1.0f;
            float tot;
            Ti[t] = T * warp_excl_prod(f, lane, &tot);
            T *= tot;
            float g = 0.0f;
            if (ok) {
                g = gcr * rgb[s * 3] + gcg * rgb[s * 3 + 1] + gcb * rgb[s * 3 + 2] + gws;
                if (d_weights) g += d_weights[s];
                float w = weights[s];
                d_rgb[s * 3] = w * gcr; d_rgb[s * 3 + 1] = w * gcg; d_rgb[s * 3 + 2] = w * gcb;
                gw[t] = g * w;
            } else {
                gw[t] = 0.0f;
            }
            // g itself is re-derived in pass 2 from the same inputs (cheaper than another register array)
        }
    }
    // pass 2 (reverse order): suffix sums S_i = sum_{k>=i} g_k w_k, then all input cotangents
    float carry = 0.0f;       // sum of g_k w_k over later chunks
    float drx = 0.f, dry = 0.f, drz = 0.f, dinv = 0.f;
#pragma unroll
    for (int t = CMP_MAX_CHUNKS - 1; t >= 0; --t) {
        if (t < chunks) {
            int i = t * 32 + lane;
            bool ok = i < n;
            int64_t s = ray * n + (ok ? i : 0);
            float tot;
            float S = warp_suffix_sum(gw[t], lane, &tot) + carry;
            carry += tot;
            if (ok) {
                const SampleEval& e = ev[t];
                float w = weights[s];
                float g = gcr * rgb[s * 3] + gcg * rgb[s * 3 + 1] + gcb * rgb[s * 3 + 2] + gws;
                if (d_weights) g += d_weights[s];
                float f = 1.0f - e.alpha + 1e-7f;
                float dalpha = g * Ti[t] - (S - g * w) / f;
                float dar = (e.alpha_raw >= 0.0f && e.alpha_raw <= 1.0f) ? dalpha : 0.0f;
                float den = e.c + 1e-5f;
                float num = e.c - e.nx + 1e-5f;
                float dc = dar * (1.0f / den - num / (den * den));
                float dnx = -dar / den;
                if (seed_c0 && i == 0) dc += S / c0;      // S at i == 0 is S_0
                float dap = dc * e.c * (1.0f - e.c);
                float dan = dnx * e.nx * (1.0f - e.nx);
                dinv += dap * e.est_prev + dan * e.est_next;
                float dprev = dap * inv_s, dnext = dan * inv_s;
                d_sdf[s] = dprev + dnext;
                float dic = (dnext - dprev) * dists[s] * 0.5f;
                float dtc = e.true_cos < 0.0f ? dic : 0.0f;
                float nxv = normal[s * 3], nyv = normal[s * 3 + 1], nzv = normal[s * 3 + 2];
                float nrm = sqrtf(nxv * nxv + nyv * nyv + nzv * nzv);
                float ke = nrm > 0.0f ? gek * 2.0f * (nrm - 1.0f) / nrm : 0.0f;
                d_normal[s * 3] = dtc * dx + ke * nxv;
                d_normal[s * 3 + 1] = dtc * dy + ke * nyv;
                d_normal[s * 3 + 2] = dtc * dz + ke * nzv;
                drx += dtc * nxv; dry += dtc * nyv; drz += dtc * nzv;
            }
        }
    }
    drx = warp_sum(drx); dry = warp_sum(dry); drz = warp_sum(drz); dinv = warp_sum(dinv);
    if (lane == 0) {
        if (d_rays_d) { d_rays_d[ray * 3] = drx; d_rays_d[ray * 3 + 1] = dry; d_rays_d[ray * 3 + 2] = drz; }
        if (d_variance && s_live) atomicAdd(d_variance, dinv * 10.0f * inv_s);
    }
}

// ------------------------------------------------------------------------------------------------
// Vectorised variants for n = 32 * S samples per ray (S = 4, 6, 8: the 128-sample render_core and the
// 192-sample fitting renderers): a lane owns S CONSECUTIVE samples, so every per-sample buffer is moved with
// 8/16-byte vector accesses (a warp instruction covers one contiguous segment of the ray), the transmittance
// is a serial product inside the lane plus one warp scan of the lane totals, and nothing is re-read.
// ------------------------------------------------------------------------------------------------
template <int K>
__device__ __forceinline__ void load_vec(const float* __restrict__ p, float* out) {
    if constexpr (K % 4 == 0) {
#pragma unroll
        for (int i = 0; i < K; i += 4) {
            const float4 v = *reinterpret_cast<const float4*>(p + i);
            out[i] = v.x; out[i + 1] = v.y; out[i + 2] = v.z; out[i + 3] = v.w;
        }
    } else {
#pragma unroll
        for (int i = 0; i < K; i += 2) {
            const float2 v = *reinterpret_cast<const float2*>(p + i);
            out[i] = v.x; out[i + 1] = v.y;
        }
    }
}
template <int K>
__device__ __forceinline__ void store_vec(float* __restrict__ p, const float* in) {
    if constexpr (K % 4 == 0) {
#pragma unroll
        for (int i = 0; i < K; i += 4) *reinterpret_cast<float4*>(p + i) = make_float4(in[i], in[i + 1], in[i + 2], in[i + 3]);
    } else {
#pragma unroll
        for (int i = 0; i < K; i += 2) *reinterpret_cast<float2*>(p + i) = make_float2(in[i], in[i + 1]);
    }
}

// A lane's 3*S consecutive floats (normal / colour cotangents) are stored through a per-warp shared-memory transpose,
// so that every warp store instruction writes one contiguous 512-byte segment (a direct store would put 16-byte
// pieces at a 48-byte stride: half-filled sectors at L2).  K = 3*S floats per lane, K % 4 == 0.
template <int K>
__device__ __forceinline__ void store_vec_coalesced(float* __restrict__ warp_base, const float* in, float4* sbuf, int lane) {
    constexpr int NV = K / 4;
#pragma unroll
    for (int j = 0; j < NV; ++j) sbuf[j * 32 + lane] = make_float4(in[4 * j], in[4 * j + 1], in[4 * j + 2], in[4 * j + 3]);
    __syncwarp();
#pragma unroll
    for (int k = 0; k < NV; ++k) {
        const int g = k * 32 + lane;                  // float4 index inside the warp's contiguous block
        reinterpret_cast<float4*>(warp_base)[g] = sbuf[(g % NV) * 32 + g / NV];
    }
    __syncwarp();
}

template <int S>
__global__ void __launch_bounds__(CMP_WARPS * 32) neus_composite_fwd_vec_kernel(
    const float* __restrict__ sdf, const float* __restrict__ normal, const float* __restrict__ rgb,
    const float* __restrict__ dists, const float* __restrict__ rays_d, const float* __restrict__ variance,
    int64_t n_rays, int seed_c0, float* __restrict__ weights, float* __restrict__ cdf,
    float* __restrict__ alpha_out, float* __restrict__ color, float* __restrict__ wsum_out,
    float* __restrict__ wmax_out, float* __restrict__ eik_out) {
    const int lane = threadIdx.x & 31;
    const int64_t ray = (int64_t)blockIdx.x * CMP_WARPS + (threadIdx.x >> 5);
    if (ray >= n_rays) return;
    const float inv_s = inv_s_of(variance);
    const float dx = rays_d[ray * 3], dy = rays_d[ray * 3 + 1], dz = rays_d[ray * 3 + 2];
    const int64_t s0 = ray * (32 * S) + lane * S;
    float vs[S], vd[S], vn[3 * S], vc[3 * S];
    load_vec<S>(sdf + s0, vs);
    load_vec<S>(dists + s0, vd);
    load_vec<3 * S>(normal + s0 * 3, vn);
    load_vec<3 * S>(rgb + s0 * 3, vc);
    float al[S], cc[S], tl[S];
    float prod = 1.0f, ek = 0.0f;
#pragma unroll
    for (int i = 0; i < S; ++i) {
        const SampleEval e = eval_sample(vs[i], vn[3 * i], vn[3 * i + 1], vn[3 * i + 2], dx, dy, dz, vd[i], inv_s);
        al[i] = e.alpha; cc[i] = e.c;
        tl[i] = prod;                        // product over the lane's earlier samples
        prod *= 1.0f - e.alpha + 1e-7f;
        const float nn = sqrtf(vn[3 * i] * vn[3 * i] + vn[3 * i + 1] * vn[3 * i + 1] + vn[3 * i + 2] * vn[3 * i + 2]) - 1.0f;
        ek += nn * nn;
    }
    float tot;
    float T = warp_excl_prod(prod, lane, &tot);
    if (seed_c0) T *= __shfl_sync(0xffffffffu, cc[0], 0);
    float w[S], cr = 0.f, cg = 0.f, cb = 0.f, ws = 0.f, wm = 0.f;
#pragma unroll
    for (int i = 0; i < S; ++i) {
        w[i] = al[i] * (T * tl[i]);
        cr += w[i] * vc[3 * i]; cg += w[i] * vc[3 * i + 1]; cb += w[i] * vc[3 * i + 2];
        ws += w[i];
        wm = fmaxf(wm, w[i]);
    }
    store_vec<S>(weights + s0, w);
    store_vec<S>(cdf + s0, cc);
    if (alpha_out) store_vec<S>(alpha_out + s0, al);
    cr = warp_sum(cr); cg = warp_sum(cg); cb = warp_sum(cb);
    ws = warp_sum(ws); ek = warp_sum(ek); wm = warp_max(wm);
    if (lane == 0) {
        color[ray * 3] = cr; color[ray * 3 + 1] = cg; color[ray * 3 + 2] = cb;
        wsum_out[ray] = ws;
        wmax_out[ray] = wm;
        eik_out[ray] = ek;
    }
}

template <int S>
__global__ void __launch_bounds__(CMP_WARPS * 32) neus_composite_bwd_vec_kernel(
    const float* __restrict__ sdf, const float* __restrict__ normal, const float* __restrict__ rgb,
    const float* __restrict__ dists, const float* __restrict__ rays_d, const float* __restrict__ variance,
    int64_t n_rays, int seed_c0, const float* __restrict__ d_color, const float* __restrict__ d_wsum,
    const float* __restrict__ d_weights, const float* __restrict__ d_eik, float* __restrict__ d_sdf,
    float* __restrict__ d_normal, float* __restrict__ d_rgb, float* __restrict__ d_rays_d,
    float* __restrict__ d_variance) {
    __shared__ float4 s_tr[(3 * S) % 4 == 0 ? CMP_WARPS * (3 * S / 4) * 32 : 1];
    const int lane = threadIdx.x & 31;
    const float inv_s_raw = expf(variance[0] * 10.0f);
    const float inv_s = fminf(fmaxf(inv_s_raw, 1e-6f), 1e6f);
    const bool s_live = inv_s_raw >= 1e-6f && inv_s_raw <= 1e6f;
    float dinv_acc = 0.0f;       // d/d inv_s summed over this warp's rays: ONE atomic per warp at the end
    for (int64_t ray = (int64_t)blockIdx.x * CMP_WARPS + (threadIdx.x >> 5); ray < n_rays;
         ray += (int64_t)gridDim.x * CMP_WARPS) {
    const float dx = rays_d[ray * 3], dy = rays_d[ray * 3 + 1], dz = rays_d[ray * 3 + 2];
    const float gcr = d_color[ray * 3], gcg = d_color[ray * 3 + 1], gcb = d_color[ray * 3 + 2];
    const float gws = d_wsum ? d_wsum[ray] : 0.0f;
    const float gek = d_eik ? d_eik[ray] : 0.0f;
    const int64_t s0 = ray * (32 * S) + lane * S;
    float vs[S], vd[S], vn[3 * S], vc[3 * S], gwt[S];
    load_vec<S>(sdf + s0, vs);
    load_vec<S>(dists + s0, vd);
    load_vec<3 * S>(normal + s0 * 3, vn);
    load_vec<3 * S>(rgb + s0 * 3, vc);
    if (d_weights) {
        load_vec<S>(d_weights + s0, gwt);
    } else {
#pragma unroll
        for (int i = 0; i < S; ++i) gwt[i] = 0.0f;
    }
    SampleEval ev[S];
    float tl[S];
    float prod = 1.0f;
#pragma unroll
    for (int i = 0; i < S; ++i) {
        ev[i] = eval_sample_fast(vs[i], vn[3 * i], vn[3 * i + 1], vn[3 * i + 2], dx, dy, dz, vd[i], inv_s);
        tl[i] = prod;
        prod *= 1.0f - ev[i].alpha + 1e-7f;
    }
    float tot;
    float T = warp_excl_prod(prod, lane, &tot);
    const float c0 = __shfl_sync(0xffffffffu, ev[0].c, 0);
    if (seed_c0) T *= c0;
    // g_i = dL/dw_i ; suffix sums of g_i w_i: inside the lane, then across lanes
    float g[S], w[S], suf[S], run = 0.0f;
#pragma unroll
    for (int i = S - 1; i >= 0; --i) {
        w[i] = ev[i].alpha * (T * tl[i]);
        g[i] = gcr * vc[3 * i] + gcg * vc[3 * i + 1] + gcb * vc[3 * i + 2] + gws + gwt[i];
        run += g[i] * w[i];
        suf[i] = run;                        // sum over this lane's samples >= i
    }
    float total;
    const float later = warp_suffix_sum(run, lane, &total) - run;     // lanes > this one
    float o_sdf[S], o_nrm[3 * S], o_rgb[3 * S];
    float drx = 0.f, dry = 0.f, drz = 0.f, dinv = 0.f;
#pragma unroll
    for (int i = 0; i < S; ++i) {
        const SampleEval& e = ev[i];
        const float Si = suf[i] + later;                 // sum over all samples k >= this one
        const float f = 1.0f - e.alpha + 1e-7f;
        const float dalpha = g[i] * (T * tl[i]) - (Si - g[i] * w[i]) * rcp_fast(f);
        const float dar = (e.alpha_raw >= 0.0f && e.alpha_raw <= 1.0f) ? dalpha : 0.0f;
        const float rden = rcp_fast(e.c + 1e-5f);
        const float num = e.c - e.nx + 1e-5f;
        float dc = dar * (rden - num * rden * rden);
        const float dnx = -dar * rden;
        if (seed_c0 && lane == 0 && i == 0) dc += Si / c0;
        const float dap = dc * e.c * (1.0f - e.c);
        const float dan = dnx * e.nx * (1.0f - e.nx);
        dinv += dap * e.est_prev + dan * e.est_next;
        const float dprev = dap * inv_s, dnext = dan * inv_s;
        o_sdf[i] = dprev + dnext;
        const float dic = (dnext - dprev) * vd[i] * 0.5f;
        const float dtc = e.true_cos < 0.0f ? dic : 0.0f;
        const float nxv = vn[3 * i], nyv = vn[3 * i + 1], nzv = vn[3 * i + 2];
        const float n2 = nxv * nxv + nyv * nyv + nzv * nzv;
        const float ke = n2 > 0.0f ? gek * 2.0f * (1.0f - rsqrtf(n2)) : 0.0f;       // 2 (|n| - 1) / |n|
        o_nrm[3 * i] = dtc * dx + ke * nxv;
        o_nrm[3 * i + 1] = dtc * dy + ke * nyv;
        o_nrm[3 * i + 2] = dtc * dz + ke * nzv;
        drx += dtc * nxv; dry += dtc * nyv; drz += dtc * nzv;
        o_rgb[3 * i] = w[i] * gcr; o_rgb[3 * i + 1] = w[i] * gcg; o_rgb[3 * i + 2] = w[i] * gcb;
    }
    store_vec<S>(d_sdf + s0, o_sdf);
    if constexpr ((3 * S) % 4 == 0) {
        float4* sbuf = s_tr + (threadIdx.x >> 5) * (3 * S / 4) * 32;
        store_vec_coalesced<3 * S>(d_normal + ray * (32 * S) * 3, o_nrm, sbuf, lane);
        store_vec_coalesced<3 * S>(d_rgb + ray * (32 * S) * 3, o_rgb, sbuf, lane);
    } else {
        store_vec<3 * S>(d_normal + s0 * 3, o_nrm);
        store_vec<3 * S>(d_rgb + s0 * 3, o_rgb);
    }
    drx = warp_sum(drx); dry = warp_sum(dry); drz = warp_sum(drz);
    dinv_acc += dinv;
    if (lane == 0 && d_rays_d) { d_rays_d[ray * 3] = drx; d_rays_d[ray * 3 + 1] = dry; d_rays_d[ray * 3 + 2] = drz; }
    }
    dinv_acc = warp_sum(dinv_acc);
    if (lane == 0 && d_variance && s_live && dinv_acc != 0.0f) atomicAdd(d_variance, dinv_acc * 10.0f * inv_s);
}

static inline bool vec_ok(int n, std::initializer_list<const void*> ptrs) {
    if (n != 128 && n != 192 && n != 256) return false;
    for (const void* p : ptrs)
        if (p && (reinterpret_cast<uintptr_t>(p) & 15) != 0) return false;
    return true;
}

}  // namespace hn

using namespace hn;

extern "C" {

int hn_neus_composite_fwd(const float* sdf, const float* normal, const float* rgb, const float* dists,
                          const float* rays_d, const float* variance, int64_t n_rays, int n,
                          int seed_with_c0, float* weights, float* cdf, float* alpha, float* color,
                          float* weight_sum, float* weight_max, float* eik, hn_stream_t stream) {
    HN_REQUIRE(n_rays >= 0 && n > 0, "hn_neus_composite_fwd: bad sizes");
    if (n_rays == 0) return HN_OK;
    HN_REQUIRE(sdf && normal && rgb && dists && rays_d && variance && weights && cdf && color && weight_sum &&
                   weight_max && eik, "hn_neus_composite_fwd: null pointer");
    const unsigned grid = (unsigned)ceil_div(n_rays, CMP_WARPS);
    cudaStream_t st = (cudaStream_t)stream;
#define HN_FWD_VEC(S)                                                                                                  \
    neus_composite_fwd_vec_kernel<S><<<grid, CMP_WARPS * 32, 0, st>>>(sdf, normal, rgb, dists, rays_d, variance, n_rays, \
                                                                      seed_with_c0, weights, cdf, alpha, color,        \
                                                                      weight_sum, weight_max, eik)
    if (vec_ok(n, {sdf, normal, rgb, dists, weights, cdf, alpha})) {
        if (n == 128) HN_FWD_VEC(4); else if (n == 192) HN_FWD_VEC(6); else HN_FWD_VEC(8);
    } else {
        neus_composite_fwd_kernel<<<grid, CMP_WARPS * 32, 0, st>>>(sdf, normal, rgb, dists, rays_d, variance, n_rays, n,
                                                                   seed_with_c0, weights, cdf, alpha, color, weight_sum,
                                                                   weight_max, eik);
    }
#undef HN_FWD_VEC
    count_launch();
    HN_CHECK_LAUNCH();
    return HN_OK;
}

int hn_neus_composite_bwd(const float* sdf, const float* normal, const float* rgb, const float* dists,
                          const float* rays_d, const float* variance, const float* weights, int64_t n_rays,
                          int n, int seed_with_c0, const float* d_color, const float* d_weight_sum,
                          const float* d_weights, const float* d_eik, float* d_sdf, float* d_normal,
                          float* d_rgb, float* d_rays_d, float* d_variance, hn_stream_t stream) {
    HN_REQUIRE(n_rays >= 0 && n > 0, "hn_neus_composite_bwd: bad sizes");
    if (n > 32 * CMP_MAX_CHUNKS) {
        set_error("hn_neus_composite_bwd: n=%d samples per ray exceeds %d", n, 32 * CMP_MAX_CHUNKS);
        return HN_ERR_UNSUPPORTED;
    }
    if (n_rays == 0) return HN_OK;
    HN_REQUIRE(sdf && normal && rgb && dists && rays_d && variance && weights && d_color && d_sdf && d_normal &&
                   d_rgb, "hn_neus_composite_bwd: null pointer");
    const unsigned grid = (unsigned)ceil_div(n_rays, CMP_WARPS);
    cudaStream_t st = (cudaStream_t)stream;
    const unsigned grid_p = (unsigned)std::min<int64_t>(grid, (int64_t)sm_count() * 16);      // persistent warps
#define HN_BWD_VEC(S)                                                                                                  \
    neus_composite_bwd_vec_kernel<S><<<grid_p, CMP_WARPS * 32, 0, st>>>(sdf, normal, rgb, dists, rays_d, variance, n_rays, \
                                                                      seed_with_c0, d_color, d_weight_sum, d_weights,  \
                                                                      d_eik, d_sdf, d_normal, d_rgb, d_rays_d, d_variance)
    if (vec_ok(n, {sdf, normal, rgb, dists, d_weights, d_sdf, d_normal, d_rgb})) {
        if (n == 128) HN_BWD_VEC(4); else if (n == 192) HN_BWD_VEC(6); else HN_BWD_VEC(8);
    } else {
        neus_composite_bwd_kernel<<<grid, CMP_WARPS * 32, 0, st>>>(sdf, normal, rgb, dists, rays_d, variance, weights, n_rays,
                                                                   n, seed_with_c0, d_color, d_weight_sum, d_weights, d_eik,
                                                                   d_sdf, d_normal, d_rgb, d_rays_d, d_variance);
    }
#undef HN_BWD_VEC
    count_launch();
    HN_CHECK_LAUNCH();
    return HN_OK;
}

}  // extern "C"
