// Two-field (hand + object) fitting compositor: per-field s-density alpha
// (NeuSRenderer_fitting.get_alpha_sample_color, utils/renderer.py:396-422; utils/renderer_batch.py:150-174)
// and the joint transmittance / colour compositing of both fields (utils/renderer.py:512-524;
// utils/renderer_batch.py:258-270), forward and backward.  One warp per ray, n <= 256 samples.
#include <algorithm>

#include "common.cuh"

namespace hn {

constexpr int FC_WARPS = 4;
constexpr int FC_MAX_CHUNKS = 8;

struct AlphaEval {
    float c, nx, alpha_raw, alpha, true_cos, est_prev, est_next;
};

__device__ __forceinline__ AlphaEval alpha_eval(float sdf, float nx_, float ny_, float nz_, float dx, float dy,
                                                float dz, float dist, float inv_s) {
    AlphaEval e;
    e.true_cos = dx * nx_ + dy * ny_ + dz * nz_;
    float ic = fminf(e.true_cos, 0.0f);                 // cos_anneal_ratio hard-coded to 1.0 upstream
    float half = ic * dist * 0.5f;
    e.est_next = sdf + half;
    e.est_prev = sdf - half;
    e.c = sigmoidf_(e.est_prev * inv_s);
    e.nx = sigmoidf_(e.est_next * inv_s);
    e.alpha_raw = (e.c - e.nx + 1e-5f) / (e.c + 1e-5f);
    e.alpha = fminf(fmaxf(e.alpha_raw, 0.0f), 1.0f);
    return e;
}

__global__ void __launch_bounds__(FC_WARPS * 32) neus_alpha_fwd_kernel(
    const float* __restrict__ sdf, const float* __restrict__ normal, const float* __restrict__ dists,
    const float* __restrict__ rays_d, const float* __restrict__ variance, int64_t n_rays, int n,
    float* __restrict__ alpha, float* __restrict__ eik) {
    const int lane = threadIdx.x & 31;
    int64_t ray = (int64_t)blockIdx.x * FC_WARPS + (threadIdx.x >> 5);
    if (ray >= n_rays) return;
    const float inv_s = fminf(fmaxf(expf(variance[0] * 10.0f), 1e-6f), 1e6f);
    const float dx = rays_d[ray * 3], dy = rays_d[ray * 3 + 1], dz = rays_d[ray * 3 + 2];
    float ek = 0.0f;
    for (int i = lane; i < n; i += 32) {
        int64_t s = ray * n + i;
        float nxv = normal[s * 3], nyv = normal[s * 3 + 1], nzv = normal[s * 3 + 2];
        AlphaEval e = alpha_eval(sdf[s], nxv, nyv, nzv, dx, dy, dz, dists[s], inv_s);
        alpha[s] = e.alpha;
        float nn = sqrtf(nxv * nxv + nyv * nyv + nzv * nzv) - 1.0f;
        ek += nn * nn;
    }
    ek = warp_sum(ek);
    if (lane == 0) eik[ray] = ek;
}

__global__ void __launch_bounds__(FC_WARPS * 32) neus_alpha_bwd_kernel(
    const float* __restrict__ sdf, const float* __restrict__ normal, const float* __restrict__ dists,
    const float* __restrict__ rays_d, const float* __restrict__ variance, int64_t n_rays, int n,
    const float* __restrict__ d_alpha, const float* __restrict__ d_eik, float* __restrict__ d_sdf,
    float* __restrict__ d_normal, float* __restrict__ d_rays_d, float* __restrict__ d_variance) {
    const int lane = threadIdx.x & 31;
    const float inv_s_raw = expf(variance[0] * 10.0f);
    const float inv_s = fminf(fmaxf(inv_s_raw, 1e-6f), 1e6f);
    const bool s_live = inv_s_raw >= 1e-6f && inv_s_raw <= 1e6f;
    float dinv_acc = 0.0f;       // ONE atomic on d_variance per warp (a per-ray atomic serialises at one L2 address)
    for (int64_t ray = (int64_t)blockIdx.x * FC_WARPS + (threadIdx.x >> 5); ray < n_rays;
         ray += (int64_t)gridDim.x * FC_WARPS) {
    const float dx = rays_d[ray * 3], dy = rays_d[ray * 3 + 1], dz = rays_d[ray * 3 + 2];
    const float gek = d_eik ? d_eik[ray] : 0.0f;
    float drx = 0.f, dry = 0.f, drz = 0.f, dinv = 0.f;
    for (int i = lane; i < n; i += 32) {
        int64_t s = ray * n + i;
        float nxv = normal[s * 3], nyv = normal[s * 3 + 1], nzv = normal[s * 3 + 2];
        float dist = dists[s];
        AlphaEval e = alpha_eval(sdf[s], nxv, nyv, nzv, dx, dy, dz, dist, inv_s);
        float dalpha = d_alpha ? d_alpha[s] : 0.0f;
        float dar = (e.alpha_raw >= 0.0f && e.alpha_raw <= 1.0f) ? dalpha : 0.0f;
        float den = e.c + 1e-5f;
        float num = e.c - e.nx + 1e-5f;
        float dc = dar * (1.0f / den - num / (den * den));
        float dnx = -dar / den;
        float dap = dc * e.c * (1.0f - e.c);
        float dan = dnx * e.nx * (1.0f - e.nx);
        dinv += dap * e.est_prev + dan * e.est_next;
        float dprev = dap * inv_s, dnext = dan * inv_s;
        d_sdf[s] = dprev + dnext;
        float dic = (dnext - dprev) * dist * 0.5f;
        float dtc = e.true_cos < 0.0f ? dic : 0.0f;
        float nrm = sqrtf(nxv * nxv + nyv * nyv + nzv * nzv);
        float ke = nrm > 0.0f ? gek * 2.0f * (nrm - 1.0f) / nrm : 0.0f;
        d_normal[s * 3] = dtc * dx + ke * nxv;
        d_normal[s * 3 + 1] = dtc * dy + ke * nyv;
        d_normal[s * 3 + 2] = dtc * dz + ke * nzv;
        drx += dtc * nxv; dry += dtc * nyv; drz += dtc * nzv;
    }
    drx = warp_sum(drx); dry = warp_sum(dry); drz = warp_sum(drz);
    dinv_acc += dinv;
    if (lane == 0 && d_rays_d) { d_rays_d[ray * 3] = drx; d_rays_d[ray * 3 + 1] = dry; d_rays_d[ray * 3 + 2] = drz; }
    }
    dinv_acc = warp_sum(dinv_acc);
    if (lane == 0 && d_variance && s_live && dinv_acc != 0.0f) atomicAdd(d_variance, dinv_acc * 10.0f * inv_s);
}

__device__ __forceinline__ float fc_excl_prod(float f, int lane, float* total) {
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
__device__ __forceinline__ float fc_suffix_sum(float v, int lane, float* total) {
    float inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        float t = __shfl_down_sync(0xffffffffu, inc, o);
        if (lane + o < 32) inc += t;
    }
    *total = __shfl_sync(0xffffffffu, inc, 0);
    return inc;
}

// T_i = prod_{j<i} (1-a_h+1e-7)(1-a_o+1e-7)  (leading ONE);  w_h = a_h T, w_o = a_o T
__global__ void __launch_bounds__(FC_WARPS * 32) fit_composite_fwd_kernel(
    const float* __restrict__ alpha_h, const float* __restrict__ rgb_h, const float* __restrict__ alpha_o,
    const float* __restrict__ rgb_o, int64_t n_rays, int n, float* __restrict__ trans,
    float* __restrict__ color, float* __restrict__ wsum) {
    const int lane = threadIdx.x & 31;
    int64_t ray = (int64_t)blockIdx.x * FC_WARPS + (threadIdx.x >> 5);
    if (ray >= n_rays) return;
    float T = 1.0f;
    float cr = 0.f, cg = 0.f, cb = 0.f, ws = 0.f;
    for (int base = 0; base < n; base += 32) {
        int i = base + lane;
        bool ok = i < n;
        int64_t s = ray * n + (ok ? i : 0);
        float ah = ok ? alpha_h[s] : 0.f, ao = ok ? alpha_o[s] : 0.f;
        float f = ok ? (1.0f - ah + 1e-7f) * (1.0f - ao + 1e-7f) : 1.0f;
        float tot;
        float Ti = T * fc_excl_prod(f, lane, &tot);
        T *= tot;
        if (ok) {
            trans[s] = Ti;
            float wh = ah * Ti, wo = ao * Ti;
            cr += wh * rgb_h[s * 3] + wo * rgb_o[s * 3];
            cg += wh * rgb_h[s * 3 + 1] + wo * rgb_o[s * 3 + 1];
            cb += wh * rgb_h[s * 3 + 2] + wo * rgb_o[s * 3 + 2];
            ws += wh + wo;
        }
    }
    cr = warp_sum(cr); cg = warp_sum(cg); cb = warp_sum(cb); ws = warp_sum(ws);
    if (lane == 0) {
        color[ray * 3] = cr; color[ray * 3 + 1] = cg; color[ray * 3 + 2] = cb;
        wsum[ray] = ws;
    }
}

__global__ void __launch_bounds__(FC_WARPS * 32) fit_composite_bwd_kernel(
    const float* __restrict__ alpha_h, const float* __restrict__ rgb_h, const float* __restrict__ alpha_o,
    const float* __restrict__ rgb_o, const float* __restrict__ trans, int64_t n_rays, int n,
    const float* __restrict__ d_color, const float* __restrict__ d_wsum, float* __restrict__ d_alpha_h,
    float* __restrict__ d_rgb_h, float* __restrict__ d_alpha_o, float* __restrict__ d_rgb_o) {
    const int lane = threadIdx.x & 31;
    int64_t ray = (int64_t)blockIdx.x * FC_WARPS + (threadIdx.x >> 5);
    if (ray >= n_rays) return;
    const float gcr = d_color ? d_color[ray * 3] : 0.f, gcg = d_color ? d_color[ray * 3 + 1] : 0.f,
                gcb = d_color ? d_color[ray * 3 + 2] : 0.f;
    const float gws = d_wsum ? d_wsum[ray] : 0.0f;
    const int chunks = (n + 31) / 32;
    float carry = 0.0f;
#pragma unroll 1
    for (int t = chunks - 1; t >= 0; --t) {
        int i = t * 32 + lane;
        bool ok = i < n;
        int64_t s = ray * n + (ok ? i : 0);
        float ah = 0.f, ao = 0.f, Ti = 0.f, gh = 0.f, go = 0.f;
        if (ok) {
            ah = alpha_h[s]; ao = alpha_o[s]; Ti = trans[s];
            gh = gcr * rgb_h[s * 3] + gcg * rgb_h[s * 3 + 1] + gcb * rgb_h[s * 3 + 2] + gws;
            go = gcr * rgb_o[s * 3] + gcg * rgb_o[s * 3 + 1] + gcb * rgb_o[s * 3 + 2] + gws;
        }
        float own = gh * ah * Ti + go * ao * Ti;
        float tot;
        float S = fc_suffix_sum(own, lane, &tot) + carry;     // sum over k >= i
        carry += tot;
        if (ok) {
            float later = S - own;                              // sum over k > i
            d_alpha_h[s] = gh * Ti - later / (1.0f - ah + 1e-7f);
            d_alpha_o[s] = go * Ti - later / (1.0f - ao + 1e-7f);
            float wh = ah * Ti, wo = ao * Ti;
            d_rgb_h[s * 3] = wh * gcr; d_rgb_h[s * 3 + 1] = wh * gcg; d_rgb_h[s * 3 + 2] = wh * gcb;
            d_rgb_o[s * 3] = wo * gcr; d_rgb_o[s * 3 + 1] = wo * gcg; d_rgb_o[s * 3 + 2] = wo * gcb;
        }
    }
}

}  // namespace hn

using namespace hn;

extern "C" {

int hn_neus_alpha_fwd(const float* sdf, const float* normal, const float* dists, const float* rays_d,
                      const float* variance, int64_t n_rays, int n, float* alpha, float* eik,
                      hn_stream_t stream) {
    HN_REQUIRE(n_rays >= 0 && n > 0, "hn_neus_alpha_fwd: bad sizes");
    if (n_rays == 0) return HN_OK;
    HN_REQUIRE(sdf && normal && dists && rays_d && variance && alpha && eik, "hn_neus_alpha_fwd: null pointer");
    neus_alpha_fwd_kernel<<<(unsigned)ceil_div(n_rays, FC_WARPS), FC_WARPS * 32, 0, (cudaStream_t)stream>>>(
        sdf, normal, dists, rays_d, variance, n_rays, n, alpha, eik);
    count_launch();
    HN_CHECK_LAUNCH();
    return HN_OK;
}

int hn_neus_alpha_bwd(const float* sdf, const float* normal, const float* dists, const float* rays_d,
                      const float* variance, int64_t n_rays, int n, const float* d_alpha, const float* d_eik,
                      float* d_sdf, float* d_normal, float* d_rays_d, float* d_variance, hn_stream_t stream) {
    HN_REQUIRE(n_rays >= 0 && n > 0, "hn_neus_alpha_bwd: bad sizes");
    if (n_rays == 0) return HN_OK;
    HN_REQUIRE(sdf && normal && dists && rays_d && variance && d_sdf && d_normal, "hn_neus_alpha_bwd: null pointer");
    const unsigned grid = (unsigned)std::min<int64_t>(ceil_div(n_rays, FC_WARPS), (int64_t)sm_count() * 16);
    neus_alpha_bwd_kernel<<<grid, FC_WARPS * 32, 0, (cudaStream_t)stream>>>(
        sdf, normal, dists, rays_d, variance, n_rays, n, d_alpha, d_eik, d_sdf, d_normal, d_rays_d, d_variance);
    count_launch();
    HN_CHECK_LAUNCH();
    return HN_OK;
}

int hn_fit_composite_fwd(const float* alpha_h, const float* rgb_h, const float* alpha_o, const float* rgb_o,
                         int64_t n_rays, int n, float* trans, float* color, float* weight_sum,
                         hn_stream_t stream) {
    HN_REQUIRE(n_rays >= 0 && n > 0, "hn_fit_composite_fwd: bad sizes");
    if (n_rays == 0) return HN_OK;
    HN_REQUIRE(alpha_h && rgb_h && alpha_o && rgb_o && trans && color && weight_sum,
               "hn_fit_composite_fwd: null pointer");
    fit_composite_fwd_kernel<<<(unsigned)ceil_div(n_rays, FC_WARPS), FC_WARPS * 32, 0, (cudaStream_t)stream>>>(
        alpha_h, rgb_h, alpha_o, rgb_o, n_rays, n, trans, color, weight_sum);
    count_launch();
    HN_CHECK_LAUNCH();
    return HN_OK;
}

int hn_fit_composite_bwd(const float* alpha_h, const float* rgb_h, const float* alpha_o, const float* rgb_o,
                         const float* trans, int64_t n_rays, int n, const float* d_color,
                         const float* d_weight_sum, float* d_alpha_h, float* d_rgb_h, float* d_alpha_o,
                         float* d_rgb_o, hn_stream_t stream) {
    HN_REQUIRE(n_rays >= 0 && n > 0, "hn_fit_composite_bwd: bad sizes");
    if (n_rays == 0) return HN_OK;
    HN_REQUIRE(alpha_h && rgb_h && alpha_o && rgb_o && trans && d_alpha_h && d_rgb_h && d_alpha_o && d_rgb_o,
               "hn_fit_composite_bwd: null pointer");
    fit_composite_bwd_kernel<<<(unsigned)ceil_div(n_rays, FC_WARPS), FC_WARPS * 32, 0, (cudaStream_t)stream>>>(
        alpha_h, rgb_h, alpha_o, rgb_o, trans, n_rays, n, d_color, d_weight_sum, d_alpha_h, d_rgb_h, d_alpha_o,
        d_rgb_o);
    count_launch();
    HN_CHECK_LAUNCH();
    return HN_OK;
}

}  // extern "C"
