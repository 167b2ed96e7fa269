// render_core_outside: the background branch of a NeuS renderer (named by the north star; the HO-NeRF reference only STORES
// n_outside -- utils/renderer.py:47,56, every config sets 0 -- and has no such method, so this follows the semantics of the
// NeuS renderer HO-NeRF's was derived from: PARITY UNPINNED, checked against oracle/honerf_oracle.render_core_outside only).
//
//   mid_z = z + dists / 2, dists = diff(z) with sample_dist appended;  p = o + d mid_z;  r = clip(|p|, 1, 1e10)
//   the NeRF is queried at the inverted-sphere coordinates (p / r, 1 / r) and the ray direction           hn_outside_points
//   alpha = 1 - exp(-softplus(density) dists);  w = alpha * cumprod(1 - alpha + 1e-7, exclusive);  c = sigmoid(raw rgb)
//   colour = sum w c (+ background (1 - sum w))                                                    hn_outside_composite_fwd/_bwd
// One warp per ray, lanes over the samples in chunks of 32 (any n), the transmittance carried from chunk to chunk; the backward
// walks the chunks in reverse with the suffix sum of g_k w_k carried.  HBM-bound: 24 B in / 36 B out per sample forward.
#include <algorithm>

#include "common.cuh"

namespace hn {

constexpr int OUT_WARPS = 4;
constexpr int OUT_MAX_CHUNKS = 16;      // backward: n <= 512 samples per ray (the per-lane transmittances are kept in registers)

__device__ __forceinline__ float out_excl_prod(float f, int lane, float* total) {
    float inc = f;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const float t = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc *= t;
    }
    *total = __shfl_sync(0xffffffffu, inc, 31);
    const float ex = __shfl_up_sync(0xffffffffu, inc, 1);
    return lane == 0 ? 1.0f : ex;
}
__device__ __forceinline__ float out_suffix_sum(float v, int lane, float* total) {
    float inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const float t = __shfl_down_sync(0xffffffffu, inc, o);
        if (lane + o < 32) inc += t;
    }
    *total = __shfl_sync(0xffffffffu, inc, 0);
    return inc;
}
// F.softplus (beta = 1, threshold = 20) and its derivative
__device__ __forceinline__ float softplus1(float x) { return x > 20.0f ? x : log1pf(expf(x)); }
__device__ __forceinline__ float softplus1_grad(float x) { return x > 20.0f ? 1.0f : 1.0f / (1.0f + expf(-x)); }

// pts4 [B*n, 4] = (p / r, 1 / r), dirs [B*n, 3] = d, dists [B, n]
__global__ void outside_points_kernel(const float* __restrict__ rays_o, const float* __restrict__ rays_d,
                                      const float* __restrict__ z_vals, float sample_dist, int64_t n_rays, int n,
                                      float* __restrict__ pts4, float* __restrict__ dirs, float* __restrict__ dists) {
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= n_rays * n) return;
    const int64_t ray = idx / n;
    const int i = (int)(idx - ray * n);
    const float z = z_vals[idx];
    const float dist = i + 1 < n ? z_vals[idx + 1] - z : sample_dist;
    const float mid = z + dist * 0.5f;
    const float ox = rays_o[ray * 3], oy = rays_o[ray * 3 + 1], oz = rays_o[ray * 3 + 2];
    const float dx = rays_d[ray * 3], dy = rays_d[ray * 3 + 1], dz = rays_d[ray * 3 + 2];
    const float px = ox + dx * mid, py = oy + dy * mid, pz = oz + dz * mid;
    const float r = fminf(fmaxf(sqrtf(px * px + py * py + pz * pz), 1.0f), 1e10f);
    *reinterpret_cast<float4*>(pts4 + idx * 4) = make_float4(px / r, py / r, pz / r, 1.0f / r);
    dirs[idx * 3] = dx; dirs[idx * 3 + 1] = dy; dirs[idx * 3 + 2] = dz;
    dists[idx] = dist;
}

__global__ void __launch_bounds__(OUT_WARPS * 32) outside_composite_fwd_kernel(
    const float* __restrict__ density, const float* __restrict__ raw_rgb, const float* __restrict__ dists,
    const float* __restrict__ background, int64_t n_rays, int n, float* __restrict__ sampled_color,
    float* __restrict__ alpha_out, float* __restrict__ weights, float* __restrict__ color) {
    const int lane = threadIdx.x & 31;
    const int64_t ray = (int64_t)blockIdx.x * OUT_WARPS + (threadIdx.x >> 5);
    if (ray >= n_rays) return;
    float T = 1.0f, cr = 0.f, cg = 0.f, cb = 0.f, ws = 0.f;
    for (int c0 = 0; c0 < n; c0 += 32) {
        const int i = c0 + lane;
        const bool ok = i < n;
        const int64_t s = ray * n + i;
        float a = 0.0f, r = 0.f, g = 0.f, b = 0.f;
        if (ok) {
            a = 1.0f - expf(-softplus1(density[s]) * dists[s]);
            r = sigmoidf_(raw_rgb[s * 3]); g = sigmoidf_(raw_rgb[s * 3 + 1]); b = sigmoidf_(raw_rgb[s * 3 + 2]);
        }
        float tot;
        const float Ti = T * out_excl_prod(ok ? 1.0f - a + 1e-7f : 1.0f, lane, &tot);
        T *= tot;
        if (ok) {
            const float w = a * Ti;
            alpha_out[s] = a;
            weights[s] = w;
            sampled_color[s * 3] = r; sampled_color[s * 3 + 1] = g; sampled_color[s * 3 + 2] = b;
            cr += w * r; cg += w * g; cb += w * b; ws += w;
        }
    }
    cr = warp_sum(cr); cg = warp_sum(cg); cb = warp_sum(cb); ws = warp_sum(ws);
    if (lane == 0) {
        if (background) {
            cr += background[0] * (1.0f - ws); cg += background[1] * (1.0f - ws); cb += background[2] * (1.0f - ws);
        }
        color[ray * 3] = cr; color[ray * 3 + 1] = cg; color[ray * 3 + 2] = cb;
    }
}

// cotangents of the four outputs (any may be NULL) -> d_density, d_raw_rgb
__global__ void __launch_bounds__(OUT_WARPS * 32) outside_composite_bwd_kernel(
    const float* __restrict__ density, const float* __restrict__ dists, const float* __restrict__ background,
    const float* __restrict__ sampled_color, const float* __restrict__ alpha, const float* __restrict__ weights,
    int64_t n_rays, int n, const float* __restrict__ g_color, const float* __restrict__ g_sampled,
    const float* __restrict__ g_alpha, const float* __restrict__ g_weights, float* __restrict__ d_density,
    float* __restrict__ d_raw) {
    const int lane = threadIdx.x & 31;
    const int64_t ray = (int64_t)blockIdx.x * OUT_WARPS + (threadIdx.x >> 5);
    if (ray >= n_rays) return;
    float gc[3] = {0.f, 0.f, 0.f};
    if (g_color) { gc[0] = g_color[ray * 3]; gc[1] = g_color[ray * 3 + 1]; gc[2] = g_color[ray * 3 + 2]; }
    const float gbg = background ? gc[0] * background[0] + gc[1] * background[1] + gc[2] * background[2] : 0.0f;
    // transmittance in front of every sample, recomputed like the forward (w / alpha is undefined where alpha == 0)
    float Tl[OUT_MAX_CHUNKS];
    {
        float T = 1.0f;
#pragma unroll
        for (int ch = 0; ch < OUT_MAX_CHUNKS; ++ch) {
            const int i = ch * 32 + lane;
            Tl[ch] = 0.0f;
            if (ch * 32 < n) {
                float tot;
                const float f = i < n ? 1.0f - alpha[ray * n + i] + 1e-7f : 1.0f;
                Tl[ch] = T * out_excl_prod(f, lane, &tot);
                T *= tot;
            }
        }
    }
    float later = 0.0f;                 // sum of G_k w_k over the samples of the chunks already visited (k beyond this chunk)
#pragma unroll
    for (int ch = OUT_MAX_CHUNKS - 1; ch >= 0; --ch) {
        if (ch * 32 >= n) continue;
        const int i = ch * 32 + lane;
        const bool ok = i < n;
        const int64_t s = ray * n + i;
        float G = 0.0f, w = 0.0f, a = 0.0f, c[3] = {0.f, 0.f, 0.f};
        if (ok) {
            w = weights[s]; a = alpha[s];
            c[0] = sampled_color[s * 3]; c[1] = sampled_color[s * 3 + 1]; c[2] = sampled_color[s * 3 + 2];
            G = gc[0] * c[0] + gc[1] * c[1] + gc[2] * c[2] - gbg + (g_weights ? g_weights[s] : 0.0f);
        }
        float tot;
        const float incl = out_suffix_sum(G * w, lane, &tot);       // this chunk's samples >= lane
        if (ok) {
            const float f = 1.0f - a + 1e-7f;
            const float beyond = incl - G * w + later;              // sum over k > i of G_k w_k
            const float dalpha = (g_alpha ? g_alpha[s] : 0.0f) + G * Tl[ch] - beyond / f;
            const float sp_d = softplus1(density[s]);
            d_density[s] = dalpha * dists[s] * expf(-sp_d * dists[s]) * softplus1_grad(density[s]);
#pragma unroll
            for (int q = 0; q < 3; ++q) {
                const float dc = (g_sampled ? g_sampled[s * 3 + q] : 0.0f) + w * gc[q];
                d_raw[s * 3 + q] = dc * c[q] * (1.0f - c[q]);
            }
        }
        later += tot;
    }
}

}  // namespace hn

using namespace hn;

extern "C" {

int hn_outside_points(const float* rays_o, const float* rays_d, const float* z_vals, float sample_dist, int64_t n_rays, int n,
                      float* pts4, float* dirs, float* dists, hn_stream_t stream) {
    HN_REQUIRE(n_rays >= 0 && n >= 1, "hn_outside_points: bad sizes");
    if (n_rays == 0) return HN_OK;
    HN_REQUIRE(rays_o && rays_d && z_vals && pts4 && dirs && dists && aligned16(pts4), "hn_outside_points: null or misaligned pointer");
    const int64_t total = n_rays * n;
    outside_points_kernel<<<(unsigned)ceil_div(total, 256), 256, 0, (cudaStream_t)stream>>>(rays_o, rays_d, z_vals, sample_dist, n_rays,
                                                                                             n, pts4, dirs, dists);
    count_launch();
    HN_CHECK_LAUNCH();
    return HN_OK;
}

int hn_outside_composite_fwd(const float* density, const float* raw_rgb, const float* dists, const float* background,
                             int64_t n_rays, int n, float* sampled_color, float* alpha, float* weights, float* color,
                             hn_stream_t stream) {
    HN_REQUIRE(n_rays >= 0 && n >= 1, "hn_outside_composite_fwd: bad sizes");
    if (n_rays == 0) return HN_OK;
    HN_REQUIRE(density && raw_rgb && dists && sampled_color && alpha && weights && color, "hn_outside_composite_fwd: null pointer");
    outside_composite_fwd_kernel<<<(unsigned)ceil_div(n_rays, OUT_WARPS), OUT_WARPS * 32, 0, (cudaStream_t)stream>>>(
        density, raw_rgb, dists, background, n_rays, n, sampled_color, alpha, weights, color);
    count_launch();
    HN_CHECK_LAUNCH();
    return HN_OK;
}

int hn_outside_composite_bwd(const float* density, const float* dists, const float* background, const float* sampled_color,
                             const float* alpha, const float* weights, int64_t n_rays, int n, const float* g_color,
                             const float* g_sampled_color, const float* g_alpha, const float* g_weights, float* d_density,
                             float* d_raw_rgb, hn_stream_t stream) {
    HN_REQUIRE(n_rays >= 0 && n >= 1 && n <= 32 * OUT_MAX_CHUNKS, "hn_outside_composite_bwd: bad sizes (n <= 512 samples per ray)");
    if (n_rays == 0) return HN_OK;
    HN_REQUIRE(density && dists && sampled_color && alpha && weights && d_density && d_raw_rgb, "hn_outside_composite_bwd: null pointer");
    outside_composite_bwd_kernel<<<(unsigned)ceil_div(n_rays, OUT_WARPS), OUT_WARPS * 32, 0, (cudaStream_t)stream>>>(
        density, dists, background, sampled_color, alpha, weights, n_rays, n, g_color, g_sampled_color, g_alpha, g_weights,
        d_density, d_raw_rgb);
    count_launch();
    HN_CHECK_LAUNCH();
    return HN_OK;
}

}  // extern "C"
