// Ray helpers and hierarchical importance sampling: replaces utils/renderer.py:10-37 (sample_pdf),
// :60-86 (up_sample), :88-105 (cat_z_vals) and the point generation at :91,:124,:216.
// All of it is HBM/latency-bound index and scan work: one warp (or thread) per ray, coalesced rows.
// Arithmetic that decides sample positions uses explicitly rounded intrinsics (no FMA contraction)
// so that positions are bit-identical to eager PyTorch given the same inputs.
#include "common.cuh"

namespace hn {

__global__ void ray_points_kernel(const float* __restrict__ o, const float* __restrict__ d,
                                  const float* __restrict__ z, int64_t total, int n,
                                  float* __restrict__ pts) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    int64_t b = i / n;
    float zz = z[i];
#pragma unroll
    for (int c = 0; c < 3; ++c) pts[i * 3 + c] = __fadd_rn(o[b * 3 + c], __fmul_rn(d[b * 3 + c], zz));
}

__global__ void mid_points_kernel(const float* __restrict__ o, const float* __restrict__ d,
                                  const float* __restrict__ z, int64_t total, int n, float sample_dist,
                                  float* __restrict__ pts, float* __restrict__ dists, float* __restrict__ dirs) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    int64_t b = i / n;
    int k = (int)(i - b * n);
    float zi = z[i];
    float dist = k + 1 < n ? __fsub_rn(z[i + 1], zi) : sample_dist;
    float mid = __fadd_rn(zi, __fmul_rn(dist, 0.5f));
    dists[i] = dist;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        float dc = d[b * 3 + c];
        pts[i * 3 + c] = __fadd_rn(o[b * 3 + c], __fmul_rn(dc, mid));
        if (dirs) dirs[i * 3 + c] = dc;          // rays_d[:, None, :].expand(B, n, 3) (utils/renderer.py:127)
    }
}

// Backward of pts = o + d * mid and dirs = expand(d): one warp per ray walks the ray's contiguous [n, 3] cotangent rows
// (coalesced), d_o = sum_i g_i, d_d = sum_i (g_i * mid_i + gdirs_i).  Replaces two slow strided torch reductions, a
// multiply and the expand's reduction per render.
__global__ void __launch_bounds__(256)
mid_points_bwd_kernel(const float* __restrict__ d_pts, const float* __restrict__ d_dirs, const float* __restrict__ z,
                      const float* __restrict__ dists, int64_t n_rays, int n, float* __restrict__ d_o,
                      float* __restrict__ d_d) {
    int64_t b = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    int lane = threadIdx.x & 31;
    if (b >= n_rays) return;
    const float* g = d_pts + b * n * 3;
    const float* gd = d_dirs ? d_dirs + b * n * 3 : nullptr;
    float ao[3] = {0.f, 0.f, 0.f}, ad[3] = {0.f, 0.f, 0.f};
    for (int e = lane; e < n * 3; e += 32) {
        int i = e / 3, c = e - i * 3;
        float gv = g[e];
        float mid = z[b * n + i] + dists[b * n + i] * 0.5f;
        float dv = gv * mid + (gd ? gd[e] : 0.f);
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            ao[k] += (c == k) ? gv : 0.f;
            ad[k] += (c == k) ? dv : 0.f;
        }
    }
#pragma unroll
    for (int k = 0; k < 3; ++k) { ao[k] = warp_sum(ao[k]); ad[k] = warp_sum(ad[k]); }
    if (lane == 0) {
#pragma unroll
        for (int k = 0; k < 3; ++k) { d_o[b * 3 + k] = ao[k]; d_d[b * 3 + k] = ad[k]; }
    }
}

// Rays into the object frame (utils/renderer.py:180-188): o' = Ro (o - To), d' = Ro d.  Ro [3,3] row-major, To [3]:
// device memory (they are trained: se3_refine / the fitting pose), so nothing is read on the host.
__global__ void rays_to_local_kernel(const float* __restrict__ o, const float* __restrict__ d,
                                     const float* __restrict__ Ro, const float* __restrict__ To, int64_t n_rays,
                                     float* __restrict__ lo, float* __restrict__ ld) {
    int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= n_rays) return;
    float R[9], T[3], ov[3], dv[3];
#pragma unroll
    for (int k = 0; k < 9; ++k) R[k] = Ro[k];
#pragma unroll
    for (int k = 0; k < 3; ++k) { T[k] = To[k]; ov[k] = __fsub_rn(o[b * 3 + k], T[k]); dv[k] = d[b * 3 + k]; }
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        lo[b * 3 + i] = R[i * 3] * ov[0] + R[i * 3 + 1] * ov[1] + R[i * 3 + 2] * ov[2];
        ld[b * 3 + i] = R[i * 3] * dv[0] + R[i * 3 + 1] * dv[1] + R[i * 3 + 2] * dv[2];
    }
}

// Its backward in ONE launch (one CTA, fixed summation order -> deterministic):
//   d_Ro[i,j] = sum_b g_lo[b,i] (o[b,j] - To[j]) + g_ld[b,i] d[b,j],   d_To[j] = - sum_b sum_i Ro[i,j] g_lo[b,i],
//   d_o[b] = Ro^T g_lo[b],  d_d[b] = Ro^T g_ld[b]  (optional).
__global__ void __launch_bounds__(256)
rays_to_local_bwd_kernel(const float* __restrict__ g_lo, const float* __restrict__ g_ld, const float* __restrict__ o,
                         const float* __restrict__ d, const float* __restrict__ Ro, const float* __restrict__ To,
                         int64_t n_rays, float* __restrict__ d_Ro, float* __restrict__ d_To, float* __restrict__ d_o,
                         float* __restrict__ d_d) {
    __shared__ float red[12][8];
    float R[9], T[3];
#pragma unroll
    for (int k = 0; k < 9; ++k) R[k] = Ro[k];
#pragma unroll
    for (int k = 0; k < 3; ++k) T[k] = To[k];
    float acc[12];
#pragma unroll
    for (int k = 0; k < 12; ++k) acc[k] = 0.f;
    for (int64_t b = threadIdx.x; b < n_rays; b += blockDim.x) {
        float go[3] = {0.f, 0.f, 0.f}, gd[3] = {0.f, 0.f, 0.f}, ov[3], dv[3];
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            if (g_lo) go[k] = g_lo[b * 3 + k];
            if (g_ld) gd[k] = g_ld[b * 3 + k];
            ov[k] = __fsub_rn(o[b * 3 + k], T[k]);
            dv[k] = d[b * 3 + k];
        }
#pragma unroll
        for (int i = 0; i < 3; ++i)
#pragma unroll
            for (int j = 0; j < 3; ++j) acc[i * 3 + j] += go[i] * ov[j] + gd[i] * dv[j];
#pragma unroll
        for (int j = 0; j < 3; ++j) {
            float to = R[j] * go[0] + R[3 + j] * go[1] + R[6 + j] * go[2];        // (Ro^T g_lo)[j]
            acc[9 + j] -= to;
            if (d_o) d_o[b * 3 + j] = to;
            if (d_d) d_d[b * 3 + j] = R[j] * gd[0] + R[3 + j] * gd[1] + R[6 + j] * gd[2];
        }
    }
    int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int k = 0; k < 12; ++k) {
        float v = warp_sum(acc[k]);
        if (lane == 0) red[k][warp] = v;
    }
    __syncthreads();
    if (threadIdx.x < 12) {
        float v = 0.f;
        for (int w = 0; w < 8; ++w) v += red[threadIdx.x][w];
        if (threadIdx.x < 9) d_Ro[threadIdx.x] = v; else d_To[threadIdx.x - 9] = v;
    }
}

// searchsorted(cdf, u, right=True): first index with cdf[idx] > u  (in [0, m])
__device__ __forceinline__ int upper_bound(const float* cdf, int m, float u) {
    int lo = 0, hi = m;
    while (lo < hi) {
        int mid = (lo + hi) >> 1;
        if (cdf[mid] > u) hi = mid; else lo = mid + 1;
    }
    return lo;
}

__device__ __forceinline__ float invert_cdf_at(const float* bins, const float* cdf, int m, float u,
                                               int* below_out, int* above_out) {
    int inds = upper_bound(cdf, m, u);
    int below = max(0, inds - 1);
    int above = min(m - 1, inds);
    float cb = cdf[below], ca = cdf[above];
    float bb = bins[below], ba = bins[above];
    float denom = __fsub_rn(ca, cb);
    if (denom < 1e-5f) denom = 1.0f;
    float t = __fdiv_rn(__fsub_rn(u, cb), denom);
    if (below_out) *below_out = below;
    if (above_out) *above_out = above;
    return __fadd_rn(bb, __fmul_rn(t, __fsub_rn(ba, bb)));
}

__global__ void inverse_cdf_kernel(const float* __restrict__ bins, const float* __restrict__ cdf,
                                   const float* __restrict__ u, int64_t n_rays, int m, int ns,
                                   float* __restrict__ samples, int64_t* __restrict__ below,
                                   int64_t* __restrict__ above) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_rays * ns) return;
    int64_t b = i / ns;
    int k = (int)(i - b * ns);
    int lo, hi;
    samples[i] = invert_cdf_at(bins + b * m, cdf + b * m, m, u[k], &lo, &hi);
    if (below) below[i] = lo;
    if (above) above[i] = hi;
}

// One warp per ray.  smem per warp: z[m], w/cdf[m].
constexpr int UP_WARPS = 4;
__global__ void __launch_bounds__(UP_WARPS * 32) up_sample_kernel(
    const float* __restrict__ z, const float* __restrict__ sdf, const float* __restrict__ u,
    int64_t n_rays, int m, int n_imp, float inv_s, float* __restrict__ new_z) {
    extern __shared__ float smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    int64_t ray = (int64_t)blockIdx.x * UP_WARPS + warp;
    if (ray >= n_rays) return;
    float* sz = smem + warp * 3 * m;
    float* ss = sz + m;      // sdf, then reused
    float* sc = ss + m;      // alpha -> weights -> cdf
    for (int j = lane; j < m; j += 32) { sz[j] = z[ray * m + j]; ss[j] = sdf[ray * m + j]; }
    __syncwarp();
    // section alphas (utils/renderer.py:64-81)
    for (int j = lane; j < m - 1; j += 32) {
        float ps = ss[j], ns = ss[j + 1], pz = sz[j], nz = sz[j + 1];
        float mid = __fmul_rn(__fadd_rn(ps, ns), 0.5f);
        float dist = __fsub_rn(nz, pz);
        float cosv = __fdiv_rn(__fsub_rn(ns, ps), __fadd_rn(dist, 1e-5f));
        float prevc = 0.0f;
        if (j > 0) {
            float pps = ss[j - 1], ppz = sz[j - 1];
            prevc = __fdiv_rn(__fsub_rn(ps, pps), __fadd_rn(__fsub_rn(pz, ppz), 1e-5f));
        }
        cosv = fminf(prevc, cosv);
        cosv = fminf(fmaxf(cosv, -1e3f), 0.0f);
        float half = __fmul_rn(__fmul_rn(cosv, dist), 0.5f);
        float prev_est = __fsub_rn(mid, half);
        float next_est = __fadd_rn(mid, half);
        float pc = sigmoidf_(__fmul_rn(prev_est, inv_s));
        float nc = sigmoidf_(__fmul_rn(next_est, inv_s));
        sc[j] = __fdiv_rn(__fadd_rn(__fsub_rn(pc, nc), 1e-5f), __fadd_rn(pc, 1e-5f));
    }
    __syncwarp();
    // transmittance cumprod and pdf cumsum: fp64 running value, every output rounded to fp32
    // (what torch's CPU cumprod/cumsum do -- SURVEY appendix B); sequential on lane 0.
    if (lane == 0) {
        double T = 1.0;
        double wsum = 0.0;
        for (int j = 0; j < m - 1; ++j) {
            float a = sc[j];
            float w = __fadd_rn(__fmul_rn(a, (float)T), 1e-5f);   // weights + 1e-5
            T *= (double)__fadd_rn(__fsub_rn(1.0f, a), 1e-7f);
            sc[j] = w;
            wsum += (double)w;
        }
        float tot = (float)wsum;
        double c = 0.0;
        float prev = 0.0f;     // cdf[0] = 0, shift by one while writing
        for (int j = 0; j < m - 1; ++j) {
            float pdf = __fdiv_rn(sc[j], tot);
            c += (double)pdf;
            sc[j] = prev;
            prev = (float)c;
        }
        sc[m - 1] = prev;
    }
    __syncwarp();
    for (int k = lane; k < n_imp; k += 32)
        new_z[ray * n_imp + k] = invert_cdf_at(sz, sc, m, u[k], nullptr, nullptr);
}

// stable merge of two sorted rows by rank
__global__ void merge_sorted_kernel(const float* __restrict__ za, int m, const float* __restrict__ zb,
                                    int k, int64_t n_rays, float* __restrict__ zo,
                                    int64_t* __restrict__ index, const float* __restrict__ sa,
                                    const float* __restrict__ sb, int64_t row_mod,
                                    float* __restrict__ so) {
    const int tot = m + k;
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_rays * tot) return;
    int64_t b = i / tot;
    int e = (int)(i - b * tot);
    const float* ra = za + b * m;
    const float* rb = zb + b * k;
    float v;
    int rank;
    if (e < m) {
        v = ra[e];
        int lo = 0, hi = k;              // number of b elements strictly less than v
        while (lo < hi) { int mid = (lo + hi) >> 1; if (rb[mid] < v) lo = mid + 1; else hi = mid; }
        rank = e + lo;
    } else {
        int j = e - m;
        v = rb[j];
        int lo = 0, hi = m;              // number of a elements <= v
        while (lo < hi) { int mid = (lo + hi) >> 1; if (ra[mid] <= v) lo = mid + 1; else hi = mid; }
        rank = j + lo;
    }
    zo[b * tot + rank] = v;
    if (index) index[b * tot + rank] = e;
    if (so) {
        int64_t src = row_mod > 0 ? b % row_mod : b;
        so[b * tot + rank] = e < m ? sa[src * m + e] : sb[src * k + (e - m)];
    }
}

// stable rank sort of each row; one block per row
__global__ void sort_rows_kernel(const float* __restrict__ x, int n, float* __restrict__ out,
                                 int64_t* __restrict__ index) {
    extern __shared__ float row[];
    int64_t b = blockIdx.x;
    for (int i = threadIdx.x; i < n; i += blockDim.x) row[i] = x[b * n + i];
    __syncthreads();
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        float v = row[i];
        int rank = 0;
        for (int j = 0; j < n; ++j) {
            float w = row[j];
            rank += (w < v) || (w == v && j < i);
        }
        out[b * n + rank] = v;
        if (index) index[b * n + rank] = i;
    }
}

}  // namespace hn

using namespace hn;

extern "C" {

int hn_ray_points(const float* rays_o, const float* rays_d, const float* z, int64_t n_rays, int n,
                  float* pts, hn_stream_t stream) {
    HN_REQUIRE(n_rays >= 0 && n > 0, "hn_ray_points: bad sizes");
    if (n_rays == 0) return HN_OK;
    HN_REQUIRE(rays_o && rays_d && z && pts, "hn_ray_points: null pointer");
    int64_t total = n_rays * n;
    ray_points_kernel<<<(unsigned)ceil_div(total, 256), 256, 0, (cudaStream_t)stream>>>(rays_o, rays_d, z, total, n, pts);
    count_launch();
    HN_CHECK_LAUNCH();
    return HN_OK;
}

int hn_mid_points(const float* rays_o, const float* rays_d, const float* z, int64_t n_rays, int n,
                  float sample_dist, float* pts, float* dists, float* dirs, hn_stream_t stream) {
    HN_REQUIRE(n_rays >= 0 && n > 0, "hn_mid_points: bad sizes");
    if (n_rays == 0) return HN_OK;
    HN_REQUIRE(rays_o && rays_d && z && pts && dists, "hn_mid_points: null pointer");
    int64_t total = n_rays * n;
    mid_points_kernel<<<(unsigned)ceil_div(total, 256), 256, 0, (cudaStream_t)stream>>>(rays_o, rays_d, z, total, n,
                                                                                       sample_dist, pts, dists, dirs);
    count_launch();
    HN_CHECK_LAUNCH();
    return HN_OK;
}

int hn_mid_points_bwd(const float* d_pts, const float* d_dirs, const float* z, const float* dists, int64_t n_rays,
                      int n, float* d_rays_o, float* d_rays_d, hn_stream_t stream) {
    HN_REQUIRE(n_rays >= 0 && n > 0, "hn_mid_points_bwd: bad sizes");
    if (n_rays == 0) return HN_OK;
    HN_REQUIRE(d_pts && z && dists && d_rays_o && d_rays_d, "hn_mid_points_bwd: null pointer");
    mid_points_bwd_kernel<<<(unsigned)ceil_div(n_rays * 32, 256), 256, 0, (cudaStream_t)stream>>>(
        d_pts, d_dirs, z, dists, n_rays, n, d_rays_o, d_rays_d);
    count_launch();
    HN_CHECK_LAUNCH();
    return HN_OK;
}

int hn_rays_to_local(const float* rays_o, const float* rays_d, const float* Ro, const float* To, int64_t n_rays,
                     float* local_o, float* local_d, hn_stream_t stream) {
    HN_REQUIRE(n_rays >= 0, "hn_rays_to_local: bad sizes");
    if (n_rays == 0) return HN_OK;
    HN_REQUIRE(rays_o && rays_d && Ro && To && local_o && local_d, "hn_rays_to_local: null pointer");
    rays_to_local_kernel<<<(unsigned)ceil_div(n_rays, 256), 256, 0, (cudaStream_t)stream>>>(rays_o, rays_d, Ro, To, n_rays,
                                                                                             local_o, local_d);
    count_launch();
    HN_CHECK_LAUNCH();
    return HN_OK;
}

int hn_rays_to_local_bwd(const float* d_local_o, const float* d_local_d, const float* rays_o, const float* rays_d,
                         const float* Ro, const float* To, int64_t n_rays, float* d_Ro, float* d_To, float* d_rays_o,
                         float* d_rays_d, hn_stream_t stream) {
    HN_REQUIRE(n_rays >= 0, "hn_rays_to_local_bwd: bad sizes");
    HN_REQUIRE(rays_o && rays_d && Ro && To && d_Ro && d_To, "hn_rays_to_local_bwd: null pointer");
    rays_to_local_bwd_kernel<<<1, 256, 0, (cudaStream_t)stream>>>(d_local_o, d_local_d, rays_o, rays_d, Ro, To, n_rays,
                                                                  d_Ro, d_To, d_rays_o, d_rays_d);
    count_launch();
    HN_CHECK_LAUNCH();
    return HN_OK;
}

int hn_up_sample(const float* z, const float* sdf, const float* u, int64_t n_rays, int m,
                 int n_importance, float inv_s, float* new_z, hn_stream_t stream) {
    HN_REQUIRE(n_rays >= 0 && m >= 2 && m <= 2048 && n_importance > 0, "hn_up_sample: bad sizes (m=%d)", m);
    if (n_rays == 0) return HN_OK;
    HN_REQUIRE(z && sdf && u && new_z, "hn_up_sample: null pointer");
    size_t smem = (size_t)UP_WARPS * 3 * m * sizeof(float);
    if (smem > 48 * 1024)
        HN_CHECK_CUDA(cudaFuncSetAttribute(up_sample_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    up_sample_kernel<<<(unsigned)ceil_div(n_rays, UP_WARPS), UP_WARPS * 32, smem, (cudaStream_t)stream>>>(
        z, sdf, u, n_rays, m, n_importance, inv_s, new_z);
    count_launch();
    HN_CHECK_LAUNCH();
    return HN_OK;
}

int hn_inverse_cdf(const float* bins, const float* cdf, const float* u, int64_t n_rays, int m,
                   int n_samples, float* samples, int64_t* below, int64_t* above, hn_stream_t stream) {
    HN_REQUIRE(n_rays >= 0 && m >= 1 && n_samples > 0, "hn_inverse_cdf: bad sizes");
    if (n_rays == 0) return HN_OK;
    HN_REQUIRE(bins && cdf && u && samples, "hn_inverse_cdf: null pointer");
    int64_t total = n_rays * n_samples;
    inverse_cdf_kernel<<<(unsigned)ceil_div(total, 256), 256, 0, (cudaStream_t)stream>>>(bins, cdf, u, n_rays, m, n_samples,
                                                                                        samples, below, above);
    count_launch();
    HN_CHECK_LAUNCH();
    return HN_OK;
}

int hn_merge_sorted(const float* z_a, int m, const float* z_b, int k, int64_t n_rays, float* z_out,
                    int64_t* index, const float* sdf_a, const float* sdf_b, int64_t sdf_row_mod,
                    float* sdf_out, hn_stream_t stream) {
    HN_REQUIRE(n_rays >= 0 && m >= 0 && k >= 0 && m + k > 0, "hn_merge_sorted: bad sizes");
    if (n_rays == 0) return HN_OK;
    HN_REQUIRE(z_a && z_b && z_out, "hn_merge_sorted: null pointer");
    HN_REQUIRE(!sdf_out || (sdf_a && sdf_b), "hn_merge_sorted: sdf_out needs sdf_a and sdf_b");
    int64_t total = n_rays * (m + k);
    merge_sorted_kernel<<<(unsigned)ceil_div(total, 256), 256, 0, (cudaStream_t)stream>>>(
        z_a, m, z_b, k, n_rays, z_out, index, sdf_a, sdf_b, sdf_row_mod, sdf_out);
    count_launch();
    HN_CHECK_LAUNCH();
    return HN_OK;
}

int hn_sort_rows(const float* x, int64_t n_rays, int n, float* out, int64_t* index, hn_stream_t stream) {
    HN_REQUIRE(n_rays >= 0 && n > 0 && n <= 4096, "hn_sort_rows: bad sizes");
    if (n_rays == 0) return HN_OK;
    HN_REQUIRE(x && out && x != out, "hn_sort_rows: null or aliased pointer");
    HN_REQUIRE(n_rays < (1ll << 31), "hn_sort_rows: too many rows for one launch");
    int threads = n <= 64 ? 64 : (n <= 128 ? 128 : 256);
    sort_rows_kernel<<<(unsigned)n_rays, threads, (size_t)n * sizeof(float), (cudaStream_t)stream>>>(x, n, out, index);
    count_launch();
    HN_CHECK_LAUNCH();
    return HN_OK;
}

}  // extern "C"
