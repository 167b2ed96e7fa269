// Loss epilogues of the render step: the caller-side loss code of the reference
// (exp_runner.py:206-227 training; fitting_single.py:253-283 and fitting_video.py:286-309 fitting) as ONE forward launch
// (all partial sums, deterministic two-level reduction, the last CTA finalises) and ONE elementwise backward launch
// that writes the cotangents the compositor backward consumes.  Replaces ~30 tiny torch launches per step.
// HBM/latency-bound scalar work: grid = min(#SM, blocks needed), coalesced loads, no atomics on the sums.
#include "common.cuh"

namespace hn {

constexpr int LOSS_THREADS = 256;
constexpr int LOSS_MAX_BLOCKS = 148 * 2;
constexpr int LOSS_SUMS = 4;

// Block-wide sum of K per-thread values; result valid in thread 0.
template <int K>
__device__ __forceinline__ void block_sum(float (&v)[K], float* smem /* [K][8] */) {
    int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int k = 0; k < K; ++k) {
        v[k] = warp_sum(v[k]);
        if (lane == 0) smem[k * 8 + warp] = v[k];
    }
    __syncthreads();
    if (threadIdx.x == 0) {
#pragma unroll
        for (int k = 0; k < K; ++k) {
            float s = 0.f;
            for (int w = 0; w < LOSS_THREADS / 32; ++w) s += smem[k * 8 + w];
            v[k] = s;
        }
    }
}

// Publishes this CTA's partial sums and tells whether it is the last CTA of the grid to do so.  `ws` holds
// [LOSS_MAX_BLOCKS][LOSS_SUMS] floats followed by one counter word that must be 0 before the first launch and is
// reset by the last CTA (self-cleaning across launches on one stream).
__device__ __forceinline__ bool publish_partials(const float (&v)[LOSS_SUMS], float* ws) {
    __shared__ bool is_last;
    if (threadIdx.x == 0) {
#pragma unroll
        for (int k = 0; k < LOSS_SUMS; ++k) ws[blockIdx.x * LOSS_SUMS + k] = v[k];
        __threadfence();
        unsigned* counter = reinterpret_cast<unsigned*>(ws + LOSS_MAX_BLOCKS * LOSS_SUMS);
        unsigned done = atomicAdd(counter, 1u);
        is_last = (done == gridDim.x - 1);
        if (is_last) *counter = 0u;
    }
    __syncthreads();
    return is_last;
}

// sums over CTAs in a fixed order with an fp64 running value (deterministic, independent of scheduling)
__device__ __forceinline__ void final_sums(const float* ws, double (&tot)[LOSS_SUMS]) {
    __threadfence();
#pragma unroll
    for (int k = 0; k < LOSS_SUMS; ++k) tot[k] = 0.0;
    for (unsigned b = 0; b < gridDim.x; ++b)
#pragma unroll
        for (int k = 0; k < LOSS_SUMS; ++k) tot[k] += (double)__ldcg(ws + b * LOSS_SUMS + k);
}

__device__ __forceinline__ float clip_p(float w) { return fminf(fmaxf(w, 1e-3f), 1.0f - 1e-3f); }

// out[0] total, [1] colour loss, [2] mask (BCE) loss, [3] psnr, [4] colour divisor used, [5] mask_sum, [6] eikonal term
__global__ void __launch_bounds__(LOSS_THREADS)
render_loss_fwd_kernel(const float* __restrict__ color, const float* __restrict__ wsum,
                       const float* __restrict__ true_rgb, const float* __restrict__ true_mask,
                       const float* __restrict__ grad_err, int64_t n, float color_div,
                       const float* __restrict__ color_div_dev, float color_weight, float mask_weight,
                       float igr_weight, float* __restrict__ ws, float* __restrict__ out) {
    __shared__ float red[LOSS_SUMS * 8];
    float v[LOSS_SUMS] = {0.f, 0.f, 0.f, 0.f};   // sum |err|, sum mask, sum BCE, sum err^2
    for (int64_t i = (int64_t)blockIdx.x * LOSS_THREADS + threadIdx.x; i < n; i += (int64_t)gridDim.x * LOSS_THREADS) {
        float m = true_mask[i];
        float l1 = 0.f, sq = 0.f;
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            float diff = __fsub_rn(color[i * 3 + c], true_rgb[i * 3 + c]);
            l1 += fabsf(__fmul_rn(diff, m));
            sq += __fmul_rn(__fmul_rn(diff, diff), m);
        }
        float p = clip_p(wsum[i]);
        // aten binary_cross_entropy: (t - 1) * max(log1p(-x), -100) - t * max(log(x), -100)
        float bce = (m - 1.0f) * fmaxf(log1pf(-p), -100.0f) - m * fmaxf(logf(p), -100.0f);
        v[0] += l1; v[1] += m; v[2] += bce; v[3] += sq;
    }
    block_sum<LOSS_SUMS>(v, red);
    if (!publish_partials(v, ws)) return;
    if (threadIdx.x == 0) {
        double tot[LOSS_SUMS];
        final_sums(ws, tot);
        float mask_sum = (float)tot[1] + 1e-5f;
        float div = color_div_dev ? *color_div_dev : (color_div > 0.f ? color_div : mask_sum);
        float color_loss = (float)tot[0] / div;
        float mask_loss = (float)(tot[2] / (double)n);
        float eik = grad_err ? *grad_err : 0.f;
        // exp_runner.py:222: 20 log10(1 / sqrt(sum((c - t)^2 m) / (mask_sum * 3)))
        float psnr = 20.0f * log10f(1.0f / sqrtf((float)tot[3] / (mask_sum * 3.0f)));
        out[0] = color_weight * color_loss + mask_weight * mask_loss + igr_weight * eik;
        out[1] = color_loss;
        out[2] = mask_loss;
        out[3] = psnr;
        out[4] = div;
        out[5] = mask_sum;
        out[6] = eik;
        out[7] = 0.f;
    }
}

__global__ void __launch_bounds__(LOSS_THREADS)
render_loss_bwd_kernel(const float* __restrict__ g_loss, const float* __restrict__ color,
                       const float* __restrict__ wsum, const float* __restrict__ true_rgb,
                       const float* __restrict__ true_mask, const float* __restrict__ fwd_out, int64_t n,
                       float color_weight, float mask_weight, float igr_weight, float* __restrict__ d_color,
                       float* __restrict__ d_wsum, float* __restrict__ d_grad_err) {
    int64_t i = (int64_t)blockIdx.x * LOSS_THREADS + threadIdx.x;
    float g = g_loss ? *g_loss : 1.0f;
    if (i == 0 && d_grad_err) *d_grad_err = g * igr_weight;
    if (i >= n) return;
    float m = true_mask[i];
    float gc = g * color_weight / fwd_out[4];
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        float e = __fmul_rn(__fsub_rn(color[i * 3 + c], true_rgb[i * 3 + c]), m);
        float s = e > 0.f ? 1.0f : (e < 0.f ? -1.0f : 0.0f);   // l1_loss backward = sign
        d_color[i * 3 + c] = gc * s * m;
    }
    float w = wsum[i];
    float p = clip_p(w);
    // aten binary_cross_entropy_backward: grad * (x - t) / max((1 - x) x, 1e-12); clamp passes on the closed interval
    float dp = (p - m) / fmaxf((1.0f - p) * p, 1e-12f);
    bool pass = (w >= 1e-3f) && (w <= 1.0f - 1e-3f);
    d_wsum[i] = pass ? g * mask_weight * dp / (float)n : 0.f;
}

// out[0] total = w_contact * contact + w_penet * penet, [1] contact loss, [2] penetration loss,
// [3] contact_num (+1e-9), [4] penet_num (+1e-9)
__global__ void __launch_bounds__(LOSS_THREADS)
interaction_loss_fwd_kernel(const float* __restrict__ sdf_h, int64_t ld_h, const float* __restrict__ sdf_o,
                            int64_t ld_o, int64_t n, float contact_thr, float w_contact, float w_penet,
                            float* __restrict__ ws, float* __restrict__ out) {
    __shared__ float red[LOSS_SUMS * 8];
    float v[LOSS_SUMS] = {0.f, 0.f, 0.f, 0.f};   // contact sum, contact count, penetration sum, penetration count
    for (int64_t i = (int64_t)blockIdx.x * LOSS_THREADS + threadIdx.x; i < n; i += (int64_t)gridDim.x * LOSS_THREADS) {
        float h = sdf_h[i * ld_h], o = sdf_o[i * ld_o];
        float a = __fadd_rn(fabsf(h), fabsf(o));
        if (a < contact_thr) { v[0] += a; v[1] += 1.0f; }
        if (o < 0.f && h < 0.f) { v[2] += a; v[3] += 1.0f; }
    }
    block_sum<LOSS_SUMS>(v, red);
    if (!publish_partials(v, ws)) return;
    if (threadIdx.x == 0) {
        double tot[LOSS_SUMS];
        final_sums(ws, tot);
        float cnum = (float)tot[1] + 1e-9f, pnum = (float)tot[3] + 1e-9f;
        float contact = (float)tot[0] / cnum, penet = (float)tot[2] / pnum;
        out[0] = w_contact * contact + w_penet * penet;
        out[1] = contact;
        out[2] = penet;
        out[3] = cnum;
        out[4] = pnum;
        out[5] = out[6] = out[7] = 0.f;
    }
}

__global__ void __launch_bounds__(LOSS_THREADS)
interaction_loss_bwd_kernel(const float* __restrict__ g_loss, const float* __restrict__ sdf_h, int64_t ld_h,
                            const float* __restrict__ sdf_o, int64_t ld_o, const float* __restrict__ fwd_out,
                            int64_t n, float contact_thr, float w_contact, float w_penet,
                            float* __restrict__ d_h, float* __restrict__ d_o) {
    int64_t i = (int64_t)blockIdx.x * LOSS_THREADS + threadIdx.x;
    if (i >= n) return;
    float g = g_loss ? *g_loss : 1.0f;
    float h = sdf_h[i * ld_h], o = sdf_o[i * ld_o];
    float a = __fadd_rn(fabsf(h), fabsf(o));
    float k = 0.f;
    if (a < contact_thr) k += w_contact / fwd_out[3];
    if (o < 0.f && h < 0.f) k += w_penet / fwd_out[4];
    k *= g;
    float sh = h > 0.f ? 1.0f : (h < 0.f ? -1.0f : 0.0f);      // abs backward = sign (0 at 0)
    float so = o > 0.f ? 1.0f : (o < 0.f ? -1.0f : 0.0f);
    d_h[i] = k * sh;
    d_o[i] = k * so;
}

static unsigned loss_grid(int64_t n) {
    int64_t blocks = ceil_div(n, LOSS_THREADS);
    int64_t cap = sm_count() > 0 ? (int64_t)sm_count() : 148;
    if (cap > LOSS_MAX_BLOCKS) cap = LOSS_MAX_BLOCKS;
    return (unsigned)(blocks < cap ? (blocks < 1 ? 1 : blocks) : cap);
}

}  // namespace hn

using namespace hn;

extern "C" {

int64_t hn_loss_ws_floats(void) { return LOSS_MAX_BLOCKS * LOSS_SUMS + 4; }

int hn_render_loss_fwd(const float* color, const float* weight_sum, const float* true_rgb, const float* true_mask,
                       const float* gradient_error, int64_t n_rays, float color_div, const float* color_div_dev,
                       float color_weight, float mask_weight, float igr_weight, float* ws, float* out,
                       hn_stream_t stream) {
    HN_REQUIRE(n_rays > 0, "hn_render_loss_fwd: n_rays must be positive (the reference's mean over 0 rays is NaN)");
    HN_REQUIRE(color && weight_sum && true_rgb && true_mask && ws && out, "hn_render_loss_fwd: null pointer");
    render_loss_fwd_kernel<<<loss_grid(n_rays), LOSS_THREADS, 0, (cudaStream_t)stream>>>(
        color, weight_sum, true_rgb, true_mask, gradient_error, n_rays, color_div, color_div_dev, color_weight,
        mask_weight, igr_weight, ws, out);
    count_launch();
    HN_CHECK_LAUNCH();
    return HN_OK;
}

int hn_render_loss_bwd(const float* g_loss, const float* color, const float* weight_sum, const float* true_rgb,
                       const float* true_mask, const float* fwd_out, int64_t n_rays, float color_weight,
                       float mask_weight, float igr_weight, float* d_color, float* d_weight_sum,
                       float* d_gradient_error, hn_stream_t stream) {
    HN_REQUIRE(n_rays > 0, "hn_render_loss_bwd: n_rays must be positive");
    HN_REQUIRE(color && weight_sum && true_rgb && true_mask && fwd_out && d_color && d_weight_sum,
               "hn_render_loss_bwd: null pointer");
    render_loss_bwd_kernel<<<(unsigned)ceil_div(n_rays, LOSS_THREADS), LOSS_THREADS, 0, (cudaStream_t)stream>>>(
        g_loss, color, weight_sum, true_rgb, true_mask, fwd_out, n_rays, color_weight, mask_weight, igr_weight,
        d_color, d_weight_sum, d_gradient_error);
    count_launch();
    HN_CHECK_LAUNCH();
    return HN_OK;
}

int hn_interaction_loss_fwd(const float* sdf_hand, int64_t ld_hand, const float* sdf_obj, int64_t ld_obj,
                            int64_t n_pts, float contact_thr, float w_contact, float w_penet, float* ws, float* out,
                            hn_stream_t stream) {
    HN_REQUIRE(n_pts > 0 && ld_hand >= 1 && ld_obj >= 1, "hn_interaction_loss_fwd: bad sizes");
    HN_REQUIRE(sdf_hand && sdf_obj && ws && out, "hn_interaction_loss_fwd: null pointer");
    interaction_loss_fwd_kernel<<<loss_grid(n_pts), LOSS_THREADS, 0, (cudaStream_t)stream>>>(
        sdf_hand, ld_hand, sdf_obj, ld_obj, n_pts, contact_thr, w_contact, w_penet, ws, out);
    count_launch();
    HN_CHECK_LAUNCH();
    return HN_OK;
}

int hn_interaction_loss_bwd(const float* g_loss, const float* sdf_hand, int64_t ld_hand, const float* sdf_obj,
                            int64_t ld_obj, const float* fwd_out, int64_t n_pts, float contact_thr, float w_contact,
                            float w_penet, float* d_sdf_hand, float* d_sdf_obj, hn_stream_t stream) {
    HN_REQUIRE(n_pts > 0 && ld_hand >= 1 && ld_obj >= 1, "hn_interaction_loss_bwd: bad sizes");
    HN_REQUIRE(sdf_hand && sdf_obj && fwd_out && d_sdf_hand && d_sdf_obj, "hn_interaction_loss_bwd: null pointer");
    interaction_loss_bwd_kernel<<<(unsigned)ceil_div(n_pts, LOSS_THREADS), LOSS_THREADS, 0, (cudaStream_t)stream>>>(
        g_loss, sdf_hand, ld_hand, sdf_obj, ld_obj, fwd_out, n_pts, contact_thr, w_contact, w_penet, d_sdf_hand,
        d_sdf_obj);
    count_launch();
    HN_CHECK_LAUNCH();
    return HN_OK;
}

}  // extern "C"
