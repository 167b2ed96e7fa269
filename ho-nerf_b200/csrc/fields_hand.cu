// Hand SDF field (HALO pose-conditioned, utils/fields.py:56-177) as ONE operator: value + feature +
// analytic normal, with the second-order backward to the weights, the points and the bone transforms;
// and the hand colour field (utils/fields.py:179-240).  Same layer-by-layer structure as fields_obj.cu:
// the 1386-wide HALO feature is produced once per point into the skip-input row [h3 (256) | feature
// (1386) | pad 2] (ld 1644), which serves as the input of layer 0 and of the skip layer 4.
#include <algorithm>

#include "common.cuh"
#include "fields_common.cuh"
#include "gemm_dispatch.cuh"
#include "halo.cuh"
#include "chain_hand_layout.cuh"
#include "gemm_bx3.cuh"

namespace hn {

namespace chain {      // chain16_hand.cu
int launch_hand16_trunk(const hn_mlp_t* m, const uint8_t* ops, int64_t n, const uint8_t* F16, const float* H0, const float* ZF4,
                        float* sdf, float* feat, int64_t ld_feat, uint8_t* const* EM, uint8_t* const* EML, cudaStream_t s);
int launch_hand16_nsweep(const hn_mlp_t* m, const uint8_t* ops, int64_t n, uint8_t* const* EM, uint8_t* const* EML,
                         uint8_t* const* D16, float* FB, cudaStream_t s);
int launch_hand16_bwd(const hn_mlp_t* m, const uint8_t* ops, int64_t n, uint8_t* const* EM, uint8_t* const* D16,
                      uint8_t* const* X16, const float* Q0, const float* QF4, const float* d_sdf, const float* d_feat,
                      int64_t ld_dfeat, float* DF, cudaStream_t s);
int hand16_pack(const hn_mlp_t* m, uint8_t* dst, cudaStream_t s);
// chain_color.cu: layers 1..3 + output of the hand colour net on the colour chain kernel (forward-only rendering)
int launch_color_tail_fwd(const hn_mlp_t* m, const uint8_t* ops, const float* Z0a, const float* Z0b, int64_t ld_z, int64_t n,
                          float* rgb, cudaStream_t s);
int64_t color_tail_bytes();
int color_tail_pack(const hn_mlp_t* m, uint8_t* dst, cudaStream_t s);
}  // namespace chain

constexpr int HROW_LD = 1644;        // [h3 256 | feature 1386 | pad 2]
constexpr int HFEAT_OFF = 256;
constexpr int HFB_LD = 1388;         // feature cotangent rows
constexpr int HALO_WARPS = 4;        // points per block (one warp per point)

__device__ __forceinline__ void load_x(const float* pts, int64_t p, float x[3]) {
    x[0] = pts[p * 3]; x[1] = pts[p * 3 + 1]; x[2] = pts[p * 3 + 2];
}

// feature rows: out[p, 0:1386] = F(x_p) (lanes 0..20 = joints), staged in smem, written coalesced
__global__ void __launch_bounds__(HALO_WARPS * 32) halo_feature_kernel(
    const float* __restrict__ pts, const float* __restrict__ bt_inv, const float* __restrict__ Tp, int64_t n,
    int64_t ppf, float* __restrict__ out, int64_t ld) {
    __shared__ float sm[HALO_WARPS][HALO_DIM + 2];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t p = (int64_t)blockIdx.x * HALO_WARPS + warp;
    if (p >= n) return;
    const int64_t f = p / ppf;
    float x[3];
    load_x(pts, p, x);
    if (lane < HALO_J) {
        HaloBase b = halo_base(bt_inv + (f * HALO_J + lane) * 16, Tp + (f * HALO_J + lane) * 3, x, lane);
        halo_feature(b, &sm[warp][lane * HALO_F]);
    }
    __syncwarp();
    float* o = out + p * ld;
    for (int i = lane; i < HALO_DIM; i += 32) o[i] = sm[warp][i];
    if (lane < 2) o[HALO_DIM + lane] = 0.0f;
}

// Same features, additionally (or only: out == NULL) as the fp16 hi / lo pair tiles the value trunk consumes straight from
// shared memory (chain16_hand.cu): per 128-point tile 22 k-blocks of 64 features, each {hi tile, lo tile} of [128 rows x 128 B]
// in the K-major SWIZZLE_128B UMMA layout; features 1386 .. 1407 and the rows past n are zero.  Launched over whole tiles.
__global__ void __launch_bounds__(HALO_WARPS * 32) halo_feature16_kernel(
    const float* __restrict__ pts, const float* __restrict__ bt_inv, const float* __restrict__ Tp, int64_t n,
    int64_t ppf, float* __restrict__ out, int64_t ld, uint8_t* __restrict__ F16) {
    constexpr int KPAD = chain::HAND_F_KBLOCKS * 64;      // 1408
    __shared__ __align__(16) float sm[HALO_WARPS][KPAD];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t p = (int64_t)blockIdx.x * HALO_WARPS + warp;
    const bool live = p < n;
    for (int i = HALO_DIM + lane; i < KPAD; i += 32) sm[warp][i] = 0.0f;
    if (live) {
        const int64_t f = p / ppf;
        float x[3];
        load_x(pts, p, x);
        if (lane < HALO_J) {
            HaloBase b = halo_base(bt_inv + (f * HALO_J + lane) * 16, Tp + (f * HALO_J + lane) * 3, x, lane);
            halo_feature<true>(b, &sm[warp][lane * HALO_F]);
        }
    } else {
        for (int i = lane; i < HALO_DIM; i += 32) sm[warp][i] = 0.0f;
    }
    __syncwarp();
    if (out && live) {
        float* o = out + p * ld;
        for (int i = lane; i < HALO_DIM; i += 32) o[i] = sm[warp][i];
        if (lane < 2) o[HALO_DIM + lane] = 0.0f;
    }
    uint8_t* tile = F16 + (size_t)(p >> 7) * chain::HAND_F16_TILE_BYTES;
    const uint32_t r = (uint32_t)(p & 127);
    for (int c = lane; c < KPAD / 8; c += 32) {          // 176 chunks of 8 features
        const float4 a = *reinterpret_cast<const float4*>(&sm[warp][c * 8]), bq = *reinterpret_cast<const float4*>(&sm[warp][c * 8 + 4]);
        uint4 hi, lo;
        chain::split2_lo16(a.x, a.y, hi.x, lo.x);
        chain::split2_lo16(a.z, a.w, hi.y, lo.y);
        chain::split2_lo16(bq.x, bq.y, hi.z, lo.z);
        chain::split2_lo16(bq.z, bq.w, hi.w, lo.w);
        uint8_t* dst = tile + (size_t)(c >> 3) * chain::HAND_F16_KB_BYTES + tc::sw128_offset(r, (uint32_t)(c & 7));
        *reinterpret_cast<uint4*>(dst) = hi;
        *reinterpret_cast<uint4*>(dst + 16384) = lo;
    }
}

// tangent rows: out[p, 0:1386] = F'(x_p) (R dn_p)
__global__ void __launch_bounds__(HALO_WARPS * 32) halo_tangent_kernel(
    const float* __restrict__ pts, const float* __restrict__ bt_inv, const float* __restrict__ Tp,
    const float* __restrict__ dn, int64_t n, int64_t ppf, float* __restrict__ out, int64_t ld) {
    __shared__ float sm[HALO_WARPS][HALO_DIM + 2];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t p = (int64_t)blockIdx.x * HALO_WARPS + warp;
    if (p >= n) return;
    const int64_t f = p / ppf;
    float x[3], t[3];
    load_x(pts, p, x);
    load_x(dn, p, t);
    if (lane < HALO_J) {
        const float* M = bt_inv + (f * HALO_J + lane) * 16;
        HaloBase b = halo_base(M, Tp + (f * HALO_J + lane) * 3, x, lane);
        float w[3];
#pragma unroll
        for (int a = 0; a < 3; ++a) w[a] = M[a * 4] * t[0] + M[a * 4 + 1] * t[1] + M[a * 4 + 2] * t[2];
        halo_jvp(b, w, &sm[warp][lane * HALO_F]);
    }
    __syncwarp();
    float* o = out + p * ld;
    for (int i = lane; i < HALO_DIM; i += 32) o[i] = sm[warp][i];
    if (lane < 2) o[HALO_DIM + lane] = 0.0f;
}

// normal[p] = sum_j R_j^T  dF_j/dq^T  FB[p, j]
__global__ void __launch_bounds__(HALO_WARPS * 32) halo_normal_kernel(
    const float* __restrict__ pts, const float* __restrict__ bt_inv, const float* __restrict__ Tp,
    const float* __restrict__ FB, int64_t ld_fb, int64_t n, int64_t ppf, float* __restrict__ normal) {
    __shared__ float sm[HALO_WARPS][HALO_DIM + 2];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t p = (int64_t)blockIdx.x * HALO_WARPS + warp;
    if (p >= n) return;
    const int64_t f = p / ppf;
    for (int i = lane; i < HALO_DIM; i += 32) sm[warp][i] = FB[p * ld_fb + i];
    __syncwarp();
    float x[3], acc[3] = {0.f, 0.f, 0.f};
    load_x(pts, p, x);
    if (lane < HALO_J) {
        const float* M = bt_inv + (f * HALO_J + lane) * 16;
        HaloBase b = halo_base(M, Tp + (f * HALO_J + lane) * 3, x, lane);
        float g[3], dummy[3], w[3] = {0.f, 0.f, 0.f};
        halo_grad_hvp<false>(b, &sm[warp][lane * HALO_F], w, g, dummy);
#pragma unroll
        for (int a = 0; a < 3; ++a) acc[a] = M[0 * 4 + a] * g[0] + M[1 * 4 + a] * g[1] + M[2 * 4 + a] * g[2];
    }
#pragma unroll
    for (int a = 0; a < 3; ++a) acc[a] = warp_sum(acc[a]);
    if (lane < 3) normal[p * 3 + lane] = lane == 0 ? acc[0] : (lane == 1 ? acc[1] : acc[2]);
}

// Backward through the embedding.  DF = cotangent of the feature (first order), FB = feature cotangent
// of the normal sweep, dn = cotangent of the normal (second order; may be NULL).
//   dq_j = F'^T DF_j + H_j(FB_j) (R_j dn);   d_x = sum_j R_j^T dq_j
//   dR_j += dq_j x^T + g_j(FB_j) dn^T ;  dt_j += dq_j ;  dT_j -= dq_j
// d_bt [frames, 21, 16] (row-major 4x4, last row untouched) and d_T [frames, 21, 3] are ACCUMULATED.
// Each warp walks a contiguous run of points and keeps the pose gradients of its joint (lane) in registers; they are
// flushed with one atomicAdd per entry when the frame changes and at the end of the run (the per-point atomics of the
// first version serialised on the 21 x 15 addresses of a frame).
__global__ void __launch_bounds__(HALO_WARPS * 32) halo_bwd_kernel(
    const float* __restrict__ pts, const float* __restrict__ bt_inv, const float* __restrict__ Tp,
    const float* __restrict__ DF, int64_t ld_df, const float* __restrict__ FB, int64_t ld_fb,
    const float* __restrict__ dn, int64_t n, int64_t ppf, int64_t pts_per_warp, float* __restrict__ d_pts,
    float* __restrict__ d_bt, float* __restrict__ d_T) {
    __shared__ float sdf_[HALO_WARPS][HALO_DIM + 2];
    __shared__ float sfb_[HALO_WARPS][HALO_DIM + 2];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t p_begin = ((int64_t)blockIdx.x * HALO_WARPS + warp) * pts_per_warp;
    const int64_t p_end = min(n, p_begin + pts_per_warp);
    float acc_bt[12], acc_T[3];
#pragma unroll
    for (int i = 0; i < 12; ++i) acc_bt[i] = 0.0f;
#pragma unroll
    for (int i = 0; i < 3; ++i) acc_T[i] = 0.0f;
    int64_t f_acc = -1;
    auto flush = [&]() {
        if (f_acc < 0 || lane >= HALO_J) return;
        if (d_bt) {
            float* db = d_bt + (f_acc * HALO_J + lane) * 16;
#pragma unroll
            for (int i = 0; i < 12; ++i) {
                if (acc_bt[i] != 0.0f) atomicAdd(&db[i], acc_bt[i]);
                acc_bt[i] = 0.0f;
            }
        }
        if (d_T) {
#pragma unroll
            for (int a = 0; a < 3; ++a) {
                if (acc_T[a] != 0.0f) atomicAdd(&d_T[(f_acc * HALO_J + lane) * 3 + a], acc_T[a]);
                acc_T[a] = 0.0f;
            }
        }
    };
    for (int64_t p = p_begin; p < p_end; ++p) {
        const int64_t f = p / ppf;
        if (f != f_acc) {
            flush();
            f_acc = f;
        }
        __syncwarp();
        for (int i = lane; i < HALO_DIM; i += 32) {
            sdf_[warp][i] = DF ? DF[p * ld_df + i] : 0.0f;
            sfb_[warp][i] = (dn && FB) ? FB[p * ld_fb + i] : 0.0f;
        }
        __syncwarp();
        float x[3], t[3] = {0.f, 0.f, 0.f}, dx[3] = {0.f, 0.f, 0.f};
        load_x(pts, p, x);
        if (dn) load_x(dn, p, t);
        if (lane < HALO_J) {
            const float* M = bt_inv + (f * HALO_J + lane) * 16;
            HaloBase b = halo_base(M, Tp + (f * HALO_J + lane) * 3, x, lane);
            if (!b.dead) {
                float gq[3], g2[3] = {0.f, 0.f, 0.f}, hv[3] = {0.f, 0.f, 0.f}, w[3] = {0.f, 0.f, 0.f}, dummy[3];
                halo_grad_hvp<false>(b, &sdf_[warp][lane * HALO_F], w, gq, dummy);
                if (dn) {
#pragma unroll
                    for (int a = 0; a < 3; ++a) w[a] = M[a * 4] * t[0] + M[a * 4 + 1] * t[1] + M[a * 4 + 2] * t[2];
                    halo_grad_hvp<true>(b, &sfb_[warp][lane * HALO_F], w, g2, hv);
                }
                float dq[3];
#pragma unroll
                for (int a = 0; a < 3; ++a) dq[a] = gq[a] + hv[a];
#pragma unroll
                for (int a = 0; a < 3; ++a) dx[a] = M[0 * 4 + a] * dq[0] + M[1 * 4 + a] * dq[1] + M[2 * 4 + a] * dq[2];
#pragma unroll
                for (int a = 0; a < 3; ++a) {
#pragma unroll
                    for (int c = 0; c < 3; ++c) acc_bt[a * 4 + c] += dq[a] * x[c] + g2[a] * t[c];
                    acc_bt[a * 4 + 3] += dq[a];
                    acc_T[a] -= dq[a];
                }
            }
        }
        if (d_pts) {
#pragma unroll
            for (int a = 0; a < 3; ++a) dx[a] = warp_sum(dx[a]);
            if (lane < 3) d_pts[p * 3 + lane] = lane == 0 ? dx[0] : (lane == 1 ? dx[1] : dx[2]);
        }
    }
    flush();
}

// ---- thread-per-point variants over TILED cotangents (HN_TC_MIXED16, chain16_hand.cu) -------------------------------------
// FB / DF arrive as column-major tiles [tile][1388 columns][128 points].  One block = one tile, one thread = one point; the 66
// columns of a joint are staged into shared memory sc[66][129] (conflict-free both ways), joints that are dead (h == 0) for
// every point of the tile are skipped without staging.
constexpr int HALO_TP = 128, HALO_TLD = 129;
constexpr int64_t HALO_TILE_FLOATS = (int64_t)HFB_LD * HALO_TP;

__device__ __forceinline__ void halo_stage_joint(const float* __restrict__ tile, int j, float* __restrict__ sc) {
    for (int idx = threadIdx.x; idx < HALO_F * HALO_TP; idx += HALO_TP) {
        const int i = idx >> 7, r = idx & 127;
        sc[i * HALO_TLD + r] = tile[(size_t)(j * HALO_F + i) * HALO_TP + r];
    }
}
// sc[i][r] += rows[(p0 + r) * ld + j * 66 + i]   (row-major addend, e.g. the cotangent the colour net sends to the feature)
__device__ __forceinline__ void halo_stage_add_rows(const float* __restrict__ rows, int64_t ld, int64_t p0, int64_t n, int j,
                                                    float* __restrict__ sc) {
    for (int idx = threadIdx.x; idx < HALO_F * HALO_TP; idx += HALO_TP) {
        const int r = idx / HALO_F, i = idx - r * HALO_F;
        if (p0 + r < n) sc[i * HALO_TLD + r] += rows[(p0 + r) * ld + j * HALO_F + i];
    }
}

__global__ void __launch_bounds__(HALO_TP) halo_normal_tiled_kernel(
    const float* __restrict__ pts, const float* __restrict__ bt_inv, const float* __restrict__ Tp,
    const float* __restrict__ FBt, int64_t n, int64_t ppf, float* __restrict__ normal) {
    // The column-major tile IS the thread-per-point layout (a warp's 32 rows of a column are one 128-byte line, every value is
    // read once): no staging through shared memory, no block-wide barrier.
    const int r = threadIdx.x;
    const int64_t p = (int64_t)blockIdx.x * HALO_TP + r;
    if (p >= n) return;
    const int64_t f = p / ppf;
    float x[3], acc[3] = {0.f, 0.f, 0.f};
    load_x(pts, p, x);
    const float* tile = FBt + (size_t)blockIdx.x * HALO_TILE_FLOATS + r;
    for (int j = 0; j < HALO_J; ++j) {
        const float* M = bt_inv + (f * HALO_J + j) * 16;
        HaloBase b = halo_base(M, Tp + (f * HALO_J + j) * 3, x, j);
        if (b.dead) continue;
        float g[3], dummy[3], w[3] = {0.f, 0.f, 0.f};
        halo_grad_hvp<false, HALO_TP>(b, tile + (size_t)j * HALO_F * HALO_TP, w, g, dummy);
#pragma unroll
        for (int a = 0; a < 3; ++a) acc[a] += M[0 * 4 + a] * g[0] + M[1 * 4 + a] * g[1] + M[2 * 4 + a] * g[2];
    }
    normal[p * 3] = acc[0]; normal[p * 3 + 1] = acc[1]; normal[p * 3 + 2] = acc[2];
}

// halo_bwd_kernel over tiled DF / FB; d_xyz (row-major, may be NULL) is added to DF while staging.  uniform_frame: every
// point of a tile belongs to one frame (pose gradients are reduced over the block before the atomics)
__global__ void __launch_bounds__(HALO_TP) halo_bwd_tiled_kernel(
    const float* __restrict__ pts, const float* __restrict__ bt_inv, const float* __restrict__ Tp,
    const float* __restrict__ DFt, const float* __restrict__ d_xyz, int64_t ld_dxyz, const float* __restrict__ FBt,
    const float* __restrict__ dn, int64_t n, int64_t ppf, int uniform_frame, float* __restrict__ d_pts,
    float* __restrict__ d_bt, float* __restrict__ d_T) {
    __shared__ float sc[HALO_F * HALO_TLD];
    __shared__ float red[HALO_TP / 32][16];
    const int r = threadIdx.x, warp = r >> 5, lane = r & 31;
    const int64_t p0 = (int64_t)blockIdx.x * HALO_TP, p = p0 + r;
    const bool live = p < n;
    const int64_t f = (live ? p : n - 1) / ppf;
    float x[3] = {0.f, 0.f, 0.f}, t[3] = {0.f, 0.f, 0.f}, dx[3] = {0.f, 0.f, 0.f};
    if (live) {
        load_x(pts, p, x);
        load_x(dn, p, t);
    }
    const float* dtile = DFt + (size_t)blockIdx.x * HALO_TILE_FLOATS;
    const float* ftile = FBt + (size_t)blockIdx.x * HALO_TILE_FLOATS;
    for (int j = 0; j < HALO_J; ++j) {
        const float* M = bt_inv + (f * HALO_J + j) * 16;
        HaloBase b = halo_base(M, Tp + (f * HALO_J + j) * 3, x, j);
        const bool on = live && !b.dead;
        if (!__syncthreads_or(on)) continue;
        // DF: straight from its column-major tile (= the thread-per-point layout) unless the colour net's row-major cotangent
        // has to be added, which goes through shared memory; FB: always straight from its tile
        float gq[3] = {0.f, 0.f, 0.f}, g2[3] = {0.f, 0.f, 0.f}, hv[3] = {0.f, 0.f, 0.f}, w[3] = {0.f, 0.f, 0.f}, dummy[3];
        if (d_xyz) {
            halo_stage_joint(dtile, j, sc);
            __syncthreads();
            halo_stage_add_rows(d_xyz, ld_dxyz, p0, n, j, sc);
            __syncthreads();
            if (on) halo_grad_hvp<false, HALO_TLD>(b, sc + r, w, gq, dummy);
        } else if (on) {
            halo_grad_hvp<false, HALO_TP>(b, dtile + (size_t)j * HALO_F * HALO_TP + r, w, gq, dummy);
        }
        float v[15];
#pragma unroll
        for (int i = 0; i < 15; ++i) v[i] = 0.0f;
        if (on) {
#pragma unroll
            for (int a = 0; a < 3; ++a) w[a] = M[a * 4] * t[0] + M[a * 4 + 1] * t[1] + M[a * 4 + 2] * t[2];
            halo_grad_hvp<true, HALO_TP>(b, ftile + (size_t)j * HALO_F * HALO_TP + r, w, g2, hv);
            float dq[3];
#pragma unroll
            for (int a = 0; a < 3; ++a) dq[a] = gq[a] + hv[a];
#pragma unroll
            for (int a = 0; a < 3; ++a) dx[a] += M[0 * 4 + a] * dq[0] + M[1 * 4 + a] * dq[1] + M[2 * 4 + a] * dq[2];
#pragma unroll
            for (int a = 0; a < 3; ++a) {
#pragma unroll
                for (int c = 0; c < 3; ++c) v[a * 4 + c] = dq[a] * x[c] + g2[a] * t[c];
                v[a * 4 + 3] = dq[a];
                v[12 + a] = -dq[a];
            }
        }
        if (d_bt || d_T) {
            if (uniform_frame) {
#pragma unroll
                for (int i = 0; i < 15; ++i) v[i] = warp_sum(v[i]);
                if (lane == 0) {
#pragma unroll
                    for (int i = 0; i < 15; ++i) red[warp][i] = v[i];
                }
                __syncthreads();
                if (r < 15) {
                    const float sum = red[0][r] + red[1][r] + red[2][r] + red[3][r];
                    if (sum != 0.0f) {
                        if (r < 12) { if (d_bt) atomicAdd(&d_bt[(f * HALO_J + j) * 16 + r], sum); }
                        else if (d_T) atomicAdd(&d_T[(f * HALO_J + j) * 3 + (r - 12)], sum);
                    }
                }
            } else if (on) {
#pragma unroll
                for (int i = 0; i < 12; ++i)
                    if (d_bt) atomicAdd(&d_bt[(f * HALO_J + j) * 16 + i], v[i]);
#pragma unroll
                for (int a = 0; a < 3; ++a)
                    if (d_T) atomicAdd(&d_T[(f * HALO_J + j) * 3 + a], v[12 + a]);
            }
        }
    }
    if (d_pts && live) {
        d_pts[p * 3] = dx[0]; d_pts[p * 3 + 1] = dx[1]; d_pts[p * 3 + 2] = dx[2];
    }
}

// D7[p, c] = s'(H7[p, c]) * w_out0[c]
__global__ void hand_normal_seed_kernel(const float* __restrict__ H7, const float* __restrict__ w_out0, int64_t n,
                                        float* __restrict__ D7) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n * 64) return;
    int c = (int)(i & 63) * 4;
    int64_t p = i >> 6;
    float4 h = ld4(H7 + p * 256 + c);
    float4 w = ld4(w_out0 + c);
    st4(D7 + p * 256 + c, make_float4(sprime_from_h(h.x) * w.x, sprime_from_h(h.y) * w.y, sprime_from_h(h.z) * w.z,
                                      sprime_from_h(h.w) * w.w));
}
__global__ void hand_sdf_head_kernel(const float* __restrict__ H7, const float* __restrict__ w_out0,
                                     const float* __restrict__ b_out, int64_t n, float* __restrict__ sdf) {
    int64_t p = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    int lane = threadIdx.x & 31;
    if (p >= n) return;
    float acc = 0.0f;
#pragma unroll
    for (int it = 0; it < 2; ++it) {
        int c = (it * 32 + lane) * 4;
        float4 h = ld4(H7 + p * 256 + c);
        float4 w = ld4(w_out0 + c);
        acc += h.x * w.x + h.y * w.y + h.z * w.z + h.w * w.w;
    }
    acc = warp_sum(acc);
    if (lane == 0) sdf[p] = acc + b_out[0];
}
__global__ void hand_assemble_dz8_kernel(const float* __restrict__ d_sdf, const float* __restrict__ d_feat,
                                         int64_t ld_dfeat, int64_t n, float* __restrict__ DZ8) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n * 260) return;
    int64_t p = i / 260;
    int j = (int)(i - p * 260);
    float v = 0.0f;
    if (j == 0) v = d_sdf ? d_sdf[p] : 0.0f;
    else if (j < 257) v = d_feat ? d_feat[p * ld_dfeat + (j - 1)] : 0.0f;
    DZ8[i] = v;
}
// dst[p, 0:cols] = src ? src[p, 0:cols] : 0   (row copies between different leading dimensions)
__global__ void copy_rows_kernel(const float* __restrict__ src, int64_t ld_src, int64_t n, int cols,
                                 float* __restrict__ dst, int64_t ld_dst) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n * cols) return;
    int64_t p = i / cols;
    int j = (int)(i - p * cols);
    dst[p * ld_dst + j] = src ? src[p * ld_src + j] : 0.0f;
}

struct HandSdfStash {
    float* HROW;   // [n,1644] = [h3 | feature | pad]
    float* H[8];   // [n,256] (H[3] aliases HROW, ld 1644)
    float* D[8];   // [n,256]
    float* FB;     // [n,1388]
    static constexpr int64_t kFloatsPerPoint = HROW_LD + 7 * 256 + 8 * 256 + HFB_LD;
    HandSdfStash(float* base, int64_t n) {
        float* p = base;
        HROW = p; p += n * HROW_LD;
        for (int l = 0; l < 8; ++l) {
            if (l == 3) { H[l] = HROW; continue; }
            H[l] = p; p += n * 256;
        }
        for (int l = 0; l < 8; ++l) { D[l] = p; p += n * 256; }
        FB = p;
    }
    int64_t ldH(int l) const { return l == 3 ? HROW_LD : 256; }
};

static int check_hand_sdf_mlp(const hn_mlp_t* m) {
    HN_REQUIRE(m && m->n_layers == 9, "hand SDF mlp must have 9 layers");
    static const int in_d[9] = {1386, 256, 256, 256, 1642, 256, 256, 256, 256};
    static const int out_d[9] = {256, 256, 256, 256, 256, 256, 256, 256, 257};
    for (int l = 0; l < 9; ++l) {
        HN_REQUIRE(m->in_dim[l] == in_d[l] && m->out_dim[l] == out_d[l], "hand SDF mlp layer %d is %dx%d, expected %dx%d",
                   l, m->out_dim[l], m->in_dim[l], out_d[l], in_d[l]);
        HN_REQUIRE(m->ld[l] >= round_up(in_d[l], 4) && m->ld[l] % 4 == 0, "bad ld for layer %d", l);
        HN_REQUIRE(m->W[l] && m->b[l] && aligned16(m->W[l]), "layer %d: null or misaligned weights", l);
    }
    return HN_OK;
}

static inline unsigned nblocks(int64_t work, int threads) { return (unsigned)ceil_div(work, threads); }

static void set_w(GemmArgs& g, const hn_mlp_t* m, int l, int col_off = 0) {
    // weights of layer l, optionally starting at input column col_off (a multiple of 4)
    g.B = m->W[l] + col_off; g.ldb = m->ld[l];
    g.BT = m->WT[l] ? m->WT[l] + (int64_t)col_off * m->ldT[l] : nullptr; g.ldbt = m->ldT[l];
    if (m->chain && col_off % 256 == 0) {      // pre-packed bf16 hi/lo operands (hn_mlp_bx3_pack), HN_TC_BF16X3 only
        const Bx3Layout L = bx3_layout(m);
        const uint8_t* base = reinterpret_cast<const uint8_t*>(m->chain);
        g.Bp = base + L.w[l]; g.bp_tile_bytes = bx3_tile_bytes(m->in_dim[l]); g.bp_kb0 = col_off / 64;
        g.btp_tile_bytes = bx3_tile_bytes(m->out_dim[l]);
        g.BTp = base + L.wt[l] + (col_off / 256) * g.btp_tile_bytes;
    }
}

// forward trunk: HROW feature -> H7.  H[] / ldH describe where each layer's activation goes.
static int hand_trunk_fwd(const hn_mlp_t* m, const float* pts, const float* bt_inv, const float* Tp, int64_t n,
                          int64_t ppf, float* HROW, float* const H[8], cudaStream_t s, int precision) {
    halo_feature_kernel<<<nblocks(n, HALO_WARPS), HALO_WARPS * 32, 0, s>>>(pts, bt_inv, Tp, n, ppf, HROW + HFEAT_OFF,
                                                                           HROW_LD);
    count_launch();
    HN_CHECK_LAUNCH();
    for (int l = 0; l < 8; ++l) {
        GemmArgs g;
        if (l == 0) { g.A = HROW + HFEAT_OFF; g.lda = HROW_LD; }
        else if (l == 4) { g.A = HROW; g.lda = HROW_LD; }
        else { g.A = H[l - 1]; g.lda = 256; }
        set_w(g, m, l);
        g.M = (int)n; g.N = 256; g.K = m->in_dim[l];
        g.C = H[l]; g.ldc = (l == 3) ? HROW_LD : 256;
        g.bias = m->b[l];
        HN_PROPAGATE((gemm_nt<EPI_BIAS_SOFTPLUS>(g, s, precision, ROLE_VALUE)));
    }
    return HN_OK;
}

// ---- HN_TC_MIXED16 (chain16_hand.cu): stash / workspace views and the operands appended to the bx3 pack -------------------
struct Hand16Stash {
    float *HROW, *FB, *RA, *RB;      // RA: H0, then D0;  RB: ZF4, then D4
    uint8_t *EM[8], *EML[8], *D16[8];
    uint8_t* F16;                    // fp16 pair tiles of the HALO feature (input of the value trunk)
    Hand16Stash(float* base, int64_t n) {
        const int64_t np = round_up(n, 128);
        float* p = base;
        HROW = p; p += np * HROW_LD;
        FB = p; p += np * HFB_LD;
        RA = p; p += np * 256;
        RB = p; p += np * 256;
        uint8_t* b = reinterpret_cast<uint8_t*>(p);
        for (int l = 0; l < 8; ++l) { EM[l] = b; b += np * 512; }
        for (int l = 0; l < 8; ++l) { EML[l] = b; b += np * 512; }
        for (int l = 0; l < 8; ++l) { D16[l] = b; b += np * 512; }
        F16 = b;
    }
};
static const uint8_t* hand16_ops(const hn_mlp_t* m) {
    const int64_t off = round_up(bx3_layout(m).total, 1024);
    if (!m->chain || m->chain_bytes < off + (int64_t)chain::hand_layout().total) return nullptr;
    return reinterpret_cast<const uint8_t*>(m->chain) + off;
}
static bool use_hand16(const hn_mlp_t* m, int precision) { return precision == HN_TC_MIXED16 && hand16_ops(m) != nullptr; }

// feature-side contractions into the chain: H0 = softplus(F W_0^T + b_0) and ZF4 = F W_4[:, 256:]^T (or, on the tangent
// features, Q0 and QF4 without bias / activation)
static int hand16_feature_in(const hn_mlp_t* m, const float* F, int64_t n, float* out0, float* out4, bool activate,
                             cudaStream_t s) {
    GemmArgs g;
    g.A = F; g.lda = HROW_LD;
    set_w(g, m, 0);
    g.M = (int)n; g.N = 256; g.K = HALO_DIM;
    g.C = out0; g.ldc = 256;
    if (activate) {
        g.bias = m->b[0];
        HN_PROPAGATE((gemm_nt<EPI_BIAS_SOFTPLUS>(g, s, HN_TC_BF16X3, ROLE_VALUE)));
    } else {
        HN_PROPAGATE((gemm_nt<EPI_STORE>(g, s, HN_TC_BF16X3, ROLE_VALUE)));
    }
    GemmArgs f;
    f.A = F; f.lda = HROW_LD;
    set_w(f, m, 4, HFEAT_OFF);
    f.M = (int)n; f.N = 256; f.K = HALO_DIM;
    f.C = out4; f.ldc = 256;
    return gemm_nt<EPI_STORE>(f, s, HN_TC_BF16X3, ROLE_VALUE);
}
}  // namespace hn

using namespace hn;

extern "C" {

int64_t hn_sdf_hand_stash_floats(int64_t n) { return std::max(n * HandSdfStash::kFloatsPerPoint, chain::hand16_stash_floats(n)); }

// The hand SDF net's packed operands: the per-layer bf16 hi/lo tiles of hn_mlp_bx3_pack, followed (1 KB aligned) by the
// chain operands of its 256 x 256 layers (HN_TC_MIXED16, chain16_hand.cu)
int64_t hn_sdf_hand_chain_bytes(const hn_mlp_t* m) {
    if (!m || m->n_layers != 9) return 0;
    return round_up(bx3_layout(m).total, 1024) + (int64_t)chain::hand_layout().total;
}
int hn_sdf_hand_chain_pack(const hn_mlp_t* m, void* buf, int64_t bytes, hn_stream_t stream) {
    HN_PROPAGATE(check_hand_sdf_mlp(m));
    HN_REQUIRE(buf && bytes >= hn_sdf_hand_chain_bytes(m) && aligned16(buf), "hn_sdf_hand_chain_pack: buffer too small or misaligned");
    HN_PROPAGATE(hn_mlp_bx3_pack(m, buf, bytes, stream));
    for (int l = 1; l < 9; ++l) HN_REQUIRE(m->WT[l], "hn_sdf_hand_chain_pack: layer %d has no transposed copy", l);
    return chain::hand16_pack(m, reinterpret_cast<uint8_t*>(buf) + round_up(bx3_layout(m).total, 1024), (cudaStream_t)stream);
}

int64_t hn_sdf_hand_ws_floats(int64_t n, int kind) {
    switch (kind) {
        case HN_WS_SDF_ONLY: return round_up(n, 128) * (HROW_LD + 2 * 256);
        case HN_WS_FWD: return 4;
        // AU4 [n,1644], U ping-pong 2x256, DZ8 260, DZ ping-pong 2x256, DF [n,1388]
        case HN_WS_BWD: return std::max(n * (HROW_LD + 2 * 256 + 260 + 2 * 256 + HFB_LD), chain::hand16_bwd_ws_floats(n));
        default: return -1;
    }
}

int hn_sdf_hand_sdf(const hn_mlp_t* mlp, const float* pts, const float* bt_inv, const float* T_pose, int64_t n,
                    int64_t pts_per_frame, float* sdf, float* ws, int64_t ws_floats, int precision,
                    hn_stream_t stream) {
    HN_PROPAGATE(check_hand_sdf_mlp(mlp));
    const bool m16 = use_hand16(mlp, precision);
    precision = base_precision(precision);
    HN_REQUIRE(precision_supported(precision), "hn_sdf_hand_sdf: precision %d not supported", precision);
    HN_REQUIRE(n >= 0 && n < (1ll << 31) && pts_per_frame > 0, "bad sizes");
    if (n == 0) return HN_OK;
    HN_REQUIRE(pts && bt_inv && T_pose && sdf && ws && ws_floats >= hn_sdf_hand_ws_floats(n, HN_WS_SDF_ONLY) &&
                   aligned16(ws), "hn_sdf_hand_sdf: null pointer or workspace too small");
    cudaStream_t s = (cudaStream_t)stream;
    if (m16) {
        // the feature goes straight into the trunk's fp16 pair tiles (no fp32 rows, no per-layer contraction)
        const int64_t np = round_up(n, 128);
        uint8_t* F16 = reinterpret_cast<uint8_t*>(ws);
        halo_feature16_kernel<<<nblocks(np, HALO_WARPS), HALO_WARPS * 32, 0, s>>>(pts, bt_inv, T_pose, n, pts_per_frame, nullptr, 0, F16);
        count_launch();
        HN_CHECK_LAUNCH();
        return chain::launch_hand16_trunk(mlp, hand16_ops(mlp), n, F16, nullptr, nullptr, sdf, nullptr, 0, nullptr, nullptr, s);
    }
    float* HROW = ws;
    float* P0 = HROW + n * HROW_LD;
    float* P1 = P0 + n * 256;
    float* H[8] = {P0, P1, P0, HROW, P0, P1, P0, P1};
    HN_PROPAGATE(hand_trunk_fwd(mlp, pts, bt_inv, T_pose, n, pts_per_frame, HROW, H, s, precision));
    hand_sdf_head_kernel<<<nblocks(n * 32, 256), 256, 0, s>>>(H[7], mlp->W[8], mlp->b[8], n, sdf);
    count_launch();
    HN_CHECK_LAUNCH();
    return HN_OK;
}

static thread_local bool t_hand_render_only = false;

int hn_sdf_hand_fwd_render(const hn_mlp_t* mlp, const float* pts, const float* bt_inv, const float* T_pose, int64_t n,
                           int64_t pts_per_frame, float* sdf, float* feat, int64_t ld_feat, float* normal,
                           float* xyz_feature, int64_t ld_xyz, float* stash, int64_t stash_floats, int precision,
                           hn_stream_t stream) {
    t_hand_render_only = true;
    const int r = hn_sdf_hand_fwd(mlp, pts, bt_inv, T_pose, n, pts_per_frame, sdf, feat, ld_feat, normal, xyz_feature, ld_xyz, stash,
                                  stash_floats, precision, stream);
    t_hand_render_only = false;
    return r;
}

int hn_sdf_hand_fwd(const hn_mlp_t* mlp, const float* pts, const float* bt_inv, const float* T_pose, int64_t n,
                    int64_t pts_per_frame, float* sdf, float* feat, int64_t ld_feat, float* normal,
                    float* xyz_feature, int64_t ld_xyz, float* stash, int64_t stash_floats, int precision,
                    hn_stream_t stream) {
    HN_PROPAGATE(check_hand_sdf_mlp(mlp));
    const bool m16 = use_hand16(mlp, precision);
    precision = base_precision(precision);
    HN_REQUIRE(precision_supported(precision), "hn_sdf_hand_fwd: precision %d not supported", precision);
    HN_REQUIRE(n >= 0 && n < (1ll << 31) && pts_per_frame > 0, "bad sizes");
    if (n == 0) return HN_OK;
    HN_REQUIRE(pts && bt_inv && T_pose && sdf && feat && normal && stash, "hn_sdf_hand_fwd: null pointer");
    HN_REQUIRE(stash_floats >= hn_sdf_hand_stash_floats(n) && aligned16(stash), "stash too small or misaligned");
    HN_REQUIRE(ld_feat >= 256 && ld_feat % 4 == 0 && aligned16(feat), "feat must be 16B aligned with ld%%4==0");
    cudaStream_t s = (cudaStream_t)stream;
    if (m16) {
        // HN_TC_MIXED16: the 256 x 256 layers as tile-chain kernels (chain16_hand.cu), the 1386-wide ends per layer
        Hand16Stash h(stash, n);
        const uint8_t* ops = hand16_ops(mlp);
        // HALO feature once: fp32 rows (the colour net's xyz_feature, a view of the stash) + the trunk's fp16 pair tiles
        halo_feature16_kernel<<<nblocks(round_up(n, 128), HALO_WARPS), HALO_WARPS * 32, 0, s>>>(pts, bt_inv, T_pose, n, pts_per_frame,
                                                                                                h.HROW + HFEAT_OFF, HROW_LD, h.F16);
        count_launch();
        HN_CHECK_LAUNCH();
        HN_PROPAGATE(chain::launch_hand16_trunk(mlp, ops, n, h.F16, nullptr, nullptr, sdf, feat, ld_feat, h.EM, h.EML, s));
        if (xyz_feature) {
            copy_rows_kernel<<<nblocks(n * HALO_DIM, 256), 256, 0, s>>>(h.HROW + HFEAT_OFF, HROW_LD, n, HALO_DIM, xyz_feature,
                                                                        ld_xyz);
            count_launch();
            HN_CHECK_LAUNCH();
        }
        HN_PROPAGATE(chain::launch_hand16_nsweep(mlp, ops, n, h.EM, h.EML, t_hand_render_only ? nullptr : h.D16, h.FB, s));
        halo_normal_tiled_kernel<<<nblocks(n, HALO_TP), HALO_TP, 0, s>>>(pts, bt_inv, T_pose, h.FB, n, pts_per_frame, normal);
        count_launch();
        HN_CHECK_LAUNCH();
        return HN_OK;
    }
    HandSdfStash st(stash, n);
    HN_PROPAGATE(hand_trunk_fwd(mlp, pts, bt_inv, T_pose, n, pts_per_frame, st.HROW, st.H, s, precision));
    hand_sdf_head_kernel<<<nblocks(n * 32, 256), 256, 0, s>>>(st.H[7], mlp->W[8], mlp->b[8], n, sdf);
    count_launch();
    HN_CHECK_LAUNCH();
    {
        GemmArgs g;
        g.A = st.H[7]; g.lda = 256;
        g.B = mlp->W[8] + mlp->ld[8]; g.ldb = mlp->ld[8];
        g.M = (int)n; g.N = 256; g.K = 256;
        g.C = feat; g.ldc = ld_feat; g.bias = mlp->b[8] + 1;
        HN_PROPAGATE((gemm_nt<EPI_STORE>(g, s, precision, ROLE_VALUE)));
    }
    if (xyz_feature) {
        copy_rows_kernel<<<nblocks(n * HALO_DIM, 256), 256, 0, s>>>(st.HROW + HFEAT_OFF, HROW_LD, n, HALO_DIM, xyz_feature,
                                                                    ld_xyz);
        count_launch();
        HN_CHECK_LAUNCH();
    }
    // normal sweep
    hand_normal_seed_kernel<<<nblocks(n * 64, 256), 256, 0, s>>>(st.H[7], mlp->W[8], n, st.D[7]);
    count_launch();
    HN_CHECK_LAUNCH();
    for (int l = 7; l >= 1; --l) {
        GemmArgs g;
        g.A = st.D[l]; g.lda = 256;
        set_w(g, mlp, l);
        g.M = (int)n; g.N = 256; g.K = 256;
        g.C = st.D[l - 1]; g.ldc = 256;
        g.aux1 = st.H[l - 1]; g.ldaux1 = st.ldH(l - 1);
        HN_PROPAGATE((gemm_nn<EPI_MUL_SPRIME>(g, s, precision, ROLE_VALUE)));
        if (l == 4) {      // the feature part of the skip input: FB = D4 @ W4[:, 256:]
            GemmArgs f;
            f.A = st.D[4]; f.lda = 256;
            set_w(f, mlp, 4, HFEAT_OFF);
            f.M = (int)n; f.N = HALO_DIM; f.K = 256;
            f.C = st.FB; f.ldc = HFB_LD;
            HN_PROPAGATE((gemm_nn<EPI_STORE>(f, s, precision, ROLE_VALUE)));
        }
    }
    {
        GemmArgs g;
        g.A = st.D[0]; g.lda = 256;
        set_w(g, mlp, 0);
        g.M = (int)n; g.N = HALO_DIM; g.K = 256;
        g.C = st.FB; g.ldc = HFB_LD; g.aux1 = st.FB; g.ldaux1 = HFB_LD;
        HN_PROPAGATE((gemm_nn<EPI_ADD_AUX>(g, s, precision, ROLE_VALUE)));
    }
    halo_normal_kernel<<<nblocks(n, HALO_WARPS), HALO_WARPS * 32, 0, s>>>(pts, bt_inv, T_pose, st.FB, HFB_LD, n,
                                                                          pts_per_frame, normal);
    count_launch();
    HN_CHECK_LAUNCH();
    return HN_OK;
}

int hn_sdf_hand_bwd(const hn_mlp_t* mlp, const float* pts, const float* bt_inv, const float* T_pose, int64_t n,
                    int64_t pts_per_frame, float* stash, const float* d_sdf, const float* d_feat, int64_t ld_dfeat,
                    const float* d_normal, const float* d_xyz_feature, int64_t ld_dxyz, float* d_pts, float* d_bt_inv,
                    float* d_T_pose, const hn_mlp_grad_t* grad, float* ws, int64_t ws_floats, int precision,
                    hn_stream_t stream) {
    HN_PROPAGATE(check_hand_sdf_mlp(mlp));
    const bool m16 = use_hand16(mlp, precision);
    precision = base_precision(precision);
    HN_REQUIRE(precision_supported(precision), "hn_sdf_hand_bwd: precision %d not supported", precision);
    HN_REQUIRE(n >= 0 && n < (1ll << 31) && pts_per_frame > 0, "bad sizes");
    if (n == 0) return HN_OK;
    HN_REQUIRE(pts && bt_inv && T_pose && stash && d_normal && ws, "hn_sdf_hand_bwd: null pointer");
    HN_REQUIRE(ws_floats >= hn_sdf_hand_ws_floats(n, HN_WS_BWD) && aligned16(ws), "workspace too small or misaligned");
    cudaStream_t s = (cudaStream_t)stream;
    if (m16) {
        HN_REQUIRE(!grad, "hn_sdf_hand_bwd: HN_TC_MIXED16 computes no weight gradients for the hand net (the forward of a call "
                          "that needs them must be made with HN_TC_BF16X3)");
        Hand16Stash h(stash, n);
        const uint8_t* ops = hand16_ops(mlp);
        const int64_t np = round_up(n, 128);
        float* AU4 = ws;                                   // [np, 1644]: tangent feature rows at + 256
        float* DF = AU4 + np * HROW_LD;                    // [np, 1388]
        float* Q0 = DF + np * HFB_LD;                      // [np, 256]: tF W_0^T
        float* QF4 = Q0 + np * 256;                        // [np, 256]: tF W_4[:, 256:]^T
        uint8_t* X16[8];
        uint8_t* b = reinterpret_cast<uint8_t*>(QF4 + np * 256);
        for (int l = 0; l < 8; ++l) { X16[l] = b; b += np * 512; }
        halo_tangent_kernel<<<nblocks(n, HALO_WARPS), HALO_WARPS * 32, 0, s>>>(pts, bt_inv, T_pose, d_normal, n, pts_per_frame,
                                                                               AU4 + HFEAT_OFF, HROW_LD);
        count_launch();
        HN_CHECK_LAUNCH();
        HN_PROPAGATE(hand16_feature_in(mlp, AU4 + HFEAT_OFF, n, Q0, QF4, false, s));
        HN_PROPAGATE(chain::launch_hand16_bwd(mlp, ops, n, h.EM, h.D16, X16, Q0, QF4, d_sdf, d_feat, ld_dfeat, DF, s));
        if (d_pts || d_bt_inv || d_T_pose) {
            const int uniform = (pts_per_frame % HALO_TP == 0 || n <= pts_per_frame) ? 1 : 0;
            halo_bwd_tiled_kernel<<<nblocks(n, HALO_TP), HALO_TP, 0, s>>>(pts, bt_inv, T_pose, DF, d_xyz_feature, ld_dxyz, h.FB,
                                                                          d_normal, n, pts_per_frame, uniform, d_pts, d_bt_inv,
                                                                          d_T_pose);
            count_launch();
            HN_CHECK_LAUNCH();
        }
        return HN_OK;
    }
    HandSdfStash st(stash, n);
    float* AU4 = ws;                                  // [n,1644] = [u3 | tangent feature | pad]
    float* U[2] = {AU4 + n * HROW_LD, AU4 + n * HROW_LD + n * 256};
    float* DZ8 = U[1] + n * 256;
    float* DZ[2] = {DZ8 + n * 260, DZ8 + n * 260 + n * 256};
    float* DF = DZ[1] + n * 256;                      // [n,1388]
    const int splits_target = 2 * sm_count();
    const bool need_input_grad = d_pts || d_bt_inv || d_T_pose;

    auto dw_gemm = [&](const float* P, int64_t ldp, int out, const float* Q, int64_t ldq, int in, int l) -> int {
        if (!grad || !grad->dW[l]) return HN_OK;
        GemmArgs g;
        g.A = P; g.lda = ldp; g.B = Q; g.ldb = ldq;
        g.M = out; g.N = in; g.K = (int)n;
        g.C = grad->dW[l]; g.ldc = mlp->ld[l];
        int tiles = (int)(ceil_div(out, 128) * ceil_div(in, 256));
        int splits = (int)max((int64_t)1, min((int64_t)ceil_div(splits_target, tiles), ceil_div(n, 256)));
        return gemm_tn(g, s, precision, splits);
    };
    auto db_sum = [&](const float* X, int64_t ldx, int cols, int l) -> int {
        if (!grad || !grad->db[l]) return HN_OK;
        return launch_colsum(X, ldx, n, cols, 1.0f, grad->db[l], s);
    };
    auto in_ptr = [&](float* skiprow, float* const prev[8], int l, const float** A, int64_t* lda) {
        if (l == 0) { *A = skiprow + HFEAT_OFF; *lda = HROW_LD; }
        else if (l == 4) { *A = skiprow; *lda = HROW_LD; }
        else { *A = prev[l - 1]; *lda = 256; }
    };

    // ---- tangent sweep ---------------------------------------------------------------------------
    halo_tangent_kernel<<<nblocks(n, HALO_WARPS), HALO_WARPS * 32, 0, s>>>(pts, bt_inv, T_pose, d_normal, n,
                                                                           pts_per_frame, AU4 + HFEAT_OFF, HROW_LD);
    count_launch();
    HN_CHECK_LAUNCH();
    float* Uout[8] = {U[0], U[1], U[0], AU4, U[0], U[1], U[0], U[1]};
    for (int l = 0; l < 8; ++l) {
        const float* A; int64_t lda;
        in_ptr(AU4, Uout, l, &A, &lda);
        HN_PROPAGATE(dw_gemm(st.D[l], 256, 256, A, lda, mlp->in_dim[l], l));
        GemmArgs g;
        g.A = A; g.lda = lda;
        set_w(g, mlp, l);
        g.M = (int)n; g.N = 256; g.K = mlp->in_dim[l];
        g.C = Uout[l]; g.ldc = (l == 3) ? HROW_LD : 256;
        g.aux1 = st.H[l]; g.ldaux1 = st.ldH(l);
        g.C2 = st.D[l]; g.ldc2 = 256;
        HN_PROPAGATE((gemm_nt<EPI_TANGENT>(g, s, precision)));
    }
    if (grad && grad->dW[8]) HN_PROPAGATE(launch_colsum(Uout[7], 256, n, 256, 1.0f, grad->dW[8], s));

    // ---- reverse sweep ---------------------------------------------------------------------------
    hand_assemble_dz8_kernel<<<nblocks(n * 260, 256), 256, 0, s>>>(d_sdf, d_feat, ld_dfeat, n, DZ8);
    count_launch();
    HN_CHECK_LAUNCH();
    // DF starts as the cotangent that reaches the feature from outside (the colour net): it enters as the aux operand of
    // the first contraction that writes DF (the skip part of layer 4), no copy
    const float* dz = DZ8;
    int64_t ld_dz = 260;
    for (int l = 8; l >= 1; --l) {
        const float* A; int64_t lda;
        in_ptr(st.HROW, st.H, l, &A, &lda);
        HN_PROPAGATE(dw_gemm(dz, ld_dz, mlp->out_dim[l], A, lda, mlp->in_dim[l], l));
        HN_PROPAGATE(db_sum(dz, ld_dz, mlp->out_dim[l], l));
        GemmArgs g;
        g.A = dz; g.lda = ld_dz;
        set_w(g, mlp, l);
        g.M = (int)n; g.N = 256; g.K = mlp->out_dim[l];
        float* out = DZ[l & 1];
        g.C = out; g.ldc = 256;
        g.aux1 = st.H[l - 1]; g.ldaux1 = st.ldH(l - 1);
        g.aux2 = st.D[l - 1]; g.ldaux2 = 256;
        HN_PROPAGATE((gemm_nn<EPI_REVERSE>(g, s, precision)));
        if (l == 4 && need_input_grad) {      // skip part: DF += DZ4 @ W4[:, 256:]
            GemmArgs f;
            f.A = dz; f.lda = ld_dz;
            set_w(f, mlp, 4, HFEAT_OFF);
            f.M = (int)n; f.N = HALO_DIM; f.K = 256;
            f.C = DF; f.ldc = HFB_LD;
            if (d_xyz_feature) {
                f.aux1 = d_xyz_feature; f.ldaux1 = ld_dxyz;
                HN_PROPAGATE((gemm_nn<EPI_ADD_AUX>(f, s, precision)));
            } else {
                HN_PROPAGATE((gemm_nn<EPI_STORE>(f, s, precision)));
            }
        }
        dz = out; ld_dz = 256;
    }
    HN_PROPAGATE(dw_gemm(dz, 256, 256, st.HROW + HFEAT_OFF, HROW_LD, HALO_DIM, 0));
    HN_PROPAGATE(db_sum(dz, 256, 256, 0));
    if (need_input_grad) {
        GemmArgs g;
        g.A = dz; g.lda = 256;
        set_w(g, mlp, 0);
        g.M = (int)n; g.N = HALO_DIM; g.K = 256;
        g.C = DF; g.ldc = HFB_LD; g.aux1 = DF; g.ldaux1 = HFB_LD;
        HN_PROPAGATE((gemm_nn<EPI_ADD_AUX>(g, s, precision)));
        // ~16 blocks per SM in flight; every warp owns a contiguous run of points
        const int64_t per_warp = max((int64_t)1, ceil_div(n, (int64_t)sm_count() * 16 * HALO_WARPS));
        halo_bwd_kernel<<<nblocks(ceil_div(n, per_warp), HALO_WARPS), HALO_WARPS * 32, 0, s>>>(
            pts, bt_inv, T_pose, DF, HFB_LD, st.FB, HFB_LD, d_normal, n, pts_per_frame, per_warp, d_pts, d_bt_inv, d_T_pose);
        count_launch();
        HN_CHECK_LAUNCH();
    }
    return HN_OK;
}

}  // extern "C"

// ==========================================================================================
// Hand colour field: input [xyz_feature 1386 | pad 2 | feature 256 | enc4(normal) 27 | pad 1] = 1672;
// the mlp's layer 0 must be packed with a 2-column gap after input column 1386 (in_dim 1671).
// ==========================================================================================
namespace hn {

constexpr int HCIN_LD = 1672, HCIN_OFF_FEAT = 1388, HCIN_OFF_NRM = 1644, HCIN_DIM = 1671;

// one thread per float4 of a [n, 1672] input row; the feature and encoding blocks start at multiples of 4 (1388, 1644) and the
// xyz rows are 16-byte aligned whenever vec_xyz (a view of the SDF stash, ld 1644)
__global__ void color_hand_input_kernel(const float* __restrict__ xyz, int64_t ld_xyz, int vec_xyz, const float* __restrict__ feat,
                                        int64_t ld_feat, const float* __restrict__ normal, int64_t n,
                                        float* __restrict__ CIN) {
    constexpr int Q = HCIN_LD / 4;      // 418 float4 per row
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n * Q) return;
    const int64_t p = i / Q;
    const int q = (int)(i - p * Q), j = q * 4;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (j + 3 < HALO_DIM) {
        const float* s = xyz + p * ld_xyz + j;
        v = vec_xyz ? *reinterpret_cast<const float4*>(s) : make_float4(s[0], s[1], s[2], s[3]);
    } else if (j < HALO_DIM) {          // 1384..1387: two feature values, then the 2-column gap
        v.x = xyz[p * ld_xyz + j];
        v.y = xyz[p * ld_xyz + j + 1];
    } else if (j >= HCIN_OFF_FEAT && j < HCIN_OFF_NRM) {
        v = *reinterpret_cast<const float4*>(feat + p * ld_feat + (j - HCIN_OFF_FEAT));
    } else if (j >= HCIN_OFF_NRM) {
        const float x[3] = {normal[p * 3], normal[p * 3 + 1], normal[p * 3 + 2]};
        const int c = j - HCIN_OFF_NRM;
        v.x = enc3_col(x, 4, c);
        v.y = enc3_col(x, 4, c + 1);
        v.z = enc3_col(x, 4, c + 2);
        v.w = c + 3 < HCIN_DIM - HCIN_OFF_NRM ? enc3_col(x, 4, c + 3) : 0.0f;
    }
    *reinterpret_cast<float4*>(CIN + p * HCIN_LD + j) = v;
}

// Render path: the first layer as TWO contractions -- over the xyz_feature rows where they lie (no copy into an input row) and
// over this small second operand A2 [n, 328] = [44 zeros | feature 256 | enc4(normal) 27 | 0], whose column c is input column
// 1344 + c of the packed first-layer weights (k-block 21 onwards; the 44 leading columns belong to the xyz part)
constexpr int HCIN2_LD = 328, HCIN2_KB0 = 21, HCIN2_OFF = HCIN2_KB0 * 64;     // 1344
__global__ void color_hand_input2_kernel(const float* __restrict__ feat, int64_t ld_feat, const float* __restrict__ normal, int64_t n,
                                         float* __restrict__ A2) {
    constexpr int Q = HCIN2_LD / 4;     // 82 float4 per row
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n * Q) return;
    const int64_t p = i / Q;
    const int j = (int)(i - p * Q) * 4 + HCIN2_OFF;      // column of the full 1672-wide row
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (j >= HCIN_OFF_FEAT && j < HCIN_OFF_NRM) {
        v = *reinterpret_cast<const float4*>(feat + p * ld_feat + (j - HCIN_OFF_FEAT));
    } else if (j >= HCIN_OFF_NRM) {
        const float x[3] = {normal[p * 3], normal[p * 3 + 1], normal[p * 3 + 2]};
        const int c = j - HCIN_OFF_NRM;
        v.x = enc3_col(x, 4, c);
        v.y = enc3_col(x, 4, c + 1);
        v.z = enc3_col(x, 4, c + 2);
        v.w = c + 3 < HCIN_DIM - HCIN_OFF_NRM ? enc3_col(x, 4, c + 3) : 0.0f;
    }
    *reinterpret_cast<float4*>(A2 + p * HCIN2_LD + (j - HCIN2_OFF)) = v;
}

// scatter of the first layer's input cotangent DCIN [n, 1672]: one thread per float4 of the xyz / feature blocks (vec: both
// destinations 16-byte aligned with ld % 4 == 0), three more threads per point for the normal through J_enc^T
__global__ void color_hand_input_bwd_kernel(const float* __restrict__ CIN, const float* __restrict__ DCIN, int64_t n,
                                            float* __restrict__ d_xyz, int64_t ld_dxyz, float* __restrict__ d_feat,
                                            int64_t ld_dfeat, float* __restrict__ d_normal, int vec) {
    constexpr int QX = (HALO_DIM + 3) / 4, QF = 64, per = QX + QF + 3;      // 347 + 64 + 3
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n * per) return;
    const int64_t p = i / per;
    const int q = (int)(i - p * per);
    const float* g = DCIN + p * HCIN_LD;
    if (q < QX) {
        if (!d_xyz) return;
        const int j = q * 4;
        const float4 v = *reinterpret_cast<const float4*>(g + j);
        float* o = d_xyz + p * ld_dxyz + j;
        if (vec && j + 3 < HALO_DIM) {
            *reinterpret_cast<float4*>(o) = v;
        } else {
            const float t[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
            for (int k = 0; k < 4; ++k)
                if (j + k < HALO_DIM) o[k] = t[k];
        }
    } else if (q < QX + QF) {
        if (!d_feat) return;
        const int j = (q - QX) * 4;
        const float4 v = *reinterpret_cast<const float4*>(g + HCIN_OFF_FEAT + j);
        float* o = d_feat + p * ld_dfeat + j;
        if (vec) *reinterpret_cast<float4*>(o) = v;
        else { o[0] = v.x; o[1] = v.y; o[2] = v.z; o[3] = v.w; }
    } else if (d_normal) {
        const int c = q - QX - QF;
        d_normal[p * 3 + c] = enc3_jt_from_enc(CIN + p * HCIN_LD + HCIN_OFF_NRM, g + HCIN_OFF_NRM, 4, c);
    }
}

__global__ void sigmoid_bwd4_kernel(const float* __restrict__ rgb, const float* __restrict__ d_rgb, int64_t n,
                                    float* __restrict__ DZ) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n * 4) return;
    int64_t p = i >> 2;
    int j = (int)(i & 3);
    float v = 0.0f;
    if (j < 3) {
        float y = rgb[p * 3 + j];
        v = d_rgb[p * 3 + j] * y * (1.0f - y);
    }
    DZ[i] = v;
}

static int check_color_hand_mlp(const hn_mlp_t* m) {
    HN_REQUIRE(m && m->n_layers == 5, "hand colour mlp must have 5 layers");
    static const int in_d[5] = {HCIN_DIM, 256, 256, 256, 256};
    static const int out_d[5] = {256, 256, 256, 256, 3};
    for (int l = 0; l < 5; ++l) {
        HN_REQUIRE(m->in_dim[l] == in_d[l] && m->out_dim[l] == out_d[l],
                   "hand colour mlp layer %d has wrong shape (layer 0 must be packed with the 2-column gap: in_dim 1671)", l);
        HN_REQUIRE(m->ld[l] >= round_up(in_d[l], 4) && m->ld[l] % 4 == 0, "bad ld for layer %d", l);
        HN_REQUIRE(m->W[l] && m->b[l] && aligned16(m->W[l]), "layer %d: null or misaligned weights", l);
    }
    return HN_OK;
}

}  // namespace hn

extern "C" {

int64_t hn_color_hand_stash_floats(int64_t n) { return n * (HCIN_LD + 4 * 256); }
int64_t hn_color_hand_ws_floats(int64_t n, int kind) {
    if (kind == HN_WS_BWD) return n * (2 * 256 + 4 + HCIN_LD);
    return 4;
}

static thread_local bool t_color_hand_render_only = false;
static const uint8_t* color_hand_tail_ops(const hn_mlp_t* m) {
    const int64_t off = round_up(bx3_layout(m).total, 1024);
    if (!m->chain || m->chain_bytes < off + chain::color_tail_bytes()) return nullptr;
    return reinterpret_cast<const uint8_t*>(m->chain) + off;
}

// The hand colour net's packed operands: the per-layer tiles of hn_mlp_bx3_pack followed (1 KB aligned) by the chain operands of
// its layers 1..4 (forward-only rendering under HN_TC_MIXED16: hn_color_hand_fwd_render)
int64_t hn_color_hand_chain_bytes(const hn_mlp_t* m) {
    if (!m || m->n_layers != 5) return 0;
    return round_up(bx3_layout(m).total, 1024) + chain::color_tail_bytes();
}
int hn_color_hand_chain_pack(const hn_mlp_t* m, void* buf, int64_t bytes, hn_stream_t stream) {
    HN_PROPAGATE(check_color_hand_mlp(m));
    HN_REQUIRE(buf && bytes >= hn_color_hand_chain_bytes(m) && aligned16(buf), "hn_color_hand_chain_pack: buffer too small or misaligned");
    HN_PROPAGATE(hn_mlp_bx3_pack(m, buf, bytes, stream));
    return chain::color_tail_pack(m, reinterpret_cast<uint8_t*>(buf) + round_up(bx3_layout(m).total, 1024), (cudaStream_t)stream);
}

int hn_color_hand_fwd_render(const hn_mlp_t* mlp, const float* xyz_feature, int64_t ld_xyz, const float* feat,
                             int64_t ld_feat, const float* normal, int64_t n, float* rgb, float* stash,
                             int64_t stash_floats, int precision, hn_stream_t stream) {
    t_color_hand_render_only = true;
    const int r = hn_color_hand_fwd(mlp, xyz_feature, ld_xyz, feat, ld_feat, normal, n, rgb, stash, stash_floats, precision, stream);
    t_color_hand_render_only = false;
    return r;
}

int hn_color_hand_fwd(const hn_mlp_t* mlp, const float* xyz_feature, int64_t ld_xyz, const float* feat,
                      int64_t ld_feat, const float* normal, int64_t n, float* rgb, float* stash,
                      int64_t stash_floats, int precision, hn_stream_t stream) {
    HN_PROPAGATE(check_color_hand_mlp(mlp));
    const bool tail = precision == HN_TC_MIXED16 && t_color_hand_render_only && color_hand_tail_ops(mlp) != nullptr;
    precision = base_precision(precision);
    HN_REQUIRE(precision_supported(precision), "hn_color_hand_fwd: precision %d not supported", precision);
    HN_REQUIRE(n >= 0 && n < (1ll << 31), "n_pts out of range");
    if (n == 0) return HN_OK;
    HN_REQUIRE(xyz_feature && feat && normal && rgb && stash, "hn_color_hand_fwd: null pointer");
    HN_REQUIRE(stash_floats >= hn_color_hand_stash_floats(n) && aligned16(stash), "stash too small or misaligned");
    cudaStream_t s = (cudaStream_t)stream;
    float* CIN = stash;
    float* R[4];
    for (int l = 0; l < 4; ++l) R[l] = stash + n * HCIN_LD + (int64_t)l * n * 256;
    HN_REQUIRE(ld_feat % 4 == 0 && aligned16(feat), "hn_color_hand_fwd: feat must be 16-byte aligned with ld %% 4 == 0");
    const int vec_xyz = (aligned16(xyz_feature) && ld_xyz % 4 == 0) ? 1 : 0;
    if (tail && vec_xyz && mlp->chain) {
        // forward-only rendering: no 1672-wide input row is assembled.  First layer = (xyz_feature rows in place) x (packed
        // k-blocks 0..21, columns past 1386 masked) + (A2) x (packed k-blocks 21..26); layers 1..3 + the sigmoid output on the
        // colour chain kernel, which adds the two partial results, the bias and the ReLU while it loads them
        float* A2 = CIN;
        float* Z1 = R[0];
        float* Z2 = R[1];
        color_hand_input2_kernel<<<nblocks(n * (HCIN2_LD / 4), 256), 256, 0, s>>>(feat, ld_feat, normal, n, A2);
        count_launch();
        HN_CHECK_LAUNCH();
        GemmArgs g1;
        g1.A = xyz_feature; g1.lda = ld_xyz;
        set_w(g1, mlp, 0);
        g1.M = (int)n; g1.N = 256; g1.K = HALO_DIM; g1.C = Z1; g1.ldc = 256;
        HN_PROPAGATE((gemm_nt<EPI_STORE>(g1, s, precision)));
        GemmArgs g2;
        g2.A = A2; g2.lda = HCIN2_LD;
        set_w(g2, mlp, 0);
        g2.B = mlp->W[0] + HCIN2_OFF; g2.bp_kb0 = HCIN2_KB0;
        g2.M = (int)n; g2.N = 256; g2.K = HCIN_DIM - HCIN2_OFF; g2.C = Z2; g2.ldc = 256;
        HN_PROPAGATE((gemm_nt<EPI_STORE>(g2, s, precision)));
        return chain::launch_color_tail_fwd(mlp, color_hand_tail_ops(mlp), Z1, Z2, 256, n, rgb, s);
    }
    color_hand_input_kernel<<<nblocks(n * (HCIN_LD / 4), 256), 256, 0, s>>>(xyz_feature, ld_xyz, vec_xyz, feat, ld_feat, normal, n, CIN);
    count_launch();
    HN_CHECK_LAUNCH();
    for (int l = 0; l < 4; ++l) {
        GemmArgs g;
        g.A = l == 0 ? CIN : R[l - 1]; g.lda = l == 0 ? HCIN_LD : 256;
        set_w(g, mlp, l);
        g.M = (int)n; g.N = 256; g.K = mlp->in_dim[l];
        g.C = R[l]; g.ldc = 256; g.bias = mlp->b[l];
        HN_PROPAGATE((gemm_nt<EPI_BIAS_RELU>(g, s, precision)));
    }
    GemmArgs g;
    g.A = R[3]; g.lda = 256; set_w(g, mlp, 4);
    g.M = (int)n; g.N = 3; g.K = 256; g.C = rgb; g.ldc = 3; g.bias = mlp->b[4];
    HN_PROPAGATE((gemm_nt<EPI_BIAS_SIGMOID>(g, s, precision)));
    return HN_OK;
}

int hn_color_hand_bwd(const hn_mlp_t* mlp, int64_t n, float* stash, const float* rgb, const float* d_rgb,
                      float* d_xyz_feature, int64_t ld_dxyz, float* d_feat, int64_t ld_dfeat, float* d_normal,
                      const hn_mlp_grad_t* grad, float* ws, int64_t ws_floats, int precision, hn_stream_t stream) {
    HN_PROPAGATE(check_color_hand_mlp(mlp));
    precision = base_precision(precision);
    HN_REQUIRE(precision_supported(precision), "hn_color_hand_bwd: precision %d not supported", precision);
    HN_REQUIRE(n >= 0 && n < (1ll << 31), "n_pts out of range");
    if (n == 0) return HN_OK;
    HN_REQUIRE(stash && rgb && d_rgb && ws, "hn_color_hand_bwd: null pointer");
    HN_REQUIRE(ws_floats >= hn_color_hand_ws_floats(n, HN_WS_BWD) && aligned16(ws), "workspace too small or misaligned");
    cudaStream_t s = (cudaStream_t)stream;
    float* CIN = stash;
    float* R[4];
    for (int l = 0; l < 4; ++l) R[l] = stash + n * HCIN_LD + (int64_t)l * n * 256;
    float* DZ[2] = {ws, ws + n * 256};
    float* DZ4 = ws + 2 * n * 256;
    float* DCIN = DZ4 + n * 4;
    const int splits_target = 2 * sm_count();
    auto dw_gemm = [&](const float* P, int64_t ldp, int out, const float* Q, int64_t ldq, int in, int l) -> int {
        if (!grad || !grad->dW[l]) return HN_OK;
        GemmArgs g;
        g.A = P; g.lda = ldp; g.B = Q; g.ldb = ldq;
        g.M = out; g.N = in; g.K = (int)n;
        g.C = grad->dW[l]; g.ldc = mlp->ld[l];
        int tiles = (int)(ceil_div(out, 128) * ceil_div(in, 256));
        int splits = (int)max((int64_t)1, min((int64_t)ceil_div(splits_target, tiles), ceil_div(n, 256)));
        return gemm_tn(g, s, precision, splits);
    };
    sigmoid_bwd4_kernel<<<nblocks(n * 4, 256), 256, 0, s>>>(rgb, d_rgb, n, DZ4);
    count_launch();
    HN_CHECK_LAUNCH();
    const float* dz = DZ4;
    int64_t ld_dz = 4;
    for (int l = 4; l >= 1; --l) {
        HN_PROPAGATE(dw_gemm(dz, ld_dz, mlp->out_dim[l], R[l - 1], 256, 256, l));
        if (grad && grad->db[l]) HN_PROPAGATE(launch_colsum(dz, ld_dz, n, mlp->out_dim[l], 1.0f, grad->db[l], s));
        GemmArgs g;
        g.A = dz; g.lda = ld_dz; set_w(g, mlp, l);
        g.M = (int)n; g.N = 256; g.K = mlp->out_dim[l];
        g.C = DZ[l & 1]; g.ldc = 256; g.aux1 = R[l - 1]; g.ldaux1 = 256;
        HN_PROPAGATE((gemm_nn<EPI_RELU_BWD>(g, s, precision)));
        dz = DZ[l & 1]; ld_dz = 256;
    }
    HN_PROPAGATE(dw_gemm(dz, 256, 256, CIN, HCIN_LD, HCIN_DIM, 0));
    if (grad && grad->db[0]) HN_PROPAGATE(launch_colsum(dz, 256, n, 256, 1.0f, grad->db[0], s));
    if (d_xyz_feature || d_feat || d_normal) {
        GemmArgs g;
        g.A = dz; g.lda = 256; set_w(g, mlp, 0);
        g.M = (int)n; g.N = HCIN_DIM; g.K = 256; g.C = DCIN; g.ldc = HCIN_LD;
        HN_PROPAGATE((gemm_nn<EPI_STORE>(g, s, precision)));
        const int vec = ((!d_xyz_feature || (aligned16(d_xyz_feature) && ld_dxyz % 4 == 0)) &&
                         (!d_feat || (aligned16(d_feat) && ld_dfeat % 4 == 0))) ? 1 : 0;
        color_hand_input_bwd_kernel<<<nblocks(n * ((HALO_DIM + 3) / 4 + 64 + 3), 256), 256, 0, s>>>(
            CIN, DCIN, n, d_xyz_feature, ld_dxyz, d_feat, ld_dfeat, d_normal, vec);
        count_launch();
        HN_CHECK_LAUNCH();
    }
    return HN_OK;
}

}  // extern "C"
