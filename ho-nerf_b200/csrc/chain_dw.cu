// Weight-gradient contractions of the fused chain path (HN_TC_BF16X3):
//     dW_l = sum over points of  P_a[p,:]^T Q_a[p,:]  (+ P_b[p,:]^T Q_b[p,:])
// for ALL layers of a network in ONE launch.  The reduction dimension is the point index, so both
// operands are consumed "MN-major" (feature index contiguous) exactly as the chain kernels leave them in
// HBM: fp32 rows are split into bf16 hi + lo while being staged into shared memory (three MMAs per
// product, fp32 accumulation in TMEM: 2 x [128 x 256] accumulators = all 512 TMEM columns).
// A CTA owns one (layer, point-range) pair; partial sums go to a workspace and a second kernel adds them
// into the packed gradient (no fp32 atomics on the 65 536-element tiles).  Bias gradients (column sums of
// the first operand) are accumulated by the staging threads on the way.
#include <algorithm>

#include "chain_common.cuh"
#include "chain_dw.cuh"

namespace hn {
namespace chain {

constexpr int DW_SWARPS = 16;            // staging / epilogue warps
constexpr int DW_THREADS = DW_SWARPS * 32 + 32;   // + 1 MMA warp
constexpr int DW_CPW = 32 / DW_SWARPS;   // 8-feature chunks per warp and operand
constexpr int DW_KP = 32;                // points per stage
constexpr int DW_OPER_BYTES = 256 * DW_KP * 2;        // one bf16 matrix of a stage: 16 KB
constexpr int DW_STAGE_BYTES = 4 * DW_OPER_BYTES;     // P_hi, P_lo, Q_hi, Q_lo
constexpr int DW_STAGES = 3;
constexpr int DW_SMEM_BYTES = DW_STAGES * DW_STAGE_BYTES + 1024;
constexpr int DW_SBO = 1024;             // next group of 8 points
constexpr int DW_LBO = 4 * 1024;         // next block of 64 features

__device__ __forceinline__ uint64_t make_desc_mn_sw128(uint32_t smem_addr) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);
    d |= (uint64_t)(DW_LBO >> 4) << 16;
    d |= (uint64_t)(DW_SBO >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}
__host__ __device__ constexpr uint32_t make_idesc_mn(uint32_t M, uint32_t N) {
    return (1u << 4) | (tc::FMT_BF16 << 7) | (tc::FMT_BF16 << 10) | (1u << 15) | (1u << 16) | ((N >> 3) << 17) |
           ((M >> 4) << 24);
}

// element address (floats) of (point p, column c) in the tiled stash layout
__host__ __device__ __forceinline__ int64_t tiled_off(int64_t p, int c) {
    return ((p >> 7) * 64 + (c >> 2)) * 512 + (p & 127) * 4 + (c & 3);
}

__device__ __forceinline__ void load8(const DwOperand& op, int64_t p, int64_t n, int fc, float* v) {
    const int c = fc * 8;
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = 0.0f;
    if (p >= n || c >= op.cols) return;
    if (op.tiled) {
        const float4 a = ld4(op.ptr + tiled_off(p, c));
        const float4 b = ld4(op.ptr + tiled_off(p, c + 4));
        v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
    } else if ((op.ld & 3) == 0 && c + 8 <= op.cols) {
        const float4 a = ld4(op.ptr + p * op.ld + c);
        const float4 b = ld4(op.ptr + p * op.ld + c + 4);
        v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
    } else {
#pragma unroll
        for (int i = 0; i < 8; ++i)
            if (c + i < op.cols) v[i] = op.ptr[p * op.ld + c + i];
    }
    if (c + 8 > op.cols) {
#pragma unroll
        for (int i = 0; i < 8; ++i)
            if (c + i >= op.cols) v[i] = 0.0f;
    }
}
// chunk of 8 features `fc` of point k (0..31) of a stage matrix
__device__ __forceinline__ void store8_mn(uint8_t* hi_base, uint8_t* lo_base, int k, int fc, const float* v) {
    uint4 hi, lo;
    split2(v[0], v[1], hi.x, lo.x);
    split2(v[2], v[3], hi.y, lo.y);
    split2(v[4], v[5], hi.z, lo.z);
    split2(v[6], v[7], hi.w, lo.w);
    const uint32_t off = (uint32_t)(fc >> 3) * DW_LBO + (uint32_t)(k >> 3) * DW_SBO + tc::sw128_offset((uint32_t)(k & 7), (uint32_t)(fc & 7));
    *reinterpret_cast<uint4*>(hi_base + off) = hi;
    *reinterpret_cast<uint4*>(lo_base + off) = lo;
}

__global__ void __launch_bounds__(DW_THREADS, 1) dw_kernel(const __grid_constant__ DwParams p) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    __shared__ uint64_t full[DW_STAGES], empty[DW_STAGES], done;
    __shared__ uint32_t tmem_base_s;
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int j = blockIdx.x / DW_SPLITS, split = blockIdx.x - j * DW_SPLITS;
    const DwJob& job = p.job[j];
    const int t0 = (int)((int64_t)p.n_tiles * split / DW_SPLITS), t1 = (int)((int64_t)p.n_tiles * (split + 1) / DW_SPLITS);
    const int n_iters = (t1 - t0) * (TILE_M / DW_KP) * job.n_pairs;       // stages this CTA runs through

    if (warp == 0) tc::tmem_alloc(&tmem_base_s, 512);
    if (tid == 32) {
        for (int s = 0; s < DW_STAGES; ++s) {
            tc::mbar_init(&full[s], DW_SWARPS * 32);
            tc::mbar_init(&empty[s], 1);
        }
        tc::mbar_init(&done, 1);
        tc::mbar_fence_init();
    }
    tc::tc_fence_before_sync();
    __syncthreads();
    tc::tc_fence_after_sync();
    const uint32_t tmem = tmem_base_s;

    if (warp == DW_SWARPS) {
        // ---- MMA issuer ---------------------------------------------------------------------------
        if (lane == 0 && n_iters > 0) {
            const uint32_t idesc = make_idesc_mn(128, (uint32_t)job.n_mma);
            uint32_t stage = 0, phase = 0;
            for (int it = 0; it < n_iters; ++it) {
                tc::mbar_wait(&full[stage], phase);
                tc::tc_fence_after_sync();
                const uint32_t base = tc::smem_u32(smem) + stage * DW_STAGE_BYTES;
                const uint32_t Phi = base, Plo = base + DW_OPER_BYTES, Qhi = base + 2 * DW_OPER_BYTES, Qlo = base + 3 * DW_OPER_BYTES;
#pragma unroll
                for (int k = 0; k < DW_KP / 16; ++k) {
                    const uint32_t koff = (uint32_t)k * 2 * DW_SBO;
#pragma unroll
                    for (int mc = 0; mc < 2; ++mc) {
                        const uint32_t moff = (uint32_t)mc * 2 * DW_LBO + koff;
                        const uint32_t acc = tmem + (uint32_t)mc * 256;
                        const uint32_t first = (it | k) != 0;
                        tc::umma_f16(acc, make_desc_mn_sw128(Plo + moff), make_desc_mn_sw128(Qhi + koff), idesc, first);
                        tc::umma_f16(acc, make_desc_mn_sw128(Phi + moff), make_desc_mn_sw128(Qlo + koff), idesc, 1);
                        tc::umma_f16(acc, make_desc_mn_sw128(Phi + moff), make_desc_mn_sw128(Qhi + koff), idesc, 1);
                    }
                }
                tc::umma_commit(&empty[stage]);
                if (++stage == DW_STAGES) { stage = 0; phase ^= 1u; }
            }
            tc::umma_commit(&done);
        }
    } else {
        // ---- staging: HBM fp32 -> bf16 hi/lo MN-major tiles; lane = point, warp w owns feature chunks w, w+8, ... --
        float bsum[DW_CPW][8];
#pragma unroll
        for (int a = 0; a < DW_CPW; ++a)
#pragma unroll
            for (int i = 0; i < 8; ++i) bsum[a][i] = 0.0f;
        uint32_t stage = 0, phase = 0;
        int it = 0;
        for (int tile = t0; tile < t1; ++tile) {
            for (int sub = 0; sub < TILE_M / DW_KP; ++sub) {
                const int64_t pnt = (int64_t)tile * TILE_M + sub * DW_KP + lane;
                for (int pair = 0; pair < job.n_pairs; ++pair, ++it) {
                    float vp[DW_CPW][8], vq[DW_CPW][8];
#pragma unroll
                    for (int a = 0; a < DW_CPW; ++a) {
                        load8(job.P[pair], pnt, p.n, warp + DW_SWARPS * a, vp[a]);
                        load8(job.Q[pair], pnt, p.n, warp + DW_SWARPS * a, vq[a]);
                    }
                    tc::mbar_wait(&empty[stage], phase ^ 1u);
                    uint8_t* base = smem + stage * DW_STAGE_BYTES;
#pragma unroll
                    for (int a = 0; a < DW_CPW; ++a) {
                        store8_mn(base, base + DW_OPER_BYTES, lane, warp + DW_SWARPS * a, vp[a]);
                        store8_mn(base + 2 * DW_OPER_BYTES, base + 3 * DW_OPER_BYTES, lane, warp + DW_SWARPS * a, vq[a]);
                        if (pair == 0) {
#pragma unroll
                            for (int i = 0; i < 8; ++i) bsum[a][i] += vp[a][i];
                        }
                    }
                    tc::fence_proxy_async_smem();
                    tc::mbar_arrive(&full[stage]);
                    if (++stage == DW_STAGES) { stage = 0; phase ^= 1u; }
                }
            }
        }
        if (job.db) {
#pragma unroll
            for (int a = 0; a < DW_CPW; ++a)
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const float s = warp_sum(bsum[a][i]);
                    const int c = (warp + DW_SWARPS * a) * 8 + i;
                    if (lane == 0 && c < job.P[0].cols && s != 0.0f) atomicAdd(job.db + c, s * job.db_scale);
                }
        }
        // ---- epilogue: partial sums -> workspace ----------------------------------------------------------
        float* part = p.part + ((size_t)j * DW_SPLITS + split) * 65536;
        const int q = warp & 3, cgrp = warp >> 2;
        constexpr int CW = 256 / (DW_SWARPS / 4);       // accumulator columns per warp
        if (n_iters > 0) {
            tc::mbar_wait(&done, 0);
            tc::tc_fence_after_sync();
        }
        for (int mc = 0; mc < 2; ++mc) {
            const int row = mc * 128 + q * 32 + lane;
            for (int blk = 0; blk < CW / 32; ++blk) {
                const int col0 = cgrp * CW + blk * 32;
                float v[32];
                if (n_iters > 0 && col0 < job.n_mma) {
                    acc_load32(tmem + (uint32_t)mc * 256, q * 32, col0, v);
                } else {
#pragma unroll
                    for (int i = 0; i < 32; ++i) v[i] = 0.0f;
                }
#pragma unroll
                for (int i = 0; i < 32; i += 4) st4(part + (size_t)row * 256 + col0 + i, make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]));
            }
        }
    }
    tc::tc_fence_before_sync();
    __syncthreads();
    if (warp == 0) tc::tmem_dealloc(tmem, 512);
}

__global__ void dw_reduce_kernel(const __grid_constant__ DwReduceParams p) {
    const DwReduceJob& jb = p.job[blockIdx.y];
    if (!jb.dW) return;
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;       // float4 index in the [256][256] tile
    const int r = idx >> 6, c = (idx & 63) * 4;
    if (r >= jb.rows || c >= jb.cols) return;
    const float* src = p.part + (size_t)blockIdx.y * DW_SPLITS * 65536 + (size_t)r * 256 + c;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int s = 0; s < DW_SPLITS; ++s) {
        const float4 v = ld4(src + (size_t)s * 65536);
        acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
    }
    float* dst = jb.dW + (size_t)(jb.row0 + r) * jb.ld;
    const float a[4] = {acc.x, acc.y, acc.z, acc.w};
#pragma unroll
    for (int i = 0; i < 4; ++i)
        if (c + i < jb.cols) dst[c + i < jb.csplit ? jb.col0 + c + i : jb.col1 + (c + i - jb.csplit)] += a[i];
}

int64_t dw_part_floats(int n_jobs) { return (int64_t)n_jobs * DW_SPLITS * 65536; }

int launch_dw(const DwParams& p, const DwReduceParams& r, cudaStream_t s) {
    static bool configured = false;
    if (!configured) {
        HN_CHECK_CUDA(cudaFuncSetAttribute(dw_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, DW_SMEM_BYTES));
        configured = true;
    }
    {
        TimingScope ts(s, TT_DW);
        dw_kernel<<<p.n_jobs * DW_SPLITS, DW_THREADS, DW_SMEM_BYTES, s>>>(p);
    }
    count_launch();
    HN_CHECK_LAUNCH();
    dw_reduce_kernel<<<dim3(65536 / 4 / 256, p.n_jobs), 256, 0, s>>>(r);
    count_launch();
    HN_CHECK_LAUNCH();
    return HN_OK;
}

}  // namespace chain
}  // namespace hn

using namespace hn;

// Diagnostics / tests: C[out, in] (row-major, ld = in_pad4) = P^T Q (+ P2^T Q2) with the production kernel.
extern "C" int hn_dw_test(const float* P, int64_t ldp, int p_tiled, int out, const float* Q, int64_t ldq, int q_tiled,
                          int in, const float* P2, const float* Q2, int64_t n, float* C, int64_t ldc, float* db,
                          float* part, int64_t part_floats, hn_stream_t stream) {
    HN_REQUIRE(P && Q && C && part && out >= 1 && out <= 256 && in >= 1 && in <= 256, "hn_dw_test: bad arguments");
    HN_REQUIRE(part_floats >= chain::dw_part_floats(1), "hn_dw_test: partial-sum workspace too small");
    chain::DwParams p;
    p.n = n; p.n_tiles = (int)ceil_div(n, chain::TILE_M); p.n_jobs = 1; p.part = part;
    chain::DwJob& j = p.job[0];
    j.P[0] = {P, ldp, out, p_tiled}; j.Q[0] = {Q, ldq, in, q_tiled};
    j.P[1] = {P2, ldp, out, p_tiled}; j.Q[1] = {Q2, ldq, in, q_tiled};
    j.n_pairs = (P2 && Q2) ? 2 : 1;
    j.n_mma = (int)round_up(in, 16);
    j.db = db; j.db_scale = 1.0f;
    chain::DwReduceParams r;
    r.part = part;
    r.job[0] = chain::reduce_job(C, (int)ldc, 0, out, in);
    return chain::launch_dw(p, r, (cudaStream_t)stream);
}
