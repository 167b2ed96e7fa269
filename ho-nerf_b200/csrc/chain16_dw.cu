// Weight gradients of the HN_TC_MIXED16 path:  dW_l = sum over points of P_a[p,:]^T Q_a[p,:] (+ P_b^T Q_b)  for all
// layers of a net in ONE launch, from the 16-bit dW-ready tiles the sweep kernels leave in HBM (chain16.cuh).
//
// The reduction dimension is the point index, so both operands are MN-major; the tiles already are in the un-swizzled
// MN-major core-matrix order (8 points x 8 features = 128 contiguous bytes; next 8 points +128 B, next 8 features
// +1024 B inside a 64-point half tile), so a 64-point operand stage is ONE cp.async.bulk of 32 KB (8 KB for the 64-wide
// encoding arrays) and nobody converts anything: warp 0 streams stages (3 x 64 KB ring, mbarrier transaction
// counts), warp 1 issues one bf16 MMA per product (M = 2 x 128 output features, N = input features, K = 16 points,
// fp32 accumulation in all 512 TMEM columns), warps 2..5 add up the bias gradient (column sums of the first operand,
// read from the landed stage) and finally move the accumulators to the split-K workspace that chain_dw.cu's
// dw_reduce_kernel folds into the packed gradient.  HBM-bound by construction: 256 KB per (layer, tile) against
// 32 MMAs of 128 cycles.
#include <algorithm>

#include "chain16.cuh"
#include "chain_dw.cuh"

namespace hn {
namespace chain {

constexpr int D16_WARPS_EPI = 4;
constexpr int D16_THREADS = 64 + D16_WARPS_EPI * 32;
constexpr int D16_P_BYTES = 64 * 256 * 2;                 // one 64-point stage of P: 32 KB
constexpr int D16_STAGE_BYTES = 2 * D16_P_BYTES;          // P + Q (Q may use less)
constexpr int D16_STAGES = 3;
constexpr int D16_SMEM_BYTES = D16_STAGES * D16_STAGE_BYTES + 1024;
constexpr uint32_t D16_SBO = 1024;                        // next 8-feature chunk (MN direction)
constexpr uint32_t D16_LBO = 128;                         // next 8 points (K direction)

static int g_dw16_swap = 0;       // diagnostics: exchange LBO / SBO in the descriptors (hn_dw16_set_debug)

__device__ __forceinline__ uint64_t make_desc_mn_none(uint32_t smem_addr, uint32_t lbo, uint32_t sbo) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);
    d |= (uint64_t)(lbo >> 4) << 16;
    d |= (uint64_t)(sbo >> 4) << 32;
    d |= (uint64_t)1 << 46;              // descriptor version 1 (Blackwell); layout type 0 = no swizzle
    return d;
}
__host__ __device__ constexpr uint32_t make_idesc_mn16(uint32_t M, uint32_t N) {
    return (1u << 4) | (tc::FMT_BF16 << 7) | (tc::FMT_BF16 << 10) | (1u << 15) | (1u << 16) | ((N >> 3) << 17) | ((M >> 4) << 24);
}

struct Dw16Args {
    Dw16Params p;
    int swap;
    uint32_t* dbg;
};

__global__ void __launch_bounds__(D16_THREADS, 1) dw16_kernel(const __grid_constant__ Dw16Args a) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    __shared__ uint64_t full[D16_STAGES], empty[D16_STAGES], done;
    __shared__ uint32_t tmem_base_s;
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    const Dw16Params& p = a.p;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int j = blockIdx.x / DW_SPLITS, split = blockIdx.x - j * DW_SPLITS;
    const Dw16Job& job = p.job[j];
    const int t0 = (int)((int64_t)p.n_tiles * split / DW_SPLITS), t1 = (int)((int64_t)p.n_tiles * (split + 1) / DW_SPLITS);
    const int n_iters = (t1 - t0) * 2 * job.n_pairs;        // 64-point stages this CTA runs through
    const uint32_t q_bytes = (uint32_t)job.q_chunks * 1024u;
    const size_t q_tile_bytes = (size_t)job.q_chunks * 2048u;

    if (warp == 0) tc::tmem_alloc(&tmem_base_s, 512);
    if (tid == 32) {
        for (int s = 0; s < D16_STAGES; ++s) {
            tc::mbar_init(&full[s], 1);
            tc::mbar_init(&empty[s], 1 + D16_WARPS_EPI);     // MMA commit + the bias-gradient warps
        }
        tc::mbar_init(&done, 1);
        tc::mbar_fence_init();
    }
    tc::tc_fence_before_sync();
    __syncthreads();
    tc::tc_fence_after_sync();
    const uint32_t tmem = tmem_base_s;

    if (warp == 0) {
        // ---- producer: stage order (tile, half, pair) --------------------------------------------------------------
        if (lane == 0) {
            uint32_t stage = 0, phase = 0;
            for (int t = t0; t < t1; ++t)
                for (int h = 0; h < 2; ++h)
                    for (int pr = 0; pr < job.n_pairs; ++pr) {
                        dbg_mark(a.dbg, 0, (uint32_t)(t << 8 | h << 1 | pr));
                        tc::mbar_wait(&empty[stage], phase ^ 1u);
                        tc::mbar_arrive_expect_tx(&full[stage], (uint32_t)D16_P_BYTES + q_bytes);
                        uint8_t* dst = smem + stage * D16_STAGE_BYTES;
                        tc::bulk_g2s(dst, job.P[pr] + (size_t)t * T16_TILE_BYTES + (size_t)h * D16_P_BYTES, D16_P_BYTES, &full[stage]);
                        tc::bulk_g2s(dst + D16_P_BYTES, job.Q[pr] + (size_t)t * q_tile_bytes + (size_t)h * q_bytes, q_bytes,
                                     &full[stage]);
                        if (++stage == D16_STAGES) { stage = 0; phase ^= 1u; }
                    }
        }
    } else if (warp == 1) {
        // ---- MMA issuer ------------------------------------------------------------------------------------------------
        if (lane == 0 && n_iters > 0) {
            const uint32_t idesc = make_idesc_mn16(128, (uint32_t)job.n_mma);
            const uint32_t lbo = a.swap ? D16_SBO : D16_LBO, sbo = a.swap ? D16_LBO : D16_SBO;
            uint32_t stage = 0, phase = 0;
            for (int it = 0; it < n_iters; ++it) {
                dbg_mark(a.dbg, 1, (uint32_t)it);
                tc::mbar_wait(&full[stage], phase);
                tc::tc_fence_after_sync();
                const uint32_t P = tc::smem_u32(smem) + stage * D16_STAGE_BYTES, Q = P + D16_P_BYTES;
#pragma unroll
                for (int k = 0; k < 4; ++k) {            // 16 points per MMA = two 8-point core-matrix groups
                    const uint32_t koff = (uint32_t)k * 2u * D16_LBO;
                    const uint64_t dq = make_desc_mn_none(Q + koff, lbo, sbo);
#pragma unroll
                    for (int mc = 0; mc < 2; ++mc)       // output features [128 mc, 128 mc + 128): chunks 16 mc ..
                        tc::umma_f16(tmem + (uint32_t)mc * 256u, make_desc_mn_none(P + (uint32_t)mc * 16u * D16_SBO + koff, lbo, sbo), dq,
                                     idesc, (it | k) != 0);
                }
                tc::umma_commit(&empty[stage]);
                if (++stage == D16_STAGES) { stage = 0; phase ^= 1u; }
            }
            tc::umma_commit(&done);
        }
    } else {
        // ---- bias gradient: column sums of P[0] stages; warp w owns chunks 8 w .. 8 w + 7, lane = point (and point + 32)
        const int w = warp - 2;
        float bsum[8][8];
#pragma unroll
        for (int c = 0; c < 8; ++c)
#pragma unroll
            for (int i = 0; i < 8; ++i) bsum[c][i] = 0.0f;
        {
            uint32_t stage = 0, phase = 0;
            for (int it = 0; it < n_iters; ++it) {
                const bool first_pair = (((job.db_mask ? job.db_mask : 1) >> (it % job.n_pairs)) & 1) != 0;
                tc::mbar_wait(&full[stage], phase);         // never run ahead of the stage (the empty count includes us)
                if (job.db && first_pair) {
                    const uint8_t* P = smem + stage * D16_STAGE_BYTES;
#pragma unroll
                    for (int c = 0; c < 8; ++c)
#pragma unroll
                        for (int hp = 0; hp < 2; ++hp) {
                            const uint4 q = *reinterpret_cast<const uint4*>(P + (uint32_t)(8 * w + c) * 1024u + (uint32_t)(hp * 32 + lane) * 16u);
                            bsum[c][0] += bf16_lo(q.x); bsum[c][1] += bf16_hi(q.x);
                            bsum[c][2] += bf16_lo(q.y); bsum[c][3] += bf16_hi(q.y);
                            bsum[c][4] += bf16_lo(q.z); bsum[c][5] += bf16_hi(q.z);
                            bsum[c][6] += bf16_lo(q.w); bsum[c][7] += bf16_hi(q.w);
                        }
                }
                __syncwarp();
                if (lane == 0) tc::mbar_arrive(&empty[stage]);
                if (++stage == D16_STAGES) { stage = 0; phase ^= 1u; }
            }
        }
        if (job.db) {
#pragma unroll
            for (int c = 0; c < 8; ++c)
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const float s = warp_sum(bsum[c][i]);
                    if (lane == 0 && (8 * w + c) * 8 + i < job.p_cols) atomicAdd(job.db + (8 * w + c) * 8 + i, s);
                }
        }
        // ---- epilogue: accumulators -> split-K workspace [256][256] ------------------------------------------------------
        float* part = p.part + ((size_t)j * DW_SPLITS + split) * 65536;
        if (n_iters > 0) {
            tc::mbar_wait(&done, 0);
            tc::tc_fence_after_sync();
        }
        const int q = warp & 3;                 // a warp reads the TMEM lane quarter warp_id % 4
        for (int mc = 0; mc < 2; ++mc) {
            const int row = mc * 128 + q * 32 + lane;
            for (int col0 = 0; col0 < 256; col0 += 32) {
                float v[32];
                if (n_iters > 0 && col0 < job.n_mma) {
                    acc_load32(tmem + (uint32_t)mc * 256u, q * 32, col0, v);
                } else {
#pragma unroll
                    for (int i = 0; i < 32; ++i) v[i] = 0.0f;
                }
#pragma unroll
                for (int i = 0; i < 32; i += 4) st4(part + (size_t)row * 256 + col0 + i, make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]));
            }
        }
    }
    if (lane == 0 && warp < 3) dbg_mark(a.dbg, warp, 0xffffffffu);
    tc::tc_fence_before_sync();
    __syncthreads();
    if (warp == 0) tc::tmem_dealloc(tmem, 512);
}

int launch_dw16(const Dw16Params& p, const DwReduceParams& r, cudaStream_t s) {
    static bool configured = false;
    if (!configured) {
        HN_CHECK_CUDA(cudaFuncSetAttribute(dw16_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, D16_SMEM_BYTES));
        configured = true;
    }
    for (int j = 0; j < p.n_jobs; ++j)
        HN_REQUIRE(p.job[j].n_mma % 16 == 0 && p.job[j].n_mma >= 16 && p.job[j].n_mma <= 8 * p.job[j].q_chunks &&
                       (p.job[j].q_chunks == 32 || p.job[j].q_chunks == 8),
                   "launch_dw16: job %d has an unsupported operand shape", j);
    Dw16Args a;
    a.p = p;
    a.swap = g_dw16_swap;
    a.dbg = dbg_slot(3);
    {
        TimingScope ts(s, TT_DW);
        dw16_kernel<<<p.n_jobs * DW_SPLITS, D16_THREADS, D16_SMEM_BYTES, s>>>(a);
    }
    count_launch();
    HN_CHECK_LAUNCH();
    return launch_dw_reduce(r, p.n_jobs, s);
}

// dW_out[0, :] += inv_scale * sum_p (d_sdf[p] * h7[p, :] + u7[p, :]);  db_out[0] += inv_scale * sum_p d_sdf[p]
// (row 0 of the output layer: the sdf value uses it directly, the normal sweep is seeded with it) from the bf16 tiles.
// grid (32 chunks, splits), 128 threads = the points of a tile.
__global__ void __launch_bounds__(128) out_row0_grad16_kernel(const uint8_t* __restrict__ H7, const uint8_t* __restrict__ U7,
                                                              const float* __restrict__ d_sdf, int64_t n, int n_tiles,
                                                              float inv_scale, float* __restrict__ dW_row0,
                                                              float* __restrict__ db0) {
    __shared__ float red[4][9];
    const int f8 = blockIdx.x, p = threadIdx.x;
    const int t0 = (int)((int64_t)n_tiles * blockIdx.y / gridDim.y), t1 = (int)((int64_t)n_tiles * (blockIdx.y + 1) / gridDim.y);
    float a[9];
#pragma unroll
    for (int i = 0; i < 9; ++i) a[i] = 0.0f;
    for (int t = t0; t < t1; ++t) {
        const int64_t pnt = (int64_t)t * TILE_M + p;
        if (pnt >= n) continue;
        const float w = d_sdf ? d_sdf[pnt] : 0.0f;
        const uint4 h = ldg16(H7 + (size_t)t * T16_TILE_BYTES + t16_off(p, f8));
        const uint4 u = ldg16(U7 + (size_t)t * T16_TILE_BYTES + t16_off(p, f8));
        a[0] += w * bf16_lo(h.x) + bf16_lo(u.x); a[1] += w * bf16_hi(h.x) + bf16_hi(u.x);
        a[2] += w * bf16_lo(h.y) + bf16_lo(u.y); a[3] += w * bf16_hi(h.y) + bf16_hi(u.y);
        a[4] += w * bf16_lo(h.z) + bf16_lo(u.z); a[5] += w * bf16_hi(h.z) + bf16_hi(u.z);
        a[6] += w * bf16_lo(h.w) + bf16_lo(u.w); a[7] += w * bf16_hi(h.w) + bf16_hi(u.w);
        a[8] += w;
    }
#pragma unroll
    for (int i = 0; i < 9; ++i) {
        a[i] = warp_sum(a[i]);
        if ((p & 31) == 0) red[p >> 5][i] = a[i];
    }
    __syncthreads();
    if (p < 9) {
        const float v = (red[0][p] + red[1][p] + red[2][p] + red[3][p]) * inv_scale;
        if (p < 8) {
            if (dW_row0) atomicAdd(dW_row0 + f8 * 8 + p, v);
        } else if (f8 == 0 && db0) {
            atomicAdd(db0, v);
        }
    }
}

int launch_out_row0_grad16(const uint8_t* H7, const uint8_t* U7, const float* d_sdf, int64_t n, int n_tiles, float inv_scale,
                           float* dW_row0, float* db0, cudaStream_t s) {
    const int splits = std::max(1, std::min(n_tiles, 64));
    out_row0_grad16_kernel<<<dim3(32, splits), 128, 0, s>>>(H7, U7, d_sdf, n, n_tiles, inv_scale, dW_row0, db0);
    count_launch();
    HN_CHECK_LAUNCH();
    return HN_OK;
}

// fp32 row-major [n, cols] -> bf16 T16 tiles (tests / diagnostics; the product's tiles are written by the sweep kernels)
__global__ void to_t16_kernel(const float* __restrict__ src, int64_t ld, int64_t n, int cols, int nch, uint8_t* __restrict__ dst) {
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;      // (point, chunk)
    const int64_t np = (n + TILE_M - 1) / TILE_M * TILE_M;
    if (idx >= np * nch) return;
    const int64_t pnt = idx / nch;
    const int f8 = (int)(idx - pnt * nch);
    float v[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = (pnt < n && f8 * 8 + i < cols) ? src[pnt * ld + f8 * 8 + i] : 0.0f;
    uint4 q;
    q.x = pack_bf16x2(v[0], v[1]); q.y = pack_bf16x2(v[2], v[3]); q.z = pack_bf16x2(v[4], v[5]); q.w = pack_bf16x2(v[6], v[7]);
    const int p = (int)(pnt & 127);
    const uint32_t off = nch == 32 ? t16_off<32>(p, f8) : t16_off<8>(p, f8);
    stg16(dst + (size_t)(pnt >> 7) * ((size_t)nch * 2048u) + off, q);
}

}  // namespace chain
}  // namespace hn

using namespace hn;

extern "C" {

int hn_dw16_set_debug(int swap_lbo_sbo) {
    chain::g_dw16_swap = swap_lbo_sbo;
    return HN_OK;
}

// Diagnostics / tests: C[out, in] (row-major, ld = ldc) = P^T Q (+ P2^T Q2) and db = colsum(P) with the production
// kernel; the fp32 operands are first rounded to bf16 tiles in `tiles` (4 * round_up(n,128) * 512 bytes).
int hn_dw16_test(const float* P, int out, const float* Q, int in, const float* P2, const float* Q2, int64_t n, float* C,
                 int64_t ldc, float* db, void* tiles, int64_t tiles_bytes, float* part, int64_t part_floats,
                 hn_stream_t stream) {
    HN_REQUIRE(P && Q && C && part && tiles && out >= 1 && out <= 256 && in >= 1 && in <= 256, "hn_dw16_test: bad arguments");
    HN_REQUIRE(part_floats >= chain::dw_part_floats(1), "hn_dw16_test: partial-sum workspace too small");
    const int64_t np = round_up(n, chain::TILE_M);
    HN_REQUIRE(tiles_bytes >= 4 * np * 512 && aligned16(tiles), "hn_dw16_test: tile workspace too small");
    cudaStream_t s = (cudaStream_t)stream;
    const int nch = in <= 64 ? 8 : 32;
    uint8_t* t = reinterpret_cast<uint8_t*>(tiles);
    uint8_t* tp[2] = {t, t + 2 * np * 512};
    uint8_t* tq[2] = {t + np * 512, t + 3 * np * 512};
    const float* ps[2] = {P, P2};
    const float* qs[2] = {Q, Q2};
    const int pairs = (P2 && Q2) ? 2 : 1;
    for (int a = 0; a < pairs; ++a) {
        chain::to_t16_kernel<<<(unsigned)ceil_div(np * 32, 256), 256, 0, s>>>(ps[a], out, n, out, 32, tp[a]);
        chain::to_t16_kernel<<<(unsigned)ceil_div(np * nch, 256), 256, 0, s>>>(qs[a], in, n, in, nch, tq[a]);
    }
    HN_CHECK_LAUNCH();
    chain::Dw16Params p = {};
    p.n_tiles = (int)(np / chain::TILE_M); p.n_jobs = 1; p.part = part;
    chain::Dw16Job& j = p.job[0];
    j.P[0] = tp[0]; j.Q[0] = tq[0]; j.P[1] = tp[pairs - 1]; j.Q[1] = tq[pairs - 1];
    j.n_pairs = pairs; j.q_chunks = nch; j.n_mma = (int)round_up(in, 16); j.db = db; j.p_cols = out;
    chain::DwReduceParams r;
    r.part = part;
    r.job[0] = chain::reduce_job(C, (int)ldc, 0, out, in);
    return chain::launch_dw16(p, r, s);
}

}  // extern "C"
