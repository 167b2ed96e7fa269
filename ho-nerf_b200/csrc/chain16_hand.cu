// Hand SDF field (HALO pose-conditioned, utils/fields.py:56-177), HN_TC_MIXED16: the 256 x 256 layers of the value trunk,
// the normal sweep, the tangent sweep and the reverse sweep as three persistent tile-chain kernels with the activations in
// tensor memory (the machinery of chain16.cuh, built for the object field), three 16-bit MMAs per product.
//
// The 1386-wide contractions INTO the chain (HALO feature -> layer 0 / skip layer 4) stay on the per-layer kernels
// (gemm_bx3.cuh) and meet it through fp32 row-major [points, 256] buffers; the ones OUT of it (cotangents of the feature)
// are extra steps of the sweeps: while D_4 (resp. D_0) is the A operand in tensor memory, six 256-column chunks of
// W_4[:, 256:] (resp. W_0) stream through the weight ring and the epilogue writes (resp. adds to) FB, kept as column-major
// [1388][128 points] fp32 tiles (coalesced for the epilogue and for the thread-per-point HALO kernels that consume it).
//   hand_trunk16_kernel   in : H0 = softplus(F W_0^T + b_0), ZF4 = F W_4[:, 256:]^T       out: sdf, feature, EM / EML tiles
//   hand_nsweep16_kernel  in : EM / EML                                                    out: D16 tiles, FB tiles
//   hand_bwd16_kernel     in : Q0 = tF W_0^T, QF4 = tF W_4[:, 256:]^T, EM, D16, d_sdf / d_feat
//                                                                                          out: DF tiles (X16 scratch)
// Used when no weight gradient is asked for (pose fitting, rendering): the weight-gradient contractions of the hand net
// stay on the per-layer path (fields_hand.cu), which keeps an fp32 stash.
#include <algorithm>

#include "chain16.cuh"
#include "chain_hand_layout.cuh"

namespace hn {
namespace chain {

// ------------------------------------------------------------------------------------------------------------------
// value trunk (layers 1..7 + feature head); TMEM columns [0,256) accumulator, [256,384) A_hi, [384,512) A_lo
// ------------------------------------------------------------------------------------------------------------------
constexpr int HT_STAGE_BYTES = 256 * 128;        // 32 KB: one k-block of a feature tile (hi + lo) or 256 weight rows; half steps use 16 KB
constexpr int HT_STAGES = 6;
constexpr int HT_HEAD_OFF = HT_STAGES * HT_STAGE_BYTES;
constexpr int HT_SMEM_BYTES = HT_HEAD_OFF + EPI_CGROUPS * TILE_M * 4 + 1024;
constexpr uint32_t HT_A_HI = 256, HT_A_LO = 384;

// ---- 2-CTA cluster helpers (MC: the two CTAs of a pair run the same weight stream; each fetches HALF of every weight stage and
// multicasts it into both CTAs' rings, which halves the L2 -> SM weight traffic the 1386-wide layers are co-limited by) ----------
__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// tcgen05.commit whose arrival lands on the mbarrier at the same CTA-relative address in every CTA of `mask` (SASS: UTCBAR.MULTICAST)
__device__ __forceinline__ void umma_commit_multicast(uint64_t* bar, uint16_t mask) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 ::"r"(tc::smem_u32(bar)), "h"(mask)
                 : "memory");
}
// bulk copy global -> the same CTA-relative shared address in every CTA of `mask`, completing on each one's mbarrier
__device__ __forceinline__ void bulk_g2s_multicast(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar, uint16_t mask) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1], %2, [%3], %4;"
                 ::"r"(tc::smem_u32(smem_dst)), "l"(gsrc), "r"(bytes), "r"(tc::smem_u32(bar)), "h"(mask)
                 : "memory");
}

struct HtBarriers {
    uint64_t full[HT_STAGES];
    uint64_t empty[HT_STAGES];
    uint64_t a_ready;
    uint64_t acc_full[2];    // one per N-half: each completes once per layer (see SwBarriers)
    uint32_t tmem_base;
};

// One MMA step of the trunk.  Half steps (ts form): acc[:, 128 h .. +128] = A (tensor memory, hi + lo fp16) @ B^T, B = 128 weight
// rows x 64 k per stage (hi stage, lo stage).  Feature steps (ss form): acc[:, 0 .. 256] (+)= F @ B^T over 22 k-blocks, A = the
// fp16 pair tiles of the HALO feature bulk-copied from HBM (one 32 KB stage per k-block), B = 256 weight rows (hi stage, lo stage).
struct HtStep {
    uint32_t b_off;
    uint16_t acc_col;
    uint8_t kblocks;
    uint8_t feature : 1;     // ss form over the feature tiles
    uint8_t no_wait : 1;     // do not wait for a_ready (second half of a layer / the feature part of the skip layer)
    uint8_t acc_in : 1;      // accumulate onto what the accumulator holds
    uint8_t commit : 2;      // 0: none (more MMAs of this layer follow), 1: this half's acc_full, 2: both
};
struct HtProgram {
    int n_steps;
    HtStep step[20];
};

struct HandTrunkParams {
    int64_t n;
    const uint8_t* F16;      // fp16 pair tiles of the HALO feature (halo_feature16_kernel); NULL: legacy inputs H0 / ZF4
    const float* H0;         // [np, 256] fp32 rows
    const float* ZF4;        // [np, 256] fp32 rows: the feature part of the skip layer's pre-activation
    float* sdf;
    float* feat;             // NULL: sdf only (no feature head, no stash)
    int64_t ld_feat;
    uint8_t* EM[8];          // fp16 T16 tiles: em = exp(-100 h) rounded to fp16 (NULL with feat == NULL)
    uint8_t* EML[8];         // fp16 T16 tiles: fp16(em - fp16(em))
    const uint8_t* chain;
    const float* bias[9];
    const float* w_out0;
    int n_tiles;
};

template <bool MC>
__global__ void __launch_bounds__(THREADS, 1)
hand_trunk16_kernel(const __grid_constant__ HandTrunkParams p, const __grid_constant__ HtProgram prog) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    __shared__ HtBarriers bar;
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    float* s_head = reinterpret_cast<float*>(smem + HT_HEAD_OFF);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (warp == 0) tc::tmem_alloc(&bar.tmem_base, 512);
    if (threadIdx.x == 32) {
        for (int s = 0; s < HT_STAGES; ++s) {
            tc::mbar_init(&bar.full[s], 1);
            // MC: a stage is free when BOTH CTAs' MMAs have released it (the peer's weight copy lands in this ring too): every
            // release is a multicast commit, so each `empty` barrier collects two arrivals per lap, in hardware
            tc::mbar_init(&bar.empty[s], MC ? 2 : 1);
        }
        tc::mbar_init(&bar.a_ready, EPI_THREADS);
        tc::mbar_init(&bar.acc_full[0], 1);
        tc::mbar_init(&bar.acc_full[1], 1);
        tc::mbar_fence_init();
    }
    tc::tc_fence_before_sync();
    __syncthreads();
    tc::tc_fence_after_sync();
    if (MC) cluster_sync_all();          // the peer's barriers are initialised before anything is sent to them
    const uint32_t tmem = bar.tmem_base;
    // tile walk.  MC: clusters walk over PAIRS of tiles, rank r takes tile 2 q + r, both CTAs of a cluster run the same number of
    // iterations (a tile index past the end is a dummy: its feature rows come from the last tile, nothing is stored)
    const uint32_t rank = MC ? cluster_ctarank() : 0u;
    const int walkers = MC ? (int)(gridDim.x >> 1) : (int)gridDim.x;
    const int me = MC ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;
    const int n_units = MC ? (p.n_tiles + 1) / 2 : p.n_tiles;
    const int n_my_tiles = n_units > me ? (n_units - me + walkers - 1) / walkers : 0;
    auto tile_of = [&](int t) -> int64_t {
        const int64_t u = (int64_t)me + (int64_t)t * walkers;
        return MC ? 2 * u + rank : u;
    };
    const bool stash = p.feat != nullptr;

    if (warp == 0) {
        if (lane == 0) {
            uint32_t stage = 0, phase = 0;
            // a stage of this CTA's own data (feature tiles)
            auto put = [&](const uint8_t* src, uint32_t bytes) {
                tc::mbar_wait(&bar.empty[stage], phase ^ 1u);
                tc::mbar_arrive_expect_tx(&bar.full[stage], bytes);
                tc::bulk_g2s(smem + stage * HT_STAGE_BYTES, src, bytes, &bar.full[stage]);
                if (++stage == HT_STAGES) { stage = 0; phase ^= 1u; }
            };
            // a stage of WEIGHTS: identical in both CTAs of the pair.  MC: once both CTAs released the stage, fetch half of it and
            // multicast it into both rings (each CTA's `full` barrier expects the whole stage: half from its own copy, half from
            // the peer's)
            auto put_w = [&](const uint8_t* src, uint32_t bytes) {
                if (!MC) { put(src, bytes); return; }
                tc::mbar_wait(&bar.empty[stage], phase ^ 1u);
                tc::mbar_arrive_expect_tx(&bar.full[stage], bytes);
                const uint32_t half = bytes >> 1;
                bulk_g2s_multicast(smem + stage * HT_STAGE_BYTES + rank * half, src + rank * half, half, &bar.full[stage], (uint16_t)3);
                if (++stage == HT_STAGES) { stage = 0; phase ^= 1u; }
            };
            for (int t = 0; t < n_my_tiles; ++t) {
                const int64_t tile = tile_of(t) < p.n_tiles ? tile_of(t) : (int64_t)p.n_tiles - 1;
                const uint8_t* ftile = p.F16 + (size_t)tile * HAND_F16_TILE_BYTES;
                for (int s = 0; s < prog.n_steps; ++s) {
                    const HtStep st = prog.step[s];
                    const uint8_t* src = p.chain + st.b_off;
                    if (st.feature) {
                        for (int kb = 0; kb < st.kblocks; ++kb) {
                            put(ftile + (size_t)kb * HAND_F16_KB_BYTES, HAND_F16_KB_BYTES);      // A: hi tile + lo tile
                            put_w(src + (size_t)kb * 65536, 32768);                              // B hi: 256 rows
                            put_w(src + (size_t)kb * 65536 + 32768, 32768);                      // B lo
                        }
                    } else {
                        for (int c = 0; c < 2 * st.kblocks; ++c) put_w(src + (size_t)c * 16384, 16384);
                    }
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            const uint32_t ring = tc::smem_u32(smem);
            const uint32_t idesc_half = tc::make_idesc(tc::FMT_F16, 128, 128), idesc_full = tc::make_idesc(tc::FMT_F16, 128, 256);
            uint32_t stage = 0, phase = 0, a_par = 0;
            auto next = [&]() {
                const uint32_t addr = ring + stage * HT_STAGE_BYTES;
                tc::mbar_wait(&bar.full[stage], phase);
                return addr;
            };
            auto commit_empty = [&](uint32_t st_idx) {
                if (MC) umma_commit_multicast(&bar.empty[st_idx], (uint16_t)3); else tc::umma_commit(&bar.empty[st_idx]);
            };
            auto release = [&]() {
                commit_empty(stage);
                if (++stage == HT_STAGES) { stage = 0; phase ^= 1u; }
            };
            for (int t = 0; t < n_my_tiles; ++t)
                for (int s = 0; s < prog.n_steps; ++s) {
                    const HtStep st = prog.step[s];
                    const uint32_t d = tmem + st.acc_col;
                    if (!st.no_wait) {
                        tc::mbar_wait(&bar.a_ready, a_par);
                        a_par ^= 1u;
                        tc::tc_fence_after_sync();
                    }
                    if (st.feature) {
                        for (int kb = 0; kb < st.kblocks; ++kb) {
                            // three stages per k-block, all released by commits behind the k-block's last MMA
                            const uint32_t a_addr = next();
                            const uint32_t st_a = stage;
                            if (++stage == HT_STAGES) { stage = 0; phase ^= 1u; }
                            const uint32_t bh_addr = next();
                            const uint32_t st_bh = stage;
                            if (++stage == HT_STAGES) { stage = 0; phase ^= 1u; }
                            const uint32_t bl_addr = next();
                            const uint32_t st_bl = stage;
                            if (++stage == HT_STAGES) { stage = 0; phase ^= 1u; }
                            tc::tc_fence_after_sync();
                            const uint64_t dAh = tc::make_smem_desc_sw128(a_addr), dAl = tc::make_smem_desc_sw128(a_addr + 16384);
                            const uint64_t dBh = tc::make_smem_desc_sw128(bh_addr), dBl = tc::make_smem_desc_sw128(bl_addr);
#pragma unroll
                            for (int k = 0; k < 4; ++k) {
                                tc::umma_f16(d, dAl + 2 * k, dBh + 2 * k, idesc_full, (kb | k | st.acc_in) != 0);
                                tc::umma_f16(d, dAh + 2 * k, dBh + 2 * k, idesc_full, 1);
                                tc::umma_f16(d, dAh + 2 * k, dBl + 2 * k, idesc_full, 1);
                            }
                            commit_empty(st_a);
                            commit_empty(st_bh);
                            commit_empty(st_bl);
                        }
                    } else {
                        for (int kb = 0; kb < st.kblocks; ++kb) {
                            const uint32_t ah = tmem + HT_A_HI + (uint32_t)kb * 32, al = tmem + HT_A_LO + (uint32_t)kb * 32;
                            uint64_t dB = tc::make_smem_desc_sw128(next());
                            tc::tc_fence_after_sync();
#pragma unroll
                            for (int k = 0; k < 4; ++k) {
                                tc::umma_f16_ts(d, al + 8 * k, dB + 2 * k, idesc_half, (kb | k | st.acc_in) != 0);
                                tc::umma_f16_ts(d, ah + 8 * k, dB + 2 * k, idesc_half, 1);
                            }
                            release();
                            dB = tc::make_smem_desc_sw128(next());
                            tc::tc_fence_after_sync();
#pragma unroll
                            for (int k = 0; k < 4; ++k) tc::umma_f16_ts(d, ah + 8 * k, dB + 2 * k, idesc_half, 1);
                            release();
                        }
                    }
                    if (st.commit == 2) {
                        tc::umma_commit(&bar.acc_full[0]);
                        tc::umma_commit(&bar.acc_full[1]);
                    } else if (st.commit == 1) {
                        tc::umma_commit(&bar.acc_full[st.acc_col ? 1 : 0]);
                    }
                }
        }
    } else {
        const int row = (warp & 3) * 32 + lane, cg = (warp - 2) >> 2;
        const uint32_t lane_base = (uint32_t)(row & ~31) << 16;
        uint32_t acc_par[2] = {0u, 0u};
        auto publish = [&]() {
            tc::tmem_st_wait();
            tc::tc_fence_before_sync();
            tc::mbar_arrive(&bar.a_ready);
        };
        auto wait_acc = [&](int hf) {
            tc::mbar_wait(&bar.acc_full[hf], acc_par[hf]);
            acc_par[hf] ^= 1u;
            tc::tc_fence_after_sync();
        };
        auto store_a = [&](int col0, const uint32_t* hi, const uint32_t* lo) {
            const uint32_t c = (uint32_t)(col0 >> 1);
            tc::tmem_st_32x32b_x8(tmem + lane_base + HT_A_HI + c, hi);
            tc::tmem_st_32x32b_x8(tmem + lane_base + HT_A_HI + c + 8, hi + 8);
            tc::tmem_st_32x32b_x8(tmem + lane_base + HT_A_LO + c, lo);
            tc::tmem_st_32x32b_x8(tmem + lane_base + HT_A_LO + c + 8, lo + 8);
        };
        for (int t = 0; t < n_my_tiles; ++t) {
            const int64_t tile = tile_of(t);
            const bool real = tile < p.n_tiles;              // MC: the odd tile's partner walks a dummy
            const int64_t gp = tile * TILE_M + row;
            const bool live = real && gp < p.n;
            // 8 columns [c, c + 8) of layer l: EM / EML stash chunks + fp16 hi / lo words of the next A operand
            auto emit8 = [&](int l, int c, const float* h, const float* em, uint32_t* hi4, uint32_t* lo4) {
                if (stash && real) {
                    const uint32_t off = t16_off(row, c >> 3);
                    uint4 q, ql;
                    split2_lo16(em[0], em[1], q.x, ql.x); split2_lo16(em[2], em[3], q.y, ql.y);
                    split2_lo16(em[4], em[5], q.z, ql.z); split2_lo16(em[6], em[7], q.w, ql.w);
                    stg16(p.EM[l] + (size_t)tile * T16_TILE_BYTES + off, q);
                    stg16(p.EML[l] + (size_t)tile * T16_TILE_BYTES + off, ql);
                }
#pragma unroll
                for (int i = 0; i < 4; ++i) split2_lo16(h[2 * i], h[2 * i + 1], hi4[i], lo4[i]);
            };
            // softplus of accumulator columns [col0, +32) of layer l (+ the feature part of the skip layer's pre-activation)
            auto act_half = [&](int l, int col0, uint32_t* hh, uint32_t* hl, float& head) {
                const float* __restrict__ bias = p.bias[l];
                float v[32];
                acc_load32(tmem, row, col0, v);
                if (l == 4 && live && !p.F16) {
                    const float* __restrict__ z = p.ZF4 + gp * 256 + col0;
#pragma unroll
                    for (int j = 0; j < 32; j += 4) {
                        const float4 a = ld4(z + j);
                        v[j] += a.x; v[j + 1] += a.y; v[j + 2] += a.z; v[j + 3] += a.w;
                    }
                }
#pragma unroll
                for (int j = 0; j < 32; j += 8) {
                    const float4 b0 = __ldg(reinterpret_cast<const float4*>(bias + col0 + j));
                    const float4 b1 = __ldg(reinterpret_cast<const float4*>(bias + col0 + j + 4));
                    const float bb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
                    float h[8], em[8];
#pragma unroll
                    for (int i = 0; i < 8; ++i) h[i] = softplus100_em(v[j + i] + bb[i], em[i]);
                    emit8(l, col0 + j, h, em, hh + (j >> 1), hl + (j >> 1));
                    if (l == 7) {
                        const float4 w0 = __ldg(reinterpret_cast<const float4*>(p.w_out0 + col0 + j));
                        const float4 w1 = __ldg(reinterpret_cast<const float4*>(p.w_out0 + col0 + j + 4));
                        head += h[0] * w0.x + h[1] * w0.y + h[2] * w0.z + h[3] * w0.w + h[4] * w1.x + h[5] * w1.y + h[6] * w1.z + h[7] * w1.w;
                    }
                }
            };
            // ---- layer 0: with the feature tiles it is an MMA step like the others (this arrival only tells the issuer that the
            //      previous tile's accumulator has been read); legacy inputs: its activation rows were computed before this
            //      kernel and become the first A operand here
#pragma unroll 1
            for (int hf = 0; hf < (p.F16 ? 0 : 2); ++hf) {
                const int col0 = 128 * hf + cg * 32;
                uint32_t hh[16], hl[16];
                const float* __restrict__ hrow = p.H0 + gp * 256 + col0;
#pragma unroll
                for (int j = 0; j < 32; j += 8) {
                    float h[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f}, em[8];
                    if (live) {
                        const float4 a = ld4(hrow + j), b = ld4(hrow + j + 4);
                        h[0] = a.x; h[1] = a.y; h[2] = a.z; h[3] = a.w; h[4] = b.x; h[5] = b.y; h[6] = b.z; h[7] = b.w;
                    }
#pragma unroll
                    for (int i = 0; i < 8; ++i) em[i] = ex2_approx(-h[i] * 144.26950408889634f);      // exp(-100 h)
                    emit8(0, col0 + j, h, em, hh + (j >> 1), hl + (j >> 1));
                }
                store_a(col0, hh, hl);
            }
            publish();
            float head = 0.0f;
            for (int l = p.F16 ? 0 : 1; l < 8; ++l) {
                uint32_t hh[16], hl[16];
                // first half (columns cg*32 ..) under the second half's MMAs
                wait_acc(0);
                act_half(l, cg * 32, hh, hl, head);
                // second half: every MMA of the layer has read A, it may be overwritten
                wait_acc(1);
                const bool feeds = l < 7 || stash;      // sdf only: layer 7 feeds no further MMA step, nothing to publish
                if (feeds) store_a(cg * 32, hh, hl);
                act_half(l, 128 + cg * 32, hh, hl, head);
                if (feeds) {
                    store_a(128 + cg * 32, hh, hl);
                    publish();
                }
            }
            // sdf = h7 . W_out[0] + b_out[0]
            s_head[cg * TILE_M + row] = head;
            tc::named_bar_sync(1, EPI_THREADS);
            if (cg == 0 && live) {
                float acc = 0.0f;
#pragma unroll
                for (int g = 0; g < EPI_CGROUPS; ++g) acc += s_head[g * TILE_M + row];
                p.sdf[gp] = acc + __ldg(p.bias[8]);
            }
            if (stash) {
                // feature head: rows 1..256 of the output layer, no activation
                for (int hf = 0; hf < 2; ++hf) {
                    wait_acc(hf);
                    const int col0 = 128 * hf + cg * 32;
                    float v[32];
                    acc_load32(tmem, row, col0, v);
                    if (live) {
                        const float* __restrict__ bias = p.bias[8] + 1;
                        float* __restrict__ fr = p.feat + gp * p.ld_feat + col0;
#pragma unroll
                        for (int j = 0; j < 32; j += 4)
                            st4(fr + j, make_float4(v[j] + __ldg(bias + col0 + j), v[j + 1] + __ldg(bias + col0 + j + 1),
                                                    v[j + 2] + __ldg(bias + col0 + j + 2), v[j + 3] + __ldg(bias + col0 + j + 3)));
                    }
                }
            }
            tc::tc_fence_before_sync();       // the accumulator reads above precede the next tile's MMAs (ordered by its a_ready)
        }
    }
    tc::tc_fence_before_sync();
    __syncthreads();
    if (MC) cluster_sync_all();          // no CTA leaves while its peer may still address its shared memory
    if (warp == 0) tc::tmem_dealloc(tmem, 512);
}

// ------------------------------------------------------------------------------------------------------------------
// normal sweep: D_7 = s'(h_7) W_out[0]; D_{l-1} = s'(h_{l-1}) * (D_l W_l), l = 7..1
// ------------------------------------------------------------------------------------------------------------------
struct HandNsweepParams {
    uint32_t* dbg;
    int64_t n;
    uint8_t* D16[8];         // NULL entries: nothing will be differentiated (rendering), the tiles are not written
    float* FB;               // column-major tiles [tile][1388][128] fp32: D_4 W_4[:, 256:] + D_0 W_0
    const uint8_t* chain;
    const float* w_out0;
    int n_tiles;
};

// One feature-side chunk step: accumulator columns [0, n) -> out[256 ch + c][row] of this tile's column-major
// [1388 columns][128 points] fp32 block (a warp's 32 rows of one column are one 128-byte line; the thread-per-point HALO
// kernels of fields_hand.cu read it the same way), or added to what the same thread stored there before
template <bool ADD>
__device__ __forceinline__ void hand_f_chunk(uint32_t tmem, uint32_t lane_base, int row, int cg, int ch, float* __restrict__ otile) {
    const int n_mma = hand_f_chunk_n(ch);
#pragma unroll
    for (int sb = 0; sb < 4; ++sb) {
        const int c = 128 * (sb >> 1) + cg * 32 + 16 * (sb & 1);
        if (c >= n_mma) continue;
        float* __restrict__ o = otile + (size_t)(256 * ch + c) * TILE_M + row;
        const int valid = 1386 - (256 * ch + c);           // columns of this sub-block inside the 1386 features (>= 16: all)
        float old[16];
        if (ADD) {
            // all sixteen loads in flight together, and before the accumulator load: written as `o[..] += g` the compiler
            // chains load -> add -> store sixteen times (one L2 round trip each)
#pragma unroll
            for (int j = 0; j < 16; ++j) old[j] = j < valid ? __ldcg(o + j * TILE_M) : 0.0f;
        }
        float g[16];
        sw_ld16(tmem, lane_base, c, g);
#pragma unroll
        for (int j = 0; j < 16; ++j)
            if (j < valid) __stcg(o + j * TILE_M, ADD ? old[j] + g[j] : g[j]);
    }
}

__global__ void __launch_bounds__(SW_THREADS, 1)
hand_nsweep16_kernel(const __grid_constant__ HandNsweepParams p, const __grid_constant__ SwProgram prog,
                     const __grid_constant__ SwInputs inputs) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    __shared__ SwBarriers bar;
    uint8_t* smem = sw_setup(smem_raw, &bar);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int n_my_tiles = p.n_tiles > (int)blockIdx.x ? (p.n_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x : 0;
    if (warp == 0) {
        if (lane == 0) sw_producer(prog, p.chain, smem, &bar, n_my_tiles);
    } else if (warp == 1) {
        if (lane == 0) sw_mma(prog, smem, &bar, n_my_tiles, p.dbg);
    } else if (warp == 2 + EPI_WARPS) {
        if (lane == 0) sw_in_producer(inputs, smem, &bar, n_my_tiles);
    } else {
        const int row = (warp & 3) * 32 + lane, cg = (warp - 2) >> 2;
        const uint32_t lane_base = (uint32_t)(row & ~31) << 16;
        const uint32_t tmem = bar.tmem_base;
        uint32_t acc_par = 0, in_par[2] = {0u, 0u};
        for (int t = 0; t < n_my_tiles; ++t) {
            const int64_t tile = (int64_t)blockIdx.x + (int64_t)t * gridDim.x;
            const int64_t gp = tile * TILE_M + row;
            const bool live = gp < p.n;
            const size_t tb = (size_t)tile * T16_TILE_BYTES;
            // 16 columns [c, c + 16) of D_lyr -> fp16 hi / lo halves of the next A operand, bf16 chunks of D16, fp32 rows
            auto emit16 = [&](int lyr, int c, const float* d) {
                {
                    uint32_t hi8[8], lo8[8];
#pragma unroll
                    for (int i = 0; i < 8; ++i) split2_lo16(d[2 * i], d[2 * i + 1], hi8[i], lo8[i]);
                    sw_st16(tmem, lane_base, SW_A0, c, hi8);
                    sw_st16(tmem, lane_base, SW_A1, c, lo8);
                }
                if (p.D16[lyr]) {
                    uint4 q0, q1;
                    q0.x = pack_bf16x2(d[0], d[1]); q0.y = pack_bf16x2(d[2], d[3]); q0.z = pack_bf16x2(d[4], d[5]); q0.w = pack_bf16x2(d[6], d[7]);
                    q1.x = pack_bf16x2(d[8], d[9]); q1.y = pack_bf16x2(d[10], d[11]); q1.z = pack_bf16x2(d[12], d[13]); q1.w = pack_bf16x2(d[14], d[15]);
                    stg16(p.D16[lyr] + tb + t16_off(row, c >> 3), q0);
                    stg16(p.D16[lyr] + tb + t16_off(row, (c >> 3) + 1), q1);
                }
            };
            float* __restrict__ fb_tile = p.FB + (size_t)tile * (1388 * TILE_M);
            auto load_sp16 = [&](int hf, int c, float* sp) {       // s' = 1 - (em_hi + em_lo) of 16 columns, from the input slot
                const uint4 a = sw_in_ld(smem, hf, 0, row, c), b = sw_in_ld(smem, hf, 0, row, c + 8);
                const uint4 al = sw_in_ld(smem, hf, 1, row, c), bl = sw_in_ld(smem, hf, 1, row, c + 8);
                const uint32_t w[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
                const uint32_t wl[8] = {al.x, al.y, al.z, al.w, bl.x, bl.y, bl.z, bl.w};
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const float2 f = f16x2_unpack(w[i]), fl = f16x2_unpack(wl[i]);
                    sp[2 * i] = 1.0f - (f.x + fl.x);
                    sp[2 * i + 1] = 1.0f - (f.y + fl.y);
                }
            };
            // ---- seed: D_7 = s'(h_7) * W_out[0] (the previous tile's MMAs are complete: A may be written) --------------------
#pragma unroll 1
            for (int hf = 0; hf < 2; ++hf) {
                sw_in_wait(&bar, hf, in_par);
#pragma unroll
                for (int sub = 0; sub < 2; ++sub) {
                    const int c = 128 * hf + cg * 32 + 16 * sub;
                    float sp[16], d[16];
                    load_sp16(hf, c, sp);
#pragma unroll
                    for (int j = 0; j < 16; j += 4) {
                        const float4 w = __ldg(reinterpret_cast<const float4*>(p.w_out0 + c + j));
                        d[j] = sp[j] * w.x; d[j + 1] = sp[j + 1] * w.y; d[j + 2] = sp[j + 2] * w.z; d[j + 3] = sp[j + 3] * w.w;
                    }
                    if (!live) {
#pragma unroll
                        for (int j = 0; j < 16; ++j) d[j] = 0.0f;
                    }
                    emit16(7, c, d);
                }
                sw_in_release(&bar, hf);
            }
            sw_publish(&bar);
            for (int l = 7; l >= 1; --l) {
                sw_wait_acc(&bar, acc_par);
                // four 16-column sub-blocks; the accumulator load of the next one is in flight while this one is processed
                float accv[2][16];
                sw_ld16_nowait(tmem, lane_base, cg * 32, accv[0]);
#pragma unroll
                for (int sb = 0; sb < 4; ++sb) {
                    const int hf = sb >> 1;
                    const int c = 128 * hf + cg * 32 + 16 * (sb & 1);
                    if ((sb & 1) == 0) sw_in_wait(&bar, hf, in_par);
                    float sp[16];
                    load_sp16(hf, c, sp);
                    tc::tmem_ld_wait();
                    if (sb < 3) sw_ld16_nowait(tmem, lane_base, 128 * ((sb + 1) >> 1) + cg * 32 + 16 * ((sb + 1) & 1), accv[(sb + 1) & 1]);
                    float* g = accv[sb & 1];
#pragma unroll
                    for (int j = 0; j < 16; ++j) g[j] = live ? sp[j] * g[j] : 0.0f;
                    emit16(l - 1, c, g);
                    if (sb & 1) sw_in_release(&bar, hf);
                }
                sw_publish(&bar);
                if (l == 5) {
                    // D_4 is the A operand: FB = D_4 W_4[:, 256:], six chunk steps, before the layer-4 step overwrites nothing
                    // (A stays) -- every arrival below releases the accumulator for the next step
#pragma unroll 1
                    for (int ch = 0; ch < HAND_F_CHUNKS; ++ch) {
                        sw_wait_acc(&bar, acc_par);
                        hand_f_chunk<false>(tmem, lane_base, row, cg, ch, fb_tile);
                        sw_publish(&bar);
                    }
                }
            }
            // D_0 is the A operand: FB += D_0 W_0 (the rows this thread wrote above)
#pragma unroll 1
            for (int ch = 0; ch < HAND_F_CHUNKS; ++ch) {
                sw_wait_acc(&bar, acc_par);
                hand_f_chunk<true>(tmem, lane_base, row, cg, ch, fb_tile);
                if (ch + 1 < HAND_F_CHUNKS) sw_publish(&bar);      // one arrival per MMA step: the last one feeds none
            }
            tc::tc_fence_before_sync();
        }
    }
    sw_teardown(&bar);
}

// ------------------------------------------------------------------------------------------------------------------
// second-order backward: tangent sweep + reverse sweep over layers 0..8 (the 1386-wide ends outside)
// ------------------------------------------------------------------------------------------------------------------
struct HandBwdParams {
    uint32_t* dbg;
    int64_t n;
    const float* Q0;         // [np, 256] fp32 rows: tF W_0^T
    const float* QF4;        // [np, 256] fp32 rows: tF W_4[:, 256:]^T
    const float* d_sdf;      // may be NULL
    const float* d_feat;     // may be NULL
    int64_t ld_dfeat;
    uint8_t* X16[8];
    float* DF;               // column-major tiles [tile][1388][128] fp32: dz_4 W_4[:, 256:] + dz_0 W_0
    const uint8_t* chain;
    const float* w_out0;
    int n_tiles;
};

__global__ void __launch_bounds__(SW_THREADS, 1)
hand_bwd16_kernel(const __grid_constant__ HandBwdParams p, const __grid_constant__ SwProgram prog,
                  const __grid_constant__ SwInputs inputs) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    __shared__ SwBarriers bar;
    uint8_t* smem = sw_setup(smem_raw, &bar);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int n_my_tiles = p.n_tiles > (int)blockIdx.x ? (p.n_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x : 0;
    if (warp == 0) {
        if (lane == 0) sw_producer(prog, p.chain, smem, &bar, n_my_tiles);
    } else if (warp == 1) {
        if (lane == 0) sw_mma(prog, smem, &bar, n_my_tiles, p.dbg);
    } else if (warp == 2 + EPI_WARPS) {
        if (lane == 0) sw_in_producer(inputs, smem, &bar, n_my_tiles);
    } else {
        const int row = (warp & 3) * 32 + lane, cg = (warp - 2) >> 2;
        const uint32_t lane_base = (uint32_t)(row & ~31) << 16;
        const uint32_t tmem = bar.tmem_base;
        uint32_t acc_par = 0, in_par[2] = {0u, 0u};
        for (int t = 0; t < n_my_tiles; ++t) {
            const int64_t tile = (int64_t)blockIdx.x + (int64_t)t * gridDim.x;
            const int64_t gp = tile * TILE_M + row;
            const bool live = gp < p.n;
            const size_t tb = (size_t)tile * T16_TILE_BYTES;
            const float gs = (live && p.d_sdf) ? p.d_sdf[gp] : 0.0f;
            float* __restrict__ df_tile = p.DF + (size_t)tile * (1388 * TILE_M);
            auto pack16 = [&](const float* v, uint32_t* w) {
#pragma unroll
                for (int i = 0; i < 8; ++i) w[i] = pack_bf16x2(v[2 * i], v[2 * i + 1]);
            };
            auto store16 = [&](uint8_t* arr, int c, const uint32_t* w) {
                stg16(arr + tb + t16_off(row, c >> 3), make_uint4(w[0], w[1], w[2], w[3]));
                stg16(arr + tb + t16_off(row, (c >> 3) + 1), make_uint4(w[4], w[5], w[6], w[7]));
            };
            // 16 values -> bf16 hi / lo halves of the next A operand (tensor memory)
            auto emit_a16 = [&](int c, const float* v) {
                uint32_t hi8[8], lo8[8];
#pragma unroll
                for (int i = 0; i < 8; ++i) split2(v[2 * i], v[2 * i + 1], hi8[i], lo8[i]);
                sw_st16(tmem, lane_base, SW_A0, c, hi8);
                sw_st16(tmem, lane_base, SW_A1, c, lo8);
            };
            auto unpack_bf16 = [&](const uint4& a, const uint4& b, float* v) {
                const uint32_t w[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
#pragma unroll
                for (int i = 0; i < 8; ++i) { v[2 * i] = bf16_lo(w[i]); v[2 * i + 1] = bf16_hi(w[i]); }
            };
            auto unpack_f16 = [&](const uint4& a, const uint4& b, float* v) {
                const uint32_t w[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const float2 f = f16x2_unpack(w[i]);
                    v[2 * i] = f.x;
                    v[2 * i + 1] = f.y;
                }
            };
            auto add_row16 = [&](const float* __restrict__ src, int c, float* q) {
                if (!live) return;
                const float* __restrict__ r = src + gp * 256 + c;
#pragma unroll
                for (int j = 0; j < 16; j += 4) {
                    const float4 a = ld4(r + j);
                    q[j] += a.x; q[j + 1] += a.y; q[j + 2] += a.z; q[j + 3] += a.w;
                }
            };
            // ---- tangent sweep: q_l = W_l u_{l-1} (q_0, and the feature part of q_4, from the contractions that ran before this
            //      kernel); u_l = s'(h_l) q_l; X_l = 100 (1 - s') D_l q_l -----------------------------------------------------
            for (int l = 0; l < 8; ++l) {
                if (l > 0) sw_wait_acc(&bar, acc_par);
#pragma unroll 1
                for (int hf = 0; hf < 2; ++hf) {
                    sw_in_wait(&bar, hf, in_par);
#pragma unroll
                    for (int sub = 0; sub < 2; ++sub) {
                        const int c = 128 * hf + cg * 32 + 16 * sub;
                        float q[16], em[16], d[16];
                        unpack_f16(sw_in_ld(smem, hf, 0, row, c), sw_in_ld(smem, hf, 0, row, c + 8), em);
                        unpack_bf16(sw_in_ld(smem, hf, 1, row, c), sw_in_ld(smem, hf, 1, row, c + 8), d);
                        if (l > 0) {
                            sw_ld16(tmem, lane_base, c, q);
                        } else {
#pragma unroll
                            for (int j = 0; j < 16; ++j) q[j] = 0.0f;
                            add_row16(p.Q0, c, q);
                        }
                        if (l == 4) add_row16(p.QF4, c, q);
#pragma unroll
                        for (int j = 0; j < 16; ++j) {
                            const float u = (1.0f - em[j]) * q[j];
                            d[j] = 100.0f * em[j] * d[j] * q[j];
                            q[j] = u;
                        }
                        if (l < 7) emit_a16(c, q);
                        uint32_t w[8];
                        pack16(d, w);
                        store16(p.X16[l], c, w);
                    }
                    sw_in_release(&bar, hf);
                }
                if (l == 7) {
                    // A operand of the output layer's reverse step: the point's row of d_feat
#pragma unroll 1
                    for (int sb = 0; sb < 4; ++sb) {
                        const int c = 128 * (sb >> 1) + cg * 32 + 16 * (sb & 1);
                        float v[16];
#pragma unroll
                        for (int j = 0; j < 16; j += 4) {
                            float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
                            if (live && p.d_feat) a = ld4(p.d_feat + gp * p.ld_dfeat + c + j);
                            v[j] = a.x; v[j + 1] = a.y; v[j + 2] = a.z; v[j + 3] = a.w;
                        }
                        emit_a16(c, v);
                    }
                }
                sw_publish(&bar);
            }
            // ---- reverse sweep: dz_{l-1} = s'(h_{l-1}) (dz_l W_l) + X_{l-1}, l = 8..1 -------------------------------------------
            for (int l = 8; l >= 1; --l) {
                // X_{l-1} was written by THIS thread during the tangent sweep: plain loads, issued before the wait for the MMAs
                const uint8_t* xp = p.X16[l - 1] + tb;
                uint4 xq[8];
#pragma unroll
                for (int i = 0; i < 8; ++i) xq[i] = ldg16(xp + t16_off(row, ((128 * (i >> 2) + cg * 32) >> 3) + (i & 3)));
                sw_wait_acc(&bar, acc_par);
#pragma unroll
                for (int hf = 0; hf < 2; ++hf) {
                    sw_in_wait(&bar, hf, in_par);
#pragma unroll
                    for (int sub = 0; sub < 2; ++sub) {
                        const int c = 128 * hf + cg * 32 + 16 * sub;
                        float da[16], em[16], x[16];
                        unpack_f16(sw_in_ld(smem, hf, 0, row, c), sw_in_ld(smem, hf, 0, row, c + 8), em);
                        unpack_bf16(xq[4 * hf + 2 * sub], xq[4 * hf + 2 * sub + 1], x);
                        sw_ld16(tmem, lane_base, c, da);
                        if (l == 8) {
#pragma unroll
                            for (int j = 0; j < 16; j += 4) {
                                const float4 w = __ldg(reinterpret_cast<const float4*>(p.w_out0 + c + j));
                                da[j] += gs * w.x; da[j + 1] += gs * w.y; da[j + 2] += gs * w.z; da[j + 3] += gs * w.w;
                            }
                        }
#pragma unroll
                        for (int j = 0; j < 16; ++j) da[j] = live ? fmaf(1.0f - em[j], da[j], x[j]) : 0.0f;
                        emit_a16(c, da);
                    }
                    sw_in_release(&bar, hf);
                }
                sw_publish(&bar);
                if (l == 5) {
                    // dz_4 is the A operand: DF = d_xyz_feature + dz_4 W_4[:, 256:] in six chunk steps
#pragma unroll 1
                    for (int ch = 0; ch < HAND_F_CHUNKS; ++ch) {
                        sw_wait_acc(&bar, acc_par);
                        hand_f_chunk<false>(tmem, lane_base, row, cg, ch, df_tile);
                        sw_publish(&bar);
                    }
                }
            }
            // dz_0 is the A operand: DF += dz_0 W_0
#pragma unroll 1
            for (int ch = 0; ch < HAND_F_CHUNKS; ++ch) {
                sw_wait_acc(&bar, acc_par);
                hand_f_chunk<true>(tmem, lane_base, row, cg, ch, df_tile);
                if (ch + 1 < HAND_F_CHUNKS) sw_publish(&bar);      // one arrival per MMA step: the last one feeds none
            }
            tc::tc_fence_before_sync();
        }
    }
    sw_teardown(&bar);
}

// ------------------------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------------------------
// stash (floats per padded point): HROW 1644 | FB 1388 | RA 256 | RB 256 | EM[8] | EML[8] | D16[8] (128 floats each) | F16 1408
//   RA: H0, RB: ZF4 (legacy inputs of the trunk); F16: the fp16 pair tiles of the HALO feature (22 k-blocks x 64 x 2 x 2 bytes)
int64_t hand16_stash_floats(int64_t n) { return round_up(n, TILE_M) * (1644 + 1388 + 512 + 24 * 128 + 1408); }
// backward workspace: AU4 1644 | DF 1388 | Q0 | QF4 (256 each) | X16[8]
int64_t hand16_bwd_ws_floats(int64_t n) { return round_up(n, TILE_M) * (1644 + 1388 + 512 + 8 * 128); }

static void sw_layer(SwProgram& prog, int& k, uint32_t off, int n_mma, int kblocks, int f16) {
    SwStep& st = prog.step[k++];
    st.b_off = off;
    st.n_mma = (uint16_t)n_mma;
    st.kblocks = (uint8_t)kblocks;
    st.f16 = (uint8_t)f16;
    st.acc_in = 0;
}

int launch_hand16_trunk(const hn_mlp_t* m, const uint8_t* ops, int64_t n, const uint8_t* F16, const float* H0, const float* ZF4,
                        float* sdf, float* feat, int64_t ld_feat, uint8_t* const* EM, uint8_t* const* EML, cudaStream_t s) {
    const HandLayout L = hand_layout();
    const int n_tiles = (int)ceil_div(n, TILE_M);
    static bool configured = false;
    static int mc_clusters = 0;      // > 0: pairs of CTAs with multicast weight stages (HONERF_HAND_MC=0 turns it off)
    if (!configured) {
        HN_CHECK_CUDA(cudaFuncSetAttribute(hand_trunk16_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, HT_SMEM_BYTES));
        HN_CHECK_CUDA(cudaFuncSetAttribute(hand_trunk16_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, HT_SMEM_BYTES));
        const bool want = getenv("HONERF_HAND_MC") ? atoi(getenv("HONERF_HAND_MC")) != 0 : true;
        if (want) {
            cudaLaunchConfig_t cfg = {};
            cfg.gridDim = dim3((unsigned)(sm_count() & ~1)); cfg.blockDim = dim3(THREADS); cfg.dynamicSmemBytes = HT_SMEM_BYTES;
            cudaLaunchAttribute at;
            at.id = cudaLaunchAttributeClusterDimension;
            at.val.clusterDim.x = 2; at.val.clusterDim.y = 1; at.val.clusterDim.z = 1;
            cfg.attrs = &at; cfg.numAttrs = 1;
            int nc = 0;
            if (cudaOccupancyMaxActiveClusters(&nc, hand_trunk16_kernel<true>, &cfg) == cudaSuccess && nc > 0) mc_clusters = nc;
            else (void)cudaGetLastError();
        }
        configured = true;
    }
    HandTrunkParams p;
    p.n = n; p.F16 = F16; p.H0 = H0; p.ZF4 = ZF4; p.sdf = sdf; p.feat = feat; p.ld_feat = ld_feat;
    for (int l = 0; l < 8; ++l) { p.EM[l] = EM ? EM[l] : nullptr; p.EML[l] = EML ? EML[l] : nullptr; }
    p.chain = ops;
    for (int l = 0; l < 9; ++l) p.bias[l] = m->b[l];
    p.w_out0 = m->W[8];
    p.n_tiles = n_tiles;
    HtProgram prog = {};
    int k = 0;
    auto feature_step = [&](int which, bool first) {
        HtStep& st = prog.step[k++];
        st.b_off = L.ntf16_off[which]; st.acc_col = 0; st.kblocks = HAND_F_KBLOCKS;
        st.feature = 1; st.no_wait = first ? 0 : 1; st.acc_in = first ? 0 : 1; st.commit = 2;
    };
    if (F16) feature_step(0, true);                                   // layer 0: F @ W_0^T
    for (int l = 1; l <= (feat ? 8 : 7); ++l) {
        for (int h = 0; h < 2; ++h) {
            HtStep& st = prog.step[k++];
            st.b_off = L.nth_off[l][h]; st.acc_col = (uint16_t)(128 * h); st.kblocks = 4;
            st.feature = 0; st.no_wait = (uint8_t)h; st.acc_in = 0;
            st.commit = (F16 && l == 4) ? 0 : 1;                      // the skip layer's accumulator is complete after its feature part
        }
        if (F16 && l == 4) feature_step(1, false);                    // + F @ W_4[:, 256:]^T onto both halves
    }
    prog.n_steps = k;
    if (F16 && mc_clusters > 0 && n_tiles >= 2) {
        const int clusters = std::min((n_tiles + 1) / 2, mc_clusters);
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3((unsigned)(2 * clusters)); cfg.blockDim = dim3(THREADS); cfg.dynamicSmemBytes = HT_SMEM_BYTES; cfg.stream = s;
        cudaLaunchAttribute at;
        at.id = cudaLaunchAttributeClusterDimension;
        at.val.clusterDim.x = 2; at.val.clusterDim.y = 1; at.val.clusterDim.z = 1;
        cfg.attrs = &at; cfg.numAttrs = 1;
        HN_CHECK_CUDA(cudaLaunchKernelEx(&cfg, hand_trunk16_kernel<true>, p, prog));
    } else {
        hand_trunk16_kernel<false><<<std::min(n_tiles, sm_count()), THREADS, HT_SMEM_BYTES, s>>>(p, prog);
    }
    count_launch();
    HN_CHECK_LAUNCH();
    return HN_OK;
}

int launch_hand16_nsweep(const hn_mlp_t* m, const uint8_t* ops, int64_t n, uint8_t* const* EM, uint8_t* const* EML,
                         uint8_t* const* D16, float* FB, cudaStream_t s) {
    const HandLayout L = hand_layout();
    const int n_tiles = (int)ceil_div(n, TILE_M);
    static bool configured = false;
    if (!configured) {
        HN_CHECK_CUDA(cudaFuncSetAttribute(hand_nsweep16_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SW_SMEM_BYTES));
        configured = true;
    }
    HandNsweepParams p;
    p.n = n; p.FB = FB; p.dbg = dbg_slot(1);
    for (int l = 0; l < 8; ++l) p.D16[l] = D16 ? D16[l] : nullptr;
    p.chain = ops;
    p.w_out0 = m->W[8];
    p.n_tiles = n_tiles;
    SwInputs in = {};
    int ne = 0;
    for (int hf = 0; hf < 2; ++hf) in.ev[ne++] = SwInEvent{EM[7], EML[7]};                  // seed
    for (int l = 7; l >= 1; --l)
        for (int hf = 0; hf < 2; ++hf) in.ev[ne++] = SwInEvent{EM[l - 1], EML[l - 1]};
    in.n_events = ne;
    SwProgram prog = {};
    int k = 0;
    for (int l = 7; l >= 1; --l) {
        sw_layer(prog, k, L.nn16_off[l], 256, 4, 1);                                        // d @ W_l
        if (l == 5)
            for (int ch = 0; ch < HAND_F_CHUNKS; ++ch) sw_layer(prog, k, L.nnf16_off[0][ch], hand_f_chunk_n(ch), 4, 1);   // D_4 @ W_4f
    }
    for (int ch = 0; ch < HAND_F_CHUNKS; ++ch) sw_layer(prog, k, L.nnf16_off[1][ch], hand_f_chunk_n(ch), 4, 1);           // D_0 @ W_0
    prog.n_steps = k;
    hand_nsweep16_kernel<<<std::min(n_tiles, sm_count()), SW_THREADS, SW_SMEM_BYTES, s>>>(p, prog, in);
    count_launch();
    HN_CHECK_LAUNCH();
    return HN_OK;
}

int launch_hand16_bwd(const hn_mlp_t* m, const uint8_t* ops, int64_t n, uint8_t* const* EM, uint8_t* const* D16,
                      uint8_t* const* X16, const float* Q0, const float* QF4, const float* d_sdf, const float* d_feat,
                      int64_t ld_dfeat, float* DF, cudaStream_t s) {
    HN_REQUIRE(!d_feat || (ld_dfeat % 4 == 0 && aligned16(d_feat)), "d_feat must be 16-byte aligned with ld %% 4 == 0");
    const HandLayout L = hand_layout();
    const int n_tiles = (int)ceil_div(n, TILE_M);
    static bool configured = false;
    if (!configured) {
        HN_CHECK_CUDA(cudaFuncSetAttribute(hand_bwd16_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SW_SMEM_BYTES));
        configured = true;
    }
    HandBwdParams p;
    p.n = n; p.Q0 = Q0; p.QF4 = QF4; p.d_sdf = d_sdf; p.d_feat = d_feat; p.ld_dfeat = ld_dfeat;
    p.DF = DF; p.dbg = dbg_slot(2);
    for (int l = 0; l < 8; ++l) p.X16[l] = X16[l];
    p.chain = ops;
    p.w_out0 = m->W[8];
    p.n_tiles = n_tiles;
    SwInputs in = {};
    int ne = 0;
    for (int l = 0; l < 8; ++l)
        for (int hf = 0; hf < 2; ++hf) in.ev[ne++] = SwInEvent{EM[l], D16[l]};              // tangent sweep
    for (int l = 8; l >= 1; --l)
        for (int hf = 0; hf < 2; ++hf) in.ev[ne++] = SwInEvent{EM[l - 1], nullptr};         // reverse sweep
    in.n_events = ne;
    SwProgram prog = {};
    int k = 0;
    for (int l = 1; l <= 7; ++l) sw_layer(prog, k, L.nt_off[l], 256, 4, 0);                 // tangent: u @ W_l^T
    for (int l = 8; l >= 1; --l) {
        sw_layer(prog, k, L.nn_off[l], 256, 4, 0);                                          // reverse: dz @ W_l
        if (l == 5)
            for (int ch = 0; ch < HAND_F_CHUNKS; ++ch) sw_layer(prog, k, L.nnf_off[0][ch], hand_f_chunk_n(ch), 4, 0);     // dz_4 @ W_4f
    }
    for (int ch = 0; ch < HAND_F_CHUNKS; ++ch) sw_layer(prog, k, L.nnf_off[1][ch], hand_f_chunk_n(ch), 4, 0);             // dz_0 @ W_0
    prog.n_steps = k;
    hand_bwd16_kernel<<<std::min(n_tiles, sm_count()), SW_THREADS, SW_SMEM_BYTES, s>>>(p, prog, in);
    count_launch();
    HN_CHECK_LAUNCH();
    return HN_OK;
}

int hand16_pack(const hn_mlp_t* m, uint8_t* dst, cudaStream_t s) {
    const HandLayout L = hand_layout();
    pack_batch_begin();
    for (int l = 1; l <= 8; ++l) {
        const int row0 = l == 8 ? 1 : 0;
        for (int h = 0; h < 2; ++h)
            HN_PROPAGATE(launch_pack_b(m->W[l], m->ld[l], pack_map(row0 + 128 * h, 0), 128, 256, 128, 4, dst + L.nth_off[l][h], s, true));
        if (l <= 7) {
            HN_PROPAGATE(launch_pack_b(m->WT[l], m->ldT[l], pack_map(0, 0), 256, 256, 256, 4, dst + L.nn16_off[l], s, true));
            HN_PROPAGATE(launch_pack_b(m->W[l], m->ld[l], pack_map(0, 0), 256, 256, 256, 4, dst + L.nt_off[l], s));
        }
        HN_PROPAGATE(launch_pack_b(m->WT[l], m->ldT[l], pack_map(0, row0), 256, 256, 256, 4, dst + L.nn_off[l], s));
    }
    // feature-side INPUT contractions of the value trunk: B(n = output, k = feature) from W_0 and W_4[:, 256:], fp16 pairs
    HN_PROPAGATE(launch_pack_b(m->W[0], m->ld[0], pack_map(0, 0), 256, 1386, 256, HAND_F_KBLOCKS, dst + L.ntf16_off[0], s, true));
    HN_PROPAGATE(launch_pack_b(m->W[4], m->ld[4], pack_map(0, 256), 256, 1386, 256, HAND_F_KBLOCKS, dst + L.ntf16_off[1], s, true));
    // feature-side chunks: rows = features of WT_4 (after the 256 h3 rows) / WT_0
    for (int w = 0; w < 2; ++w)
        for (int ch = 0; ch < HAND_F_CHUNKS; ++ch) {
            const int l = w == 0 ? 4 : 0, r0 = (w == 0 ? 256 : 0) + 256 * ch;
            HN_PROPAGATE(launch_pack_b(m->WT[l], m->ldT[l], pack_map(r0, 0), hand_f_chunk_valid(ch), 256, hand_f_chunk_n(ch), 4,
                                       dst + L.nnf16_off[w][ch], s, true));
            HN_PROPAGATE(launch_pack_b(m->WT[l], m->ldT[l], pack_map(r0, 0), hand_f_chunk_valid(ch), 256, hand_f_chunk_n(ch), 4,
                                       dst + L.nnf_off[w][ch], s));
        }
    return pack_batch_flush(s);
}

}  // namespace chain
}  // namespace hn
