// HN_TC_MIXED16: the object SDF field with a 16-bit activation stash and TMEM-resident operands.
//
// What changes against the HN_TC_BF16X3 chain kernels (chain_obj.cu), and why (profiles/r02_precision_table.md):
//  * every contraction of the forward and of the sweeps keeps three 16-bit MMAs per product (hi + lo operand pairs:
//    measured on the B200, two MMAs per product in the normal sweep move normals by 3e-4, which flips enough ReLUs of
//    the colour net to move ITS weight gradients by 3-4 %); what changes is the STASH: 16-bit, so the weight gradients
//    take their operands as stored (one bf16 MMA per product): normals 2e-5, weight gradients <= 4e-3 relative against
//    the 1e-2 bound;
//  * the A operand of every sweep layer lives in tensor memory (accumulator 256 + A_hi 128 + A_lo 128 = 512 columns,
//    written by the epilogue with tcgen05.st, consumed by the `ts` form of tcgen05.mma): shared memory holds nothing
//    but the weight ring and the input slots the stash tiles are bulk-copied into one half layer ahead of their use;
//  * everything the sweeps exchange through HBM is 16-bit and "dW-ready": [128 points x 256 features] tiles in the
//    un-swizzled MN-major core-matrix order the weight-gradient MMAs consume (8 points x 8 features = 128 contiguous
//    bytes), grouped so that a warp's access (32 consecutive points, one 8-feature chunk) is one 512-byte segment and a
//    64-point operand stage of the weight-gradient kernel is ONE 32 KB bulk copy with no conversion.
//    Stored: EM = exp(-100 h) = 1 - softplus'(z) as fp16 (the sweeps need no transcendental), A16 = h as bf16 (weight-
//    gradient operand), D16 / U16 / X16 / DZ16 as bf16: 28 bytes per element and layer through HBM instead of 56.
#pragma once
#include "chain_common.cuh"

namespace hn {
namespace chain {

// ---- 16-bit stash tiles ---------------------------------------------------------------------------------------------
// tile = [half (64 points)][chunk f8 (8 features, NCH per point)][64 points][16 bytes]
constexpr int T16_TILE_BYTES = TILE_M * 256 * 2;         // 64 KB  (NCH = 32)
constexpr int T16N_TILE_BYTES = TILE_M * 64 * 2;         // 16 KB  (NCH = 8: the 64-wide encoding arrays)
template <int NCH = 32>
__host__ __device__ __forceinline__ uint32_t t16_off(int p, int f8) {
    return (uint32_t)(p >> 6) * (uint32_t)(NCH * 1024) + (uint32_t)f8 * 1024u + (uint32_t)(p & 63) * 16u;
}
__device__ __forceinline__ uint4 ldg16(const uint8_t* p) { return *reinterpret_cast<const uint4*>(p); }
__device__ __forceinline__ void stg16(uint8_t* p, uint4 v) { *reinterpret_cast<uint4*>(p) = v; }

__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
    uint32_t r;
    asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
    return r;
}
__device__ __forceinline__ uint32_t pack_f16x2(float lo, float hi) {
    uint32_t r;
    asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
    return r;
}
__device__ __forceinline__ float bf16_lo(uint32_t w) { return __uint_as_float(w << 16); }
__device__ __forceinline__ float bf16_hi(uint32_t w) { return __uint_as_float(w & 0xffff0000u); }
__device__ __forceinline__ float2 f16x2_unpack(uint32_t w) { return __half22float2(*reinterpret_cast<const __half2*>(&w)); }

// softplus(beta = 100) and em = exp(-100 h) = 1 - sigmoid(100 z) from ONE exponential:
//   e = exp(-|100 z|);  h = max(z, 0) + 0.01 log(1 + e);  em = (z >= 0 ? e : 1) / (1 + e)
__device__ __forceinline__ float softplus100_em(float z, float& em) {
    const float e = ex2_approx(-fabsf(z) * 144.26950408889634f);
    const float t = 1.0f + e;
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(t));
    em = (z >= 0.0f ? e : 1.0f) * r;
    return fmaf(lg2_approx(t), 0.006931471805599453f, fmaxf(z, 0.0f));
}

// ---- sweep kernels: A operand in tensor memory, double-buffered -------------------------------------------------------
// Shared memory: weight ring | two input slots | one fp32 scratch tile.  The 16-bit stash tiles an epilogue needs (EM,
// D16) are pulled in by a dedicated producer warp with bulk copies, one N-half (128 columns x 128 points x up to two
// arrays = 64 KB) ahead of the epilogue that reads them, so the epilogue warps never wait on a global load.
constexpr int SW_THREADS = 64 + EPI_THREADS + 32;   // weight producer, MMA issuer, 16 epilogue warps, input producer
constexpr int SW_STAGE_BYTES = 256 * 128;        // one weight stage: [256 rows x 64 k] 16-bit (hi and lo stages alternate)
constexpr int SW_STAGES = 3;
constexpr int SW_IN_OFF = SW_STAGES * SW_STAGE_BYTES;
constexpr int SW_IN_ARRAY_BYTES = 128 * 128 * 2; // one array's N-half: [2 point halves][16 chunks][64 points][16 B]
constexpr int SW_IN_SLOT_BYTES = 2 * SW_IN_ARRAY_BYTES;
constexpr int SW_SMEM_BYTES = SW_IN_OFF + 2 * SW_IN_SLOT_BYTES + 1024;
static_assert(SW_SMEM_BYTES <= 232448, "sweep kernels: shared memory budget");
constexpr uint32_t SW_A0 = 256, SW_A1 = 384;     // TMEM columns of A_hi / A_lo (accumulator: [0, 256))
constexpr int SW_MAX_STEPS = 36;

// one layer: acc[128, n_mma] (TMEM columns 0 ..) = A[:, 0 : 64 kblocks] @ B^T with A = A_hi + A_lo in tensor memory:
//   A_lo B_hi + A_hi B_hi + A_hi B_lo  (three MMAs of N = n_mma per 16-wide k step: 128 cycles each at N = 256, long enough
//   to hide the ~90 cycles one elected thread needs to issue the next one -- 128-column half layers were issue-bound)
struct SwStep {
    uint32_t b_off;          // operand at chain + b_off: kblocks x { hi tile [n_mma x 128 B], lo tile [n_mma x 128 B] }
    uint16_t n_mma;
    uint8_t kblocks;
    uint8_t f16 : 1;         // fp16 operands (normal sweep) instead of bf16
    uint8_t acc_in : 1;      // accumulate onto what the epilogue pre-seeded in the accumulator (tangent sweep, skip layer input)
};
struct SwProgram {
    int n_steps;
    SwStep step[SW_MAX_STEPS];
};
// inputs of one epilogue half-step: N-half `hf` of up to two T16 arrays (pointers to tile 0; b may be NULL)
struct SwInEvent {
    const uint8_t* a;
    const uint8_t* b;
};
struct SwInputs {
    int n_events;            // per tile; event e uses slot e & 1 and N-half e & 1
    SwInEvent ev[SW_MAX_STEPS];
};
struct SwBarriers {
    uint64_t full[SW_STAGES];
    uint64_t empty[SW_STAGES];
    uint64_t in_full[2];
    uint64_t in_empty[2];
    uint64_t a_ready;
    uint64_t acc_full;       // completes once per layer, and a_ready (every epilogue thread) separates two completions: a slow
                             // thread can never see it two phases ahead (a barrier that flips twice aliases its parity)
    uint32_t tmem_base;
};

__device__ __forceinline__ uint8_t* sw_setup(uint8_t* smem_raw, SwBarriers* bar) {
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    const int warp = threadIdx.x >> 5;
    if (warp == 0) tc::tmem_alloc(&bar->tmem_base, 512);
    if (threadIdx.x == 32) {
        for (int s = 0; s < SW_STAGES; ++s) {
            tc::mbar_init(&bar->full[s], 1);
            tc::mbar_init(&bar->empty[s], 1);
        }
        for (int s = 0; s < 2; ++s) {
            tc::mbar_init(&bar->in_full[s], 1);
            tc::mbar_init(&bar->in_empty[s], EPI_THREADS);
        }
        tc::mbar_init(&bar->a_ready, EPI_THREADS);
        tc::mbar_init(&bar->acc_full, 1);
        tc::mbar_fence_init();
    }
    tc::tc_fence_before_sync();
    __syncthreads();
    tc::tc_fence_after_sync();
    return smem;
}
__device__ __forceinline__ void sw_teardown(SwBarriers* bar) {
    tc::tc_fence_before_sync();
    __syncthreads();
    if ((threadIdx.x >> 5) == 0) tc::tmem_dealloc(bar->tmem_base, 512);
}

// Diagnostics (hn_chain16_set_debug): a host-mapped buffer of [4 kernels][4 instances][148 CTAs][8] words; every role
// of every CTA records (tile << 8 | step) as it goes, 0xffffffff when done -- readable from the host while a kernel hangs.
extern uint32_t* g_m16_dbg;
__device__ __forceinline__ void dbg_mark(uint32_t* dbg, int role, uint32_t v) {
    if (dbg) {
        *reinterpret_cast<volatile uint32_t*>(dbg + blockIdx.x * 8 + role) = v;
    }
}
uint32_t* dbg_slot(int kernel_id);

// warp 0, lane 0
__device__ __forceinline__ void sw_producer(const SwProgram& prog, const uint8_t* __restrict__ chain_w, uint8_t* smem,
                                            SwBarriers* bar, int n_my_tiles, uint32_t* dbg = nullptr) {
    uint32_t stage = 0, phase = 0;
    for (int t = 0; t < n_my_tiles; ++t)
        for (int s = 0; s < prog.n_steps; ++s) {
            dbg_mark(dbg, 0, (uint32_t)(t << 8 | s));
            const SwStep st = prog.step[s];
            const uint32_t bytes = (uint32_t)st.n_mma * 128u;
            const uint8_t* src = chain_w + st.b_off;
            for (int kb = 0; kb < st.kblocks; ++kb)
                for (int ps = 0; ps < 2; ++ps) {
                    tc::mbar_wait(&bar->empty[stage], phase ^ 1u);
                    tc::mbar_arrive_expect_tx(&bar->full[stage], bytes);
                    tc::bulk_g2s(smem + stage * SW_STAGE_BYTES, src + (size_t)(2 * kb + ps) * bytes, bytes, &bar->full[stage]);
                    if (++stage == SW_STAGES) { stage = 0; phase ^= 1u; }
                }
        }
    dbg_mark(dbg, 0, 0xffffffffu);
}

// warp 18, lane 0: the stash tiles of every epilogue half-step, one half-step ahead
__device__ __forceinline__ void sw_in_producer(const SwInputs& in, uint8_t* smem, SwBarriers* bar, int n_my_tiles,
                                               uint32_t* dbg = nullptr) {
    uint32_t par[2] = {0u, 0u};
    for (int t = 0; t < n_my_tiles; ++t) {
        const size_t tb = ((size_t)blockIdx.x + (size_t)t * gridDim.x) * T16_TILE_BYTES;
        for (int e = 0; e < in.n_events; ++e) {
            dbg_mark(dbg, 3, (uint32_t)(t << 8 | e));
            const int slot = e & 1;
            const SwInEvent ev = in.ev[e];
            tc::mbar_wait(&bar->in_empty[slot], par[slot] ^ 1u);
            par[slot] ^= 1u;
            tc::mbar_arrive_expect_tx(&bar->in_full[slot], ev.b ? 2u * SW_IN_ARRAY_BYTES : (uint32_t)SW_IN_ARRAY_BYTES);
            uint8_t* dst = smem + SW_IN_OFF + slot * SW_IN_SLOT_BYTES;
#pragma unroll
            for (int ph = 0; ph < 2; ++ph) {       // point halves: 16 chunks x 1 KB each, contiguous in the tile
                tc::bulk_g2s(dst + ph * 16384, ev.a + tb + (size_t)ph * 32768 + (size_t)slot * 16384, 16384, &bar->in_full[slot]);
                if (ev.b)
                    tc::bulk_g2s(dst + SW_IN_ARRAY_BYTES + ph * 16384, ev.b + tb + (size_t)ph * 32768 + (size_t)slot * 16384, 16384,
                                 &bar->in_full[slot]);
            }
        }
    }
    dbg_mark(dbg, 3, 0xffffffffu);
}
// epilogue side: wait for / release the input slot of N-half hf
__device__ __forceinline__ void sw_in_wait(SwBarriers* bar, int hf, uint32_t* par) {
    tc::mbar_wait(&bar->in_full[hf], par[hf]);
    par[hf] ^= 1u;
}
__device__ __forceinline__ void sw_in_release(SwBarriers* bar, int hf) { tc::mbar_arrive(&bar->in_empty[hf]); }
// the 16-byte chunk of (row, column c) of array `arr` in input slot hf (c inside N-half hf)
__device__ __forceinline__ uint4 sw_in_ld(const uint8_t* smem, int hf, int arr, int row, int c) {
    return *reinterpret_cast<const uint4*>(smem + SW_IN_OFF + hf * SW_IN_SLOT_BYTES + arr * SW_IN_ARRAY_BYTES + (row >> 6) * 16384 +
                                           ((c >> 3) & 15) * 1024 + (row & 63) * 16);
}

// [x, sin / cos(2^k x_c)] encoding (utils/fields.py:13-20): column `col` of its tangent J_e(x) dn, from the stored encoding
// e (column-major tile entry of the point: e[128 j] = e_j)
__device__ __forceinline__ float enc_tangent_col(const float* __restrict__ e, float dn0, float dn1, float dn2, int col) {
    if (col < 3) return col == 0 ? dn0 : (col == 1 ? dn1 : dn2);
    if (col >= 63) return 0.0f;
    const int jj = col - 3, c = jj / 20, r = jj - c * 20, k = r >= 10 ? r - 10 : r;
    const float f = (float)(1 << k);
    // column 3 + 20 c + k = sin(2^k x_c) -> f cos dn_c;  column 3 + 20 c + 10 + k = cos -> -f sin dn_c
    const float other = e[(r >= 10 ? col - 10 : col + 10) * TILE_M];
    return (r >= 10 ? -f : f) * other * (c == 0 ? dn0 : (c == 1 ? dn1 : dn2));
}

// warp 1, lane 0
__device__ __forceinline__ void sw_mma(const SwProgram& prog, uint8_t* smem, SwBarriers* bar, int n_my_tiles,
                                       uint32_t* dbg = nullptr) {
    const uint32_t tmem = bar->tmem_base;
    const uint32_t ring = tc::smem_u32(smem);
    uint32_t stage = 0, phase = 0, a_par = 0;
    long long t_a = 0, t_w = 0, tt;
    const long long t0 = clock64();
    for (int t = 0; t < n_my_tiles; ++t)
        for (int s = 0; s < prog.n_steps; ++s) {
            dbg_mark(dbg, 1, (uint32_t)(t << 8 | s));
            const SwStep st = prog.step[s];
            const uint32_t idesc = tc::make_idesc(st.f16 ? tc::FMT_F16 : tc::FMT_BF16, 128, st.n_mma);
            tt = clock64();
            tc::mbar_wait(&bar->a_ready, a_par);
            t_a += clock64() - tt;
            a_par ^= 1u;
            tc::tc_fence_after_sync();
            for (int kb = 0; kb < st.kblocks; ++kb) {
                const uint32_t ah = tmem + SW_A0 + (uint32_t)kb * 32u, al = tmem + SW_A1 + (uint32_t)kb * 32u;
                // stage "hi": A_lo B_hi + A_hi B_hi
                tt = clock64();
                tc::mbar_wait(&bar->full[stage], phase);
                t_w += clock64() - tt;
                tc::tc_fence_after_sync();
                uint64_t dB = tc::make_smem_desc_sw128(ring + stage * SW_STAGE_BYTES);
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    tc::umma_f16_ts(tmem, al + 8u * k, dB + 2 * k, idesc, (kb | k | st.acc_in) != 0);
                    tc::umma_f16_ts(tmem, ah + 8u * k, dB + 2 * k, idesc, 1);
                }
                tc::umma_commit(&bar->empty[stage]);
                if (++stage == SW_STAGES) { stage = 0; phase ^= 1u; }
                // stage "lo": A_hi B_lo
                tt = clock64();
                tc::mbar_wait(&bar->full[stage], phase);
                t_w += clock64() - tt;
                tc::tc_fence_after_sync();
                dB = tc::make_smem_desc_sw128(ring + stage * SW_STAGE_BYTES);
#pragma unroll
                for (int k = 0; k < 4; ++k) tc::umma_f16_ts(tmem, ah + 8u * k, dB + 2 * k, idesc, 1);
                tc::umma_commit(&bar->empty[stage]);
                if (++stage == SW_STAGES) { stage = 0; phase ^= 1u; }
            }
            tc::umma_commit(&bar->acc_full);
        }
    dbg_mark(dbg, 1, 0xffffffffu);
    dbg_mark(dbg, 4, (uint32_t)(t_a >> 4));                  // cycles / 16: MMA issuer waiting for the A operand
    dbg_mark(dbg, 5, (uint32_t)(t_w >> 4));                  // ... for weight stages
    dbg_mark(dbg, 6, (uint32_t)((clock64() - t0) >> 4));     // ... total
}

// epilogue-side handshake
__device__ __forceinline__ void sw_wait_acc(SwBarriers* bar, uint32_t& par) {
    tc::mbar_wait(&bar->acc_full, par);
    par ^= 1u;
    tc::tc_fence_after_sync();
}
__device__ __forceinline__ void sw_publish(SwBarriers* bar) {
    tc::tmem_st_wait();
    tc::tc_fence_before_sync();
    tc::mbar_arrive(&bar->a_ready);
}
// 16 accumulator columns [col0, col0 + 16) of this thread's lane
__device__ __forceinline__ void sw_ld16(uint32_t tmem, uint32_t lane_base, int col0, float* v) {
    tc::tmem_ld_32x32b_x16(tmem + lane_base + (uint32_t)col0, v);
    tc::tmem_ld_wait();
}
__device__ __forceinline__ void sw_ld16_nowait(uint32_t tmem, uint32_t lane_base, int col0, float* v) {
    tc::tmem_ld_32x32b_x16(tmem + lane_base + (uint32_t)col0, v);
}
// 16 operand columns [col0, col0 + 16) (8 packed words) into `abuf` (SW_A0: hi halves, SW_A1: lo halves)
__device__ __forceinline__ void sw_st16(uint32_t tmem, uint32_t lane_base, uint32_t abuf, int col0, const uint32_t* w) {
    tc::tmem_st_32x32b_x8(tmem + lane_base + abuf + (uint32_t)(col0 >> 1), w);
}

// ---- launchers (chain16_obj.cu / chain16_dw.cu) -------------------------------------------------------------------------
int64_t m16_stash_floats(int64_t n);
int64_t m16_bwd_ws_floats(int64_t n);
int launch_m16_fwd(const hn_mlp_t* m, const float* pts, int64_t n, float inv_scale, float* sdf, float* feat, int64_t ld_feat,
                   float* normal, float* stash, cudaStream_t s);
int launch_m16_bwd(const hn_mlp_t* m, int64_t n, float inv_scale, const float* stash, const float* d_sdf, const float* d_feat,
                   int64_t ld_dfeat, const float* d_normal, float* d_pts, const hn_mlp_grad_t* grad, float* ws, cudaStream_t s);

// weight-gradient kernel over 16-bit dW-ready tiles
struct Dw16Job {
    const uint8_t* P[3];     // [tiles] T16 tiles (256 features): M operand (output features)
    const uint8_t* Q[3];     // [tiles] T16 tiles of q_chunks * 8 features: N operand (input features)
    int n_pairs;             // dW = sum over pairs of P[i]^T Q[i]  (three pairs = hi/lo operand split: Ph Qh + Pl Qh + Ph Ql)
    int db_mask;             // pairs whose P enters the bias gradient (0 = pair 0 only)
    int q_chunks;            // 32 (T16) or 8 (T16N)
    int n_mma;               // UMMA N: multiple of 16, <= 8 * q_chunks
    float* db;               // += column sums of P[0] (may be NULL)
    int p_cols;              // valid columns of P (length of db)
};
struct Dw16Params {
    int n_tiles;
    int n_jobs;
    Dw16Job job[12];
    float* part;             // [n_jobs][DW_SPLITS][256][256]
};

struct DwReduceParams;
int launch_dw16(const Dw16Params& p, const DwReduceParams& r, cudaStream_t s);
int launch_out_row0_grad16(const uint8_t* H7, const uint8_t* U7, const float* d_sdf, int64_t n, int n_tiles, float inv_scale,
                           float* dW_row0, float* db0, cudaStream_t s);

}  // namespace chain
}  // namespace hn
