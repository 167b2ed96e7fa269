// Fused tile-chain MLP kernels on tcgen05 (HN_TC_BF16X3): shared machinery.
//
// One persistent CTA per SM walks over 128-point tiles.  A tile's activations never leave the SM
// between layers: they live in shared memory as TWO bf16 matrices (hi + lo, hi = bf16(x),
// lo = bf16(x - hi)) in the canonical K-major SWIZZLE_128B UMMA layout, and every layer is
//     acc[128, N] (fp32, TMEM) = A_hi B_hi^T + A_lo B_hi^T + A_hi B_lo^T
// i.e. three bf16 MMAs per product, which restores ~16 mantissa bits (oracle/analytic.py:
// bf16x3 gives 8e-6 max-abs on the SDF and 5e-5 relative on weight gradients, against the
// 1e-3 / 1e-2 north-star bounds; a single bf16 or tf32 pass does not).
//
// Weights are packed ONCE per parameter version (hn_*_chain_pack) as bf16 hi/lo tiles already in
// the UMMA shared-memory layout, so a warp-specialised producer streams them L2 -> shared memory
// with plain 1-D bulk copies (cp.async.bulk + mbarrier transaction counts), 32 KB per stage.
//
// Warp roles (320 threads):  warp 0 = weight producer, warp 1 = MMA issuer (one thread),
// warps 2..9 = epilogue (TMEM -> registers -> activation math -> next layer's A operand in shared
// memory, plus whatever the mode streams to / from HBM).
#pragma once
#include <cuda_bf16.h>
#include <cuda_fp16.h>

#include "common.cuh"
#include "tc_common.cuh"

namespace hn {
namespace chain {

constexpr int TILE_M = 128;
constexpr int EPI_WARPS = 16;
constexpr int EPI_THREADS = EPI_WARPS * 32;
constexpr int THREADS = 64 + EPI_THREADS;
constexpr int KB_BYTES = TILE_M * 128;            // one 64-column k-block of an A matrix: 16 KB
constexpr int A_KBLOCKS = 4;                      // 256 activation columns
constexpr int A_LO_OFF = A_KBLOCKS * KB_BYTES;    // 64 KB
constexpr int RING_OFF = 2 * A_KBLOCKS * KB_BYTES;   // 128 KB
constexpr int STAGE_BYTES = 256 * 128;            // 32 KB: [256 rows x 64 bf16]
constexpr int STAGES = 3;
constexpr int SMEM_BYTES = RING_OFF + STAGES * STAGE_BYTES + 1024;   // + alignment slack
constexpr int TMEM_COLS = 256;
constexpr int MAX_STEPS = 24;

// One layer of a chain: acc[128, n_mma] = A[:, 64*a_kb0 : 64*(a_kb0+kblocks)] @ B^T.
// B is stored at chain + b_off as kblocks x { hi tile [n_mma x 128 B], lo tile [n_mma x 128 B] }.
struct Step {
    uint32_t b_off;
    uint16_t n_mma;
    uint8_t kblocks;
    uint8_t a_kb0 : 4;
    uint8_t acc_in : 1;      // accumulate onto the previous step's result (a layer whose K is split in two steps)
    uint8_t f16 : 1;         // A and B are fp16 (hi + lo) pairs instead of bf16 pairs: ~2^-21 instead of ~2^-17
                             // relative, but fp16 range -- only for the O(1) forward values (value trunk), never
                             // for cotangents (kind::f16 cannot mix fp16 and bf16 operands in one MMA)
    uint8_t no_wait : 1;     // chain_ts.cu: second column half of a layer, same A operand: do not wait for a_ready
    uint16_t acc_col;        // chain_ts.cu: first TMEM column of this step's accumulator
};
struct Program {
    int n_steps;
    Step step[MAX_STEPS];
};

__host__ __device__ inline uint32_t b_operand_bytes(int n_mma, int kblocks) {
    return (uint32_t)kblocks * 2u * (uint32_t)n_mma * 128u;
}

struct Barriers {
    uint64_t full[STAGES];
    uint64_t empty[STAGES];
    uint64_t a_ready;     // epilogue -> MMA: the A operand of the next step is in shared memory
    uint64_t acc_full;    // MMA -> epilogue: the accumulator of this step is complete
    uint32_t tmem_base;
};

__device__ __forceinline__ float4 ld4(const float* p) { return *reinterpret_cast<const float4*>(p); }
__device__ __forceinline__ void st4(float* p, float4 v) { *reinterpret_cast<float4*>(p) = v; }

// The [points, 256] fp32 arrays the chain kernels exchange through HBM (H, D, U, X, DZ) are TILED:
// [tile][column / 4][row in tile (128)][4 floats], so that the epilogue's access pattern (thread = row,
// four consecutive columns per access) is one contiguous 512-byte segment per warp instruction, and the
// weight-gradient kernel (lane = point) reads them coalesced as well.  n is padded to whole tiles.
constexpr int64_t TILE_FLOATS = (int64_t)TILE_M * 256;
__device__ __forceinline__ int toff(int row, int col) { return (col >> 2) * 512 + row * 4 + (col & 3); }

// The 64-wide per-point arrays (encoding E, its cotangents EB / DE, its tangent UE) are accessed one column at a time by
// threads that each own a row, so their tiles are column-major: [tile][64 columns][128 rows] -- a warp's 32 rows of one
// column are one 128-byte line.  eoff(point, 0) + 128 * col addresses column `col`.
__host__ __device__ __forceinline__ int64_t eoff(int64_t point, int col = 0) {
    return ((point >> 7) << 13) + (int64_t)col * TILE_M + (point & 127);
}

// ---- bf16 hi/lo split ------------------------------------------------------------------------
__device__ __forceinline__ void split2(float a, float b, uint32_t& hi, uint32_t& lo) {
    // written out so that widening hi back to fp32 is one shift / one mask (the bf162 intrinsics cost two more per pair)
    uint32_t h;
    asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(h) : "f"(b), "f"(a));      // low half = a, high half = b
    const float ah = __uint_as_float(h << 16), bh = __uint_as_float(h & 0xffff0000u);
    asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(lo) : "f"(b - bh), "f"(a - ah));
    hi = h;
}
// fp16 pair: hi = fp16(x), lo = fp16(x - hi)
__device__ __forceinline__ void split2_lo16(float a, float b, uint32_t& hi, uint32_t& lo) {
    __half2 h = __floats2half2_rn(a, b);
    float2 hf = __half22float2(h);
    __half2 l = __floats2half2_rn(a - hf.x, b - hf.y);
    hi = *reinterpret_cast<uint32_t*>(&h);
    lo = *reinterpret_cast<uint32_t*>(&l);
}
// 8 consecutive columns [col, col+8) (col % 8 == 0) of row `row`
template <bool LO16 = false>
__device__ __forceinline__ void a_store8(uint8_t* smem, int row, int col, const float* v) {
    uint4 hi, lo;
    if (LO16) {
        split2_lo16(v[0], v[1], hi.x, lo.x);
        split2_lo16(v[2], v[3], hi.y, lo.y);
        split2_lo16(v[4], v[5], hi.z, lo.z);
        split2_lo16(v[6], v[7], hi.w, lo.w);
    } else {
        split2(v[0], v[1], hi.x, lo.x);
        split2(v[2], v[3], hi.y, lo.y);
        split2(v[4], v[5], hi.z, lo.z);
        split2(v[6], v[7], hi.w, lo.w);
    }
    uint32_t off = (uint32_t)(col >> 6) * KB_BYTES + tc::sw128_offset((uint32_t)row, (uint32_t)((col & 63) >> 3));
    *reinterpret_cast<uint4*>(smem + off) = hi;
    *reinterpret_cast<uint4*>(smem + A_LO_OFF + off) = lo;
}
template <bool LO16 = false>
__device__ __forceinline__ void a_store1(uint8_t* smem, int row, int col, float v) {
    uint32_t off = (uint32_t)(col >> 6) * KB_BYTES + tc::sw128_offset((uint32_t)row, (uint32_t)((col & 63) >> 3)) +
                   (uint32_t)(col & 7) * 2u;
    if (LO16) {
        const __half h = __float2half_rn(v);
        *reinterpret_cast<__half*>(smem + off) = h;
        *reinterpret_cast<__half*>(smem + A_LO_OFF + off) = __float2half_rn(v - __half2float(h));
    } else {
        const __nv_bfloat16 h = __float2bfloat16_rn(v);
        *reinterpret_cast<__nv_bfloat16*>(smem + off) = h;
        *reinterpret_cast<__nv_bfloat16*>(smem + A_LO_OFF + off) = __float2bfloat16_rn(v - __bfloat162float(h));
    }
}
// read back 8 consecutive columns as fp32 (hi + lo)
template <bool LO16 = false>
__device__ __forceinline__ void a_load8(const uint8_t* smem, int row, int col, float* v) {
    uint32_t off = (uint32_t)(col >> 6) * KB_BYTES + tc::sw128_offset((uint32_t)row, (uint32_t)((col & 63) >> 3));
    uint4 hi = *reinterpret_cast<const uint4*>(smem + off);
    uint4 lo = *reinterpret_cast<const uint4*>(smem + A_LO_OFF + off);
    const uint32_t h[4] = {hi.x, hi.y, hi.z, hi.w}, l[4] = {lo.x, lo.y, lo.z, lo.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        if (LO16) {
            const float2 a = __half22float2(*reinterpret_cast<const __half2*>(&h[i]));
            const float2 b = __half22float2(*reinterpret_cast<const __half2*>(&l[i]));
            v[2 * i] = a.x + b.x;
            v[2 * i + 1] = a.y + b.y;
        } else {
            v[2 * i] = __uint_as_float(h[i] << 16) + __uint_as_float(l[i] << 16);
            v[2 * i + 1] = __uint_as_float(h[i] & 0xffff0000u) + __uint_as_float(l[i] & 0xffff0000u);
        }
    }
}

__device__ __forceinline__ float ex2_approx(float x) {
    float r;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}
__device__ __forceinline__ float lg2_approx(float x) {     // x in [1, 2] here: no denormal handling needed
    float r;
    asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}
// softplus(beta=100) = max(z, 0) + 0.01 log1p(exp(-|100 z|)).  torch's threshold branch (identity for 100 z > 20)
// agrees with this to fp32 rounding: there the log1p term is < 2.1e-11, far under an ulp of z >= 0.2.  7 issue slots
// (2 MUFU) per element: the raw ex2 / lg2 approximations skip the range fix-ups of __expf / __logf, which cannot
// trigger here (the exponent is <= 0, the log argument lies in [1, 2]).
__device__ __forceinline__ float softplus100_fast(float z) {
    const float e = ex2_approx(-fabsf(z) * 144.26950408889634f);          // exp(-|100 z|)
    return fmaf(lg2_approx(1.0f + e), 0.006931471805599453f, fmaxf(z, 0.0f));   // + 0.01 ln2 lg2(1 + e)
}

// ---- setup / teardown (all threads) -----------------------------------------------------------
__device__ __forceinline__ uint8_t* chain_setup(uint8_t* smem_raw, Barriers* bar) {
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    const int warp = threadIdx.x >> 5;
    if (warp == 0) tc::tmem_alloc(&bar->tmem_base, TMEM_COLS);
    if (threadIdx.x == 32) {
        for (int s = 0; s < STAGES; ++s) {
            tc::mbar_init(&bar->full[s], 1);
            tc::mbar_init(&bar->empty[s], 1);
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
__device__ __forceinline__ void chain_teardown(Barriers* bar) {
    tc::tc_fence_before_sync();
    __syncthreads();
    if ((threadIdx.x >> 5) == 0) tc::tmem_dealloc(bar->tmem_base, TMEM_COLS);
}

// ---- warp 0: stream the packed weights of every step of every tile of this CTA -----------------
__device__ __forceinline__ void producer_loop(const Program& prog, const uint8_t* __restrict__ chain_w,
                                              uint8_t* smem, Barriers* bar, int n_my_tiles) {
    if ((threadIdx.x & 31) != 0) return;
    uint32_t stage = 0, phase = 0;
    for (int t = 0; t < n_my_tiles; ++t) {
        for (int s = 0; s < prog.n_steps; ++s) {
            const Step st = prog.step[s];
            const uint32_t bytes = (uint32_t)st.n_mma * 128u;
            const uint8_t* src = chain_w + st.b_off;
            for (int c = 0; c < 2 * st.kblocks; ++c) {
                tc::mbar_wait(&bar->empty[stage], phase ^ 1u);
                tc::mbar_arrive_expect_tx(&bar->full[stage], bytes);
                tc::bulk_g2s(smem + RING_OFF + stage * STAGE_BYTES, src + (size_t)c * bytes, bytes, &bar->full[stage]);
                if (++stage == STAGES) { stage = 0; phase ^= 1u; }
            }
        }
    }
}

// ---- warp 1: issue the MMAs ----------------------------------------------------------------------
// prof (may be NULL): per CTA {cycles waiting for the A operand, cycles waiting for weights, total cycles}
__device__ __forceinline__ void mma_loop(const Program& prog, uint8_t* smem, Barriers* bar, int n_my_tiles,
                                         long long* prof = nullptr) {
    if ((threadIdx.x & 31) != 0) return;
    long long t_a = 0, t_w = 0, t0 = clock64(), tt;
    const uint32_t tmem = bar->tmem_base;
    const uint32_t a_hi = tc::smem_u32(smem), a_lo = a_hi + A_LO_OFF, ring = a_hi + RING_OFF;
    uint32_t stage = 0, phase = 0, a_par = 0;
    for (int t = 0; t < n_my_tiles; ++t) {
        for (int s = 0; s < prog.n_steps; ++s) {
            const Step st = prog.step[s];
            const uint32_t idesc = tc::make_idesc(st.f16 ? tc::FMT_F16 : tc::FMT_BF16, 128, st.n_mma);
            tt = clock64();
            tc::mbar_wait(&bar->a_ready, a_par);
            t_a += clock64() - tt;
            if (prof) prof[148 * 4 + blockIdx.x * 32 + s] += clock64() - tt;      // per-step epilogue wait
            if (prof && t == 0 && s == 0) prof[148 * 4 + blockIdx.x * 32 + 31] += clock64() - tt;   // ... of the first tile
            a_par ^= 1u;
            tc::tc_fence_after_sync();
            for (int kb = 0; kb < st.kblocks; ++kb) {
                const uint64_t dAh = tc::make_smem_desc_sw128(a_hi + (uint32_t)(st.a_kb0 + kb) * KB_BYTES);
                const uint64_t dAl = tc::make_smem_desc_sw128(a_lo + (uint32_t)(st.a_kb0 + kb) * KB_BYTES);
                // stage "hi": A_lo B_hi^T + A_hi B_hi^T
                tt = clock64();
                tc::mbar_wait(&bar->full[stage], phase);
                t_w += clock64() - tt;
                tc::tc_fence_after_sync();
                uint64_t dB = tc::make_smem_desc_sw128(ring + stage * STAGE_BYTES);
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    tc::umma_f16(tmem, dAl + 2 * k, dB + 2 * k, idesc, (kb | k | st.acc_in) != 0);
                    tc::umma_f16(tmem, dAh + 2 * k, dB + 2 * k, idesc, 1);
                }
                tc::umma_commit(&bar->empty[stage]);
                if (++stage == STAGES) { stage = 0; phase ^= 1u; }
                // stage "lo": A_hi B_lo^T
                tt = clock64();
                tc::mbar_wait(&bar->full[stage], phase);
                t_w += clock64() - tt;
                tc::tc_fence_after_sync();
                dB = tc::make_smem_desc_sw128(ring + stage * STAGE_BYTES);
#pragma unroll
                for (int k = 0; k < 4; ++k) tc::umma_f16(tmem, dAh + 2 * k, dB + 2 * k, idesc, 1);
                tc::umma_commit(&bar->empty[stage]);
                if (++stage == STAGES) { stage = 0; phase ^= 1u; }
            }
            tc::umma_commit(&bar->acc_full);
        }
    }
    if (prof) {
        prof[blockIdx.x * 4 + 0] = t_a;
        prof[blockIdx.x * 4 + 1] = t_w;
        prof[blockIdx.x * 4 + 2] = clock64() - t0;
    }
}

// All CTAs run the same program, so without help their HBM-heavy epilogues and their MMA phases line up
// across the chip (HBM saturated, then idle).  Shifting CTAs by a fraction of a layer period spreads the
// memory phases over time.  Called once by the epilogue warps before the first tile.
__device__ __forceinline__ void stagger_start(int cycles_per_slot, int slots) {
    const long long until = clock64() + (long long)(blockIdx.x % slots) * cycles_per_slot;
    while (clock64() < until) {
    }
}

// epilogue side of the handshake
__device__ __forceinline__ void epi_publish_a(Barriers* bar) {
    tc::tc_fence_before_sync();       // orders this thread's tcgen05.ld before the MMAs that overwrite the accumulator
    tc::fence_proxy_async_smem();     // generic-proxy stores of the A operand -> visible to the MMA's async proxy
    tc::mbar_arrive(&bar->a_ready);
}
__device__ __forceinline__ void epi_wait_acc(Barriers* bar, uint32_t& par) {
    tc::mbar_wait(&bar->acc_full, par);
    par ^= 1u;
    tc::tc_fence_after_sync();
}
// rows 32*(warp%4) + lane (hardware: a warp reads the TMEM lane quarter warp_id % 4)
// column group cg = 0..EPI_WARPS/4-1 of (256 / (EPI_WARPS/4)) columns
constexpr int EPI_CGROUPS = EPI_WARPS / 4;
constexpr int EPI_COLS = 256 / EPI_CGROUPS;
__device__ __forceinline__ void epi_coords(int& row, int& cg) {
    const int warp = threadIdx.x >> 5;
    row = (warp & 3) * 32 + (threadIdx.x & 31);
    cg = (warp - 2) >> 2;
}
__device__ __forceinline__ void acc_load32(uint32_t tmem_base, int row, int col0, float* v) {
    tc::tmem_ld_32x32b_x32(tmem_base + ((uint32_t)(row & ~31) << 16) + (uint32_t)col0, v);
    tc::tmem_ld_wait();
}
__device__ __forceinline__ void acc_load32_nowait(uint32_t tmem_base, int row, int col0, float* v) {
    tc::tmem_ld_32x32b_x32(tmem_base + ((uint32_t)(row & ~31) << 16) + (uint32_t)col0, v);
}

// Streams this thread's EPI_COLS accumulator columns in sub-blocks of SUB (8 or 16) through
// f(col0, v[SUB], aux), with up to two tiled fp32 input arrays (a0, a1: tile base pointers) software-pipelined
// one sub-block ahead in registers.  The first loads are issued BEFORE the wait for the accumulator, so their
// latency overlaps the MMAs.  SUB = 8 keeps two arrays within the 96-register budget of the 576-thread CTA.
template <int NA, int SUB, class F>
__device__ __forceinline__ void epi_stream(Barriers* bar, uint32_t& acc_par, uint32_t tmem, int row, int cg, bool live,
                                           bool active, const float* __restrict__ a0, const float* __restrict__ a1,
                                           F&& f) {
    constexpr int NSB = EPI_COLS / SUB, NQ = SUB / 4;
    float4 buf[2][NA][NQ];
    auto issue = [&](int sb, int slot) {
#pragma unroll
        for (int q = 0; q < NQ; ++q) {
            const int off = toff(row, cg * EPI_COLS + sb * SUB + q * 4);
            buf[slot][0][q] = live ? ld4(a0 + off) : make_float4(0.f, 0.f, 0.f, 0.f);
            if (NA > 1) buf[slot][NA - 1][q] = live ? ld4(a1 + off) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
    };
    if (active) issue(0, 0);
    epi_wait_acc(bar, acc_par);
    if (!active) return;
#pragma unroll
    for (int sb = 0; sb < NSB; ++sb) {
        if (sb + 1 < NSB) issue(sb + 1, (sb + 1) & 1);
        const int col0 = cg * EPI_COLS + sb * SUB;
        float v[SUB];
        const uint32_t taddr = tmem + ((uint32_t)(row & ~31) << 16) + (uint32_t)col0;
        if (SUB == 16) tc::tmem_ld_32x32b_x16(taddr, v); else tc::tmem_ld_32x32b_x8(taddr, v);
        tc::tmem_ld_wait();
        f(col0, v, buf[sb & 1]);
    }
}

// ---- packing ---------------------------------------------------------------------------------------
// dst tile element (n, k) <- src[(row0 + n) * ld + col0 + k] for n < rows, k < cols, else 0;
// layout: k-block kb at kb * 2 * n_pad * 128 B: hi tile then lo tile, each [n_pad x 128 B] SW128.
// Row / column maps let an operand gather two ranges of the source: n < rsplit -> row0 + n, else
// row1 + (n - rsplit); k < ksplit -> col0 + k, else col1 + (k - ksplit).
struct PackMap {
    int row0, rsplit, row1, col0, ksplit, col1;
};
inline PackMap pack_map(int row0, int col0) { return PackMap{row0, 1 << 30, 0, col0, 1 << 30, 0}; }
int launch_pack_b(const float* src, int64_t ld, PackMap map, int rows, int cols, int n_pad, int kblocks,
                  uint8_t* dst, cudaStream_t stream, bool lo16 = false);
inline int launch_pack_b(const float* src, int64_t ld, int row0, int col0, int rows, int cols, int n_pad, int kblocks,
                         uint8_t* dst, cudaStream_t stream) {
    return launch_pack_b(src, ld, pack_map(row0, col0), rows, cols, n_pad, kblocks, dst, stream);
}
// All operands of a net in ONE launch: launch_pack_b calls made between pack_batch_begin() and pack_batch_flush()
// are only recorded (up to PACK_MAX_JOBS) and run as one kernel (blockIdx.y = operand).
constexpr int PACK_MAX_JOBS = 48;
void pack_batch_begin();
int pack_batch_flush(cudaStream_t stream);

}  // namespace chain
}  // namespace hn
