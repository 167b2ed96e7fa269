// Weight-gradient contractions of the fused chain path (HN_TC_BF16X3):
//     dW_l = sum over points of  P_a[p,:]^T Q_a[p,:]  (+ P_b[p,:]^T Q_b[p,:])
// for ALL layers of a network in ONE launch.  The reduction dimension is the point index, so both
// operands are consumed "MN-major" (feature index contiguous) exactly as the chain kernels leave them in
// HBM.  Every staging thread runs its own cp.async pipeline (fp32 pieces land in a private shared-memory
// slot several stages ahead, so HBM latency is never exposed and no registers are held across it), then
// splits its values into bf16 hi + lo in the UMMA operand tiles (three MMAs per product, fp32
// accumulation in TMEM: 2 x [128 x 256] accumulators = all 512 TMEM columns).
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
constexpr int DW_KP = 16;                // points per stage (one UMMA K step)
constexpr int DW_OPER_BYTES = 256 * DW_KP * 2;        // one bf16 matrix of a stage: 8 KB
constexpr int DW_STAGE_BYTES = 4 * DW_OPER_BYTES;     // P_hi, P_lo, Q_hi, Q_lo
constexpr int DW_STAGES = 3;
constexpr int DW_RAW_SLOTS = 4;          // fp32 landing slots of the per-thread cp.async pipeline
constexpr int DW_RAW_BYTES = 4 * DW_SWARPS * 32 * 16; // 4 x 16 bytes per staging thread: 32 KB
constexpr int DW_SMEM_BYTES = DW_STAGES * DW_STAGE_BYTES + DW_RAW_SLOTS * DW_RAW_BYTES + 1024;
constexpr int DW_SBO = 1024;                     // next group of 8 points
constexpr int DW_LBO = (DW_KP / 8) * 1024;       // next block of 64 features
static_assert(DW_KP * 32 == DW_SWARPS * 32, "one (point, 8-feature chunk) pair of each operand per staging thread");

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

__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src, uint32_t src_bytes) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
    asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}

// Per-thread view of one operand: the thread owns columns [8 fc, 8 fc + 8) of point k of every 16-point group, read
// as two 16-byte pieces.  The source pointer only ever advances by constant strides, so issuing a stage costs a
// handful of instructions whatever the operand's layout.
// Column-major tiles hold 4 consecutive POINTS of one column in 16 bytes, so there the warp shares its landing slots:
// the warp's 32 rows x 16 columns are 64 pieces, lane L fetches pieces L and L + 32 (column 16 w + p / 4, points
// 4 (p % 4) ..) into its own two slots and every lane then reads its 8 columns out of the warp's slots.
struct DwStream {
    const float* ptr;      // piece 0 of the next group to fetch
    int step;              // floats to the next 16-point group
    uint32_t flags;        // bit 0 / 1: piece 0 / 1 holds operand columns; bit 2: tiled layout; bit 3: column-major tiles
                           // (bits 8..15: their column count); bits 4..7: valid columns of the chunk
    int rem;               // column-major: points from the piece's first point (first group) to the end of the operand

    __device__ __forceinline__ void init(const DwOperand& op, int tile0, int64_t n, int tid) {
        const int k = tid & (DW_KP - 1), fc = tid >> 4;
        const int c = fc * 8;
        const int n_valid = max(0, min(8, op.cols - c));
        flags = (n_valid > 0 ? 1u : 0u) | (n_valid > 4 ? 2u : 0u) | ((uint32_t)n_valid << 4);
        rem = 0;
        if (op.tiled == 2) {
            const int lane = tid & 31, col = 16 * (tid >> 5) + (lane >> 2), q = lane & 3;
            flags = (col < op.cols ? 1u : 0u) | (col + 8 < op.cols ? 2u : 0u) | (8u << 4) | 8u | ((uint32_t)op.ld << 8);
            ptr = op.ptr + ((int64_t)tile0 * op.ld + col) * TILE_M + q * 4;
            step = DW_KP;
            rem = (int)max((int64_t)-4, min((int64_t)1 << 30, n - ((int64_t)tile0 * TILE_M + q * 4)));
            if (!(flags & 1u)) { ptr = op.ptr; step = 0; flags = (flags & ~(255u << 8)) | (1u << 8); }      // never moves
        } else if (n_valid == 0) {          // nothing to read: keep a valid address and never move
            ptr = op.ptr; step = 0;
        } else if (op.tiled == 1) {
            ptr = op.ptr + ((int64_t)tile0 * 64 + fc * 2) * 512 + k * 4;
            step = DW_KP * 4;
            flags |= 4u;
        } else {
            ptr = op.ptr + ((int64_t)tile0 * TILE_M + k) * op.ld + c;
            step = DW_KP * (int)op.ld;
        }
    }
    // start fetching group g (`live`: the thread's point exists; `last`: g is the last group of a 128-point tile)
    template <bool CM>       // CM: the job has column-major operands (compiled out of the common loop)
    __device__ __forceinline__ void fetch(int g, bool live, bool last, uint32_t dst0, uint32_t dst1) {
        const bool tiled = (flags & 4u) != 0u;
        if (CM && (flags & 8u)) {
            const uint32_t bytes = 4u * (uint32_t)max(0, min(4, rem - g * DW_KP));       // ragged last tile: zero-filled
            cp_async16(dst0, ptr, (flags & 1u) ? bytes : 0u);
            cp_async16(dst1, ptr + ((flags & 2u) ? 8 * TILE_M : 0), (flags & 2u) ? bytes : 0u);
            ptr += step + (last ? (int)((flags >> 8) & 255u) * TILE_M - TILE_M : 0);
            return;
        }
        cp_async16(dst0, ptr, (live && (flags & 1u)) ? 16u : 0u);
        cp_async16(dst1, ptr + ((flags & 2u) ? (tiled ? 512 : 4) : 0), (live && (flags & 2u)) ? 16u : 0u);
        ptr += step + ((tiled && last) ? 64 * 512 - 512 : 0);
    }
    template <bool CM>
    __device__ __forceinline__ void take(uint32_t src0, uint32_t src1, int tid, float* v) const {
        if (CM && (flags & 8u)) {           // (after a warp sync) column 8 fc + i of point k: 64 i + 4 k into the warp's slots
            const uint32_t base = (((tid >> 4) & 1) ? src1 : src0) - (uint32_t)(tid & 31) * 16u + (uint32_t)(tid & (DW_KP - 1)) * 4u;
#pragma unroll
            for (int i = 0; i < 8; ++i) asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v[i]) : "r"(base + 64u * i) : "memory");
            return;
        }
        asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v[0]), "=f"(v[1]), "=f"(v[2]), "=f"(v[3]) : "r"(src0) : "memory");
        asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v[4]), "=f"(v[5]), "=f"(v[6]), "=f"(v[7]) : "r"(src1) : "memory");
        const int n_valid = (int)((flags >> 4) & 15u);
        if (n_valid < 8) {       // columns past the operand inside a fetched piece
#pragma unroll
            for (int i = 0; i < 8; ++i)
                if (i >= n_valid) v[i] = 0.0f;
        }
    }
};
__device__ __forceinline__ bool dw_unaligned(const DwOperand& op) {
    return !op.tiled && (((op.ld & 3) | (int)(reinterpret_cast<uintptr_t>(op.ptr) & 15)) != 0);
}
// synchronous scalar read of a chunk (operands that cannot be fetched in aligned 16-byte pieces)
__device__ __forceinline__ void load8_scalar(const DwOperand& op, int64_t pnt, int64_t n, int fc, float* v) {
    const int c = fc * 8;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        float x = 0.0f;
        if (pnt < n && c + i < op.cols)
            x = op.tiled == 1   ? op.ptr[((pnt >> 7) * 64 + ((c + i) >> 2)) * 512 + (pnt & 127) * 4 + ((c + i) & 3)]
                : op.tiled == 2 ? op.ptr[((pnt >> 7) * op.ld + c + i) * TILE_M + (pnt & 127)]
                                : op.ptr[pnt * op.ld + c + i];
        v[i] = x;
    }
}
// split 8 values into bf16 hi + lo and store them at the thread's chunk of the hi / lo stage matrices
__device__ __forceinline__ void store8_mn(uint32_t hi_addr, uint32_t lo_addr, const float* v) {
    uint4 hi, lo;
    split2(v[0], v[1], hi.x, lo.x);
    split2(v[2], v[3], hi.y, lo.y);
    split2(v[4], v[5], hi.z, lo.z);
    split2(v[6], v[7], hi.w, lo.w);
    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(hi_addr), "r"(hi.x), "r"(hi.y), "r"(hi.z), "r"(hi.w) : "memory");
    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(lo_addr), "r"(lo.x), "r"(lo.y), "r"(lo.z), "r"(lo.w) : "memory");
}

// The common staging loop of a CTA (all operands readable in aligned pieces): see dw_kernel.
template <bool CM>
__device__ __forceinline__ void stage_streams(const DwJob& job, int64_t n, int t0, int n_groups, int tid, uint32_t smem0,
                                              uint32_t raw0, uint32_t soff, uint64_t* full, uint64_t* empty, uint32_t& stage,
                                              uint32_t& phase, float* bsum) {
    const int k = tid & (DW_KP - 1), lane = tid & 31;
    const bool q_used = (tid >> 4) * 8 < job.n_mma;      // the MMA reads n_mma features of Q: a narrow operand costs its warps only
    const int64_t pnt0 = (int64_t)t0 * TILE_M + k;
    DwStream sp[2], sq[2];
#pragma unroll
    for (int a = 0; a < 2; ++a) {
        sp[a].init(job.P[a < job.n_pairs ? a : 0], t0, n, tid);
        sq[a].init(job.Q[a < job.n_pairs ? a : 0], t0, n, tid);
    }
    // groups of this thread's point that exist (the last tile may be ragged)
    const int g_live = (int)max((int64_t)0, min((int64_t)n_groups, (n - pnt0 + DW_KP - 1) / DW_KP));
    uint32_t slot_head = 0, slot_cur = 0;
    auto issue = [&](int pair, int g) {          // `pair` is a compile-time constant at every call site
        if (g < n_groups) {
            const uint32_t dst = raw0 + slot_head * DW_RAW_BYTES;
            const bool live = g < g_live;
            const bool last = (g & (TILE_M / DW_KP - 1)) == TILE_M / DW_KP - 1;
            sp[pair].template fetch<CM>(g, live, last, dst, dst + 512 * 16);
            if (q_used) sq[pair].template fetch<CM>(g, live, last, dst + 2 * 512 * 16, dst + 3 * 512 * 16);
        }
        cp_async_commit();
        if (++slot_head == DW_RAW_SLOTS) slot_head = 0;
    };
    auto take = [&](int pair, float* vp, float* vq) {
        const uint32_t src = raw0 + slot_cur * DW_RAW_BYTES;
        sp[pair].template take<CM>(src, src + 512 * 16, tid, vp);
        if (q_used) sq[pair].template take<CM>(src + 2 * 512 * 16, src + 3 * 512 * 16, tid, vq);
        if (pair == 0) {
#pragma unroll
            for (int i = 0; i < 8; ++i) bsum[i] += vp[i];
        }
        if (++slot_cur == DW_RAW_SLOTS) slot_cur = 0;
    };
    // Two stages per trip (their count is even): one proxy fence and one warp sync for both, and the loads of
    // one overlap the conversion of the other.  Stages s+3, s+4 are fetched into the slots stages s, s+1 vacate.
    auto two_stages = [&](int pa, int pb, int pc, int gc, int pd, int gd) {
        cp_async_wait<1>();
        if (CM) __syncwarp();       // the warp's landing slots are shared: everybody's pieces have landed ...
        float vp0[8], vq0[8], vp1[8], vq1[8];
        take(pa, vp0, vq0);
        take(pb, vp1, vq1);
        if (CM) __syncwarp();       // ... and have been read before they are refilled
        issue(pc, gc);
        issue(pd, gd);
        const uint32_t st0 = stage, ph0 = phase;
        if (++stage == DW_STAGES) { stage = 0; phase ^= 1u; }
        const uint32_t st1 = stage, ph1 = phase;
        if (++stage == DW_STAGES) { stage = 0; phase ^= 1u; }
        tc::mbar_wait(&empty[st0], ph0 ^ 1u);
        const uint32_t b0 = smem0 + st0 * DW_STAGE_BYTES + soff;
        store8_mn(b0, b0 + DW_OPER_BYTES, vp0);
        if (q_used) store8_mn(b0 + 2 * DW_OPER_BYTES, b0 + 3 * DW_OPER_BYTES, vq0);
        tc::mbar_wait(&empty[st1], ph1 ^ 1u);
        const uint32_t b1 = smem0 + st1 * DW_STAGE_BYTES + soff;
        store8_mn(b1, b1 + DW_OPER_BYTES, vp1);
        if (q_used) store8_mn(b1 + 2 * DW_OPER_BYTES, b1 + 3 * DW_OPER_BYTES, vq1);
        tc::fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) {
            tc::mbar_arrive(&full[st0]);
            tc::mbar_arrive(&full[st1]);
        }
    };
    static_assert(DW_RAW_SLOTS == 4 && (TILE_M / DW_KP) % 2 == 0, "look-ahead pattern below: three stages in flight");
    if (job.n_pairs == 2) {
        // stage order (g,0) (g,1) (g+1,0) ...
        issue(0, 0); issue(1, 0); issue(0, 1);
        for (int g = 0; g < n_groups; ++g) two_stages(0, 1, 1, g + 1, 0, g + 2);
    } else {
        issue(0, 0); issue(0, 1); issue(0, 2);
        for (int g = 0; g < n_groups; g += 2) two_stages(0, 0, 0, g + 3, 0, g + 4);
    }
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
            tc::mbar_init(&full[s], DW_SWARPS);
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
        // ---- staging: HBM fp32 -> (cp.async, DW_RAW_SLOTS deep, private to the thread) -> bf16 hi/lo MN-major tiles.
        //      Thread = (point k of the 16-point group, 8-feature chunk fc) of both operands.  Stages run over
        //      (group, operand pair); the fetch of the stage DW_RAW_SLOTS - 1 ahead is issued before each conversion.
        const int k = tid & (DW_KP - 1), fc = tid >> 4;
        const uint32_t smem0 = tc::smem_u32(smem);
        const uint32_t raw0 = smem0 + DW_STAGES * DW_STAGE_BYTES + (uint32_t)tid * 16;
        const uint32_t soff = (uint32_t)(fc >> 3) * DW_LBO + (uint32_t)(k >> 3) * DW_SBO + tc::sw128_offset((uint32_t)(k & 7), (uint32_t)(fc & 7));
        const int n_groups = (t1 - t0) * (TILE_M / DW_KP);
        const int64_t pnt0 = (int64_t)t0 * TILE_M + k;
        float bsum[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) bsum[i] = 0.0f;
        uint32_t stage = 0, phase = 0;
        auto publish = [&](const float* vp, const float* vq) {      // convert + hand one stage to the MMA warp
            tc::mbar_wait(&empty[stage], phase ^ 1u);
            const uint32_t base = smem0 + stage * DW_STAGE_BYTES + soff;
            store8_mn(base, base + DW_OPER_BYTES, vp);
            store8_mn(base + 2 * DW_OPER_BYTES, base + 3 * DW_OPER_BYTES, vq);
            tc::fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0) tc::mbar_arrive(&full[stage]);
            if (++stage == DW_STAGES) { stage = 0; phase ^= 1u; }
        };
        bool unaligned = false, colmajor = false;
        for (int a = 0; a < job.n_pairs; ++a) {
            unaligned = unaligned || dw_unaligned(job.P[a]) || dw_unaligned(job.Q[a]);
            colmajor = colmajor || job.P[a].tiled == 2 || job.Q[a].tiled == 2;
        }
        if (unaligned) {
            // ---- rare: an operand with an odd leading dimension; plain loads, latency exposed --------------------
            for (int g = 0; g < n_groups; ++g)
                for (int pair = 0; pair < job.n_pairs; ++pair) {
                    float vp[8], vq[8];
                    load8_scalar(job.P[pair], pnt0 + (int64_t)g * DW_KP, p.n, fc, vp);
                    load8_scalar(job.Q[pair], pnt0 + (int64_t)g * DW_KP, p.n, fc, vq);
                    if (pair == 0) {
#pragma unroll
                        for (int i = 0; i < 8; ++i) bsum[i] += vp[i];
                    }
                    publish(vp, vq);
                }
        } else {
            if (colmajor) stage_streams<true>(job, p.n, t0, n_groups, tid, smem0, raw0, soff, full, empty, stage, phase, bsum);
            else stage_streams<false>(job, p.n, t0, n_groups, tid, smem0, raw0, soff, full, empty, stage, phase, bsum);
        }
        cp_async_wait<0>();
        if (job.db) {
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                float s = bsum[i];
#pragma unroll
                for (int o = DW_KP / 2; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
                const int c = fc * 8 + i;
                if (k == 0 && c < job.P[0].cols && s != 0.0f) atomicAdd(job.db + c, s * job.db_scale);
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

int launch_dw_reduce(const DwReduceParams& r, int n_jobs, cudaStream_t s) {
    dw_reduce_kernel<<<dim3(65536 / 4 / 256, n_jobs), 256, 0, s>>>(r);
    count_launch();
    HN_CHECK_LAUNCH();
    return HN_OK;
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
    return launch_dw_reduce(r, p.n_jobs, s);
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
