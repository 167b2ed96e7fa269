// Per-layer dense contraction of the HN_TC_BF16X3 precision for nets without a fused chain kernel (the hand field:
// 1386 / 1642 / 1669-wide layers):  C[M,N] = epi( A[M,K] * Bp ),  A = fp32 activations (K contiguous),
// Bp = weights PRE-PACKED as bf16 hi/lo tiles in the tcgen05 shared-memory layout (gemm_bx3_pack below; K-major,
// SWIZZLE_128B, one 64-wide k-block = hi tile [256 x 128 B] + lo tile).  Same fused epilogues as gemm_tc.cuh.
//
// Warp-specialised, no CTA-wide barrier in the K loop:
//   warps 0..7  stage A: global fp32 -> registers (two k-blocks of prefetch) -> bf16 hi/lo -> 3-stage smem ring
//   warp  8     issues the MMAs (three per product: A_lo B_hi + A_hi B_hi + A_hi B_lo, fp32 accumulation in TMEM)
//   warp  9     streams the packed weights with bulk copies into a 2-stage ring
// then warps 0..7 run the epilogue (TMEM -> registers -> fused epilogue, coalesced through shared memory).
// Against the TF32x3 kernel: half the MMA passes per product, no weight staging / splitting in the loop, and the
// staging never waits on a block-wide barrier.
#pragma once
#include "chain_common.cuh"
#include "gemm_tc.cuh"

namespace hn {

constexpr int BX_SWARPS = 8;
constexpr int BX_THREADS = (BX_SWARPS + 2) * 32;
constexpr int BX_A_STAGES = 3, BX_B_STAGES = 2;
constexpr int BX_A_HALF = TC_BM * 128;            // one [128 x 128 B] tile: 16 KB
constexpr int BX_A_BYTES = 2 * BX_A_HALF;         // hi + lo
constexpr int BX_B_HALF = 256 * 128;              // one [256 x 128 B] tile: 32 KB
constexpr int BX_B_BYTES = 2 * BX_B_HALF;         // hi + lo: the size of a packed k-block
constexpr int BX_SMEM_BYTES = BX_A_STAGES * BX_A_BYTES + BX_B_STAGES * BX_B_BYTES + 1024;

// bytes of a packed operand with `rows` output rows (padded to 256-row tiles) and reduction length k
inline int64_t bx3_operand_bytes(int rows, int k) { return ceil_div(rows, 256) * ceil_div(k, 64) * (int64_t)BX_B_BYTES; }
inline int64_t bx3_tile_bytes(int k) { return ceil_div(k, 64) * (int64_t)BX_B_BYTES; }

// Packed operands of a whole net, in one buffer (hn_mlp_t::chain of a net without a chain kernel): per layer the
// operand of x @ W^T (rows = out, k = in) and of d @ W (rows = in, k = out).
struct Bx3Layout {
    int64_t w[HN_MAX_LAYERS], wt[HN_MAX_LAYERS], total;
};
inline Bx3Layout bx3_layout(const hn_mlp_t* m) {
    Bx3Layout L;
    int64_t off = 0;
    for (int l = 0; l < m->n_layers; ++l) {
        L.w[l] = off;
        off += bx3_operand_bytes(m->out_dim[l], m->in_dim[l]);
        L.wt[l] = off;
        off += bx3_operand_bytes(m->in_dim[l], m->out_dim[l]);
    }
    L.total = off;
    return L;
}

template <int EPI>
__global__ void __launch_bounds__(BX_THREADS, 1) gemm_bx3_kernel(const GemmArgs g) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    __shared__ uint64_t a_full[BX_A_STAGES], a_empty[BX_A_STAGES], b_full[BX_B_STAGES], b_empty[BX_B_STAGES], done;
    __shared__ uint32_t tmem_base_s;
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* smem_b = smem + BX_A_STAGES * BX_A_BYTES;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int64_t m0 = (int64_t)blockIdx.x * TC_BM;
    const int n0 = blockIdx.y * TC_BN;
    const int n_tile = min(TC_BN, g.N - n0);
    const int n_mma = (n_tile + 15) & ~15;
    const int kchunks = (g.K + 63) >> 6;

    if (warp == 0) tc::tmem_alloc(&tmem_base_s, 256);
    if (tid == 32) {
        for (int s = 0; s < BX_A_STAGES; ++s) { tc::mbar_init(&a_full[s], BX_SWARPS); tc::mbar_init(&a_empty[s], 1); }
        for (int s = 0; s < BX_B_STAGES; ++s) { tc::mbar_init(&b_full[s], 1); tc::mbar_init(&b_empty[s], 1); }
        tc::mbar_init(&done, 1);
        tc::mbar_fence_init();
    }
    tc::tc_fence_before_sync();
    __syncthreads();
    tc::tc_fence_after_sync();
    const uint32_t tmem_base = tmem_base_s;

    if (warp == BX_SWARPS + 1) {
        // ---- weights: one packed k-block (n_mma rows of the hi tile, n_mma rows of the lo tile) per stage -----------
        if (lane == 0) {
            const uint8_t* src = g.Bp + (int64_t)blockIdx.y * g.bp_tile_bytes + (int64_t)g.bp_kb0 * BX_B_BYTES;
            const uint32_t bytes = (uint32_t)n_mma * 128u;
            uint32_t stage = 0, phase = 0;
            for (int kc = 0; kc < kchunks; ++kc) {
                tc::mbar_wait(&b_empty[stage], phase ^ 1u);
                tc::mbar_arrive_expect_tx(&b_full[stage], 2 * bytes);
                uint8_t* dst = smem_b + stage * BX_B_BYTES;
                tc::bulk_g2s(dst, src + (int64_t)kc * BX_B_BYTES, bytes, &b_full[stage]);
                tc::bulk_g2s(dst + BX_B_HALF, src + (int64_t)kc * BX_B_BYTES + BX_B_HALF, bytes, &b_full[stage]);
                if (++stage == BX_B_STAGES) { stage = 0; phase ^= 1u; }
            }
        }
    } else if (warp == BX_SWARPS) {
        // ---- MMA issuer -------------------------------------------------------------------------------------------------
        if (lane == 0) {
            const uint32_t idesc = tc::make_idesc(tc::FMT_BF16, 128, (uint32_t)n_mma);
            uint32_t sa = 0, pa = 0, sb = 0, pb = 0;
            for (int kc = 0; kc < kchunks; ++kc) {
                tc::mbar_wait(&a_full[sa], pa);
                tc::mbar_wait(&b_full[sb], pb);
                tc::tc_fence_after_sync();
                const uint32_t a_addr = tc::smem_u32(smem + sa * BX_A_BYTES), b_addr = tc::smem_u32(smem_b + sb * BX_B_BYTES);
                const uint64_t dAh = tc::make_smem_desc_sw128(a_addr), dAl = tc::make_smem_desc_sw128(a_addr + BX_A_HALF);
                const uint64_t dBh = tc::make_smem_desc_sw128(b_addr), dBl = tc::make_smem_desc_sw128(b_addr + BX_B_HALF);
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    tc::umma_f16(tmem_base, dAl + 2 * k, dBh + 2 * k, idesc, (kc | k) != 0);
                    tc::umma_f16(tmem_base, dAh + 2 * k, dBh + 2 * k, idesc, 1);
                    tc::umma_f16(tmem_base, dAh + 2 * k, dBl + 2 * k, idesc, 1);
                }
                tc::umma_commit(&a_empty[sa]);
                tc::umma_commit(&b_empty[sb]);
                if (++sa == BX_A_STAGES) { sa = 0; pa ^= 1u; }
                if (++sb == BX_B_STAGES) { sb = 0; pb ^= 1u; }
            }
            tc::umma_commit(&done);
        }
    } else {
        // ---- A staging: item = (row, 8 consecutive k) = two float4 in, one 16-byte chunk of the hi and lo tiles out.
        //      Thread t owns items t, t + 256, t + 512, t + 768: rows t/8 + 32 i, chunk t % 8.
        const int c8 = tid & 7, r0 = tid >> 3;
        const float* arow[4];
        bool rlive[4];
        uint32_t soff[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int r = r0 + 32 * i;
            rlive[i] = m0 + r < g.M;
            arow[i] = g.A + (rlive[i] ? (m0 + r) : 0) * g.lda + c8 * 8;
            soff[i] = tc::sw128_offset((uint32_t)r, (uint32_t)c8);
        }
        float4 buf[3][4][2];
        auto load = [&](int kc, float4 (*r)[2]) {
            if (kc >= kchunks) return;
            const int gk = kc * 64 + c8 * 8;
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                // nothing here may depend on the loaded VALUE (the k tail is masked at store time)
                r[i][0] = (rlive[i] && gk < g.K) ? ld4(arow[i] + kc * 64) : make_float4(0.f, 0.f, 0.f, 0.f);
                r[i][1] = (rlive[i] && gk + 4 < g.K) ? ld4(arow[i] + kc * 64 + 4) : make_float4(0.f, 0.f, 0.f, 0.f);
            }
        };
        uint32_t sa = 0, pa = 0;
        auto process = [&](int kc, float4 (*r)[2]) {
            const int gk = kc * 64 + c8 * 8;
            tc::mbar_wait(&a_empty[sa], pa ^ 1u);
            uint8_t* base = smem + sa * BX_A_BYTES;
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                float v[8] = {r[i][0].x, r[i][0].y, r[i][0].z, r[i][0].w, r[i][1].x, r[i][1].y, r[i][1].z, r[i][1].w};
                if (gk + 8 > g.K) {       // activations are not guaranteed finite past column K
#pragma unroll
                    for (int j = 0; j < 8; ++j)
                        if (gk + j >= g.K) v[j] = 0.0f;
                }
                uint4 hi, lo;
                chain::split2(v[0], v[1], hi.x, lo.x);
                chain::split2(v[2], v[3], hi.y, lo.y);
                chain::split2(v[4], v[5], hi.z, lo.z);
                chain::split2(v[6], v[7], hi.w, lo.w);
                *reinterpret_cast<uint4*>(base + soff[i]) = hi;
                *reinterpret_cast<uint4*>(base + BX_A_HALF + soff[i]) = lo;
            }
            tc::fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0) tc::mbar_arrive(&a_full[sa]);
            if (++sa == BX_A_STAGES) { sa = 0; pa ^= 1u; }
        };
        load(0, buf[0]);
        load(1, buf[1]);
        for (int kc = 0; kc < kchunks; kc += 3) {
            load(kc + 2, buf[2]);
            process(kc, buf[0]);
            if (kc + 1 < kchunks) {
                load(kc + 3, buf[0]);
                process(kc + 1, buf[1]);
            }
            if (kc + 2 < kchunks) {
                load(kc + 4, buf[1]);
                process(kc + 2, buf[2]);
            }
        }
        // ---- epilogue: warp w -> TMEM lanes 32*(w%4).., column half (w/4); the staging ring is free by now ---------------
        tc::mbar_wait(&done, 0);
        tc::tc_fence_after_sync();
        float* T = reinterpret_cast<float*>(smem) + warp * 32 * TC_TPAD;
        const int q = warp & 3, half = warp >> 2;
        for (int nb = half * 4; nb < half * 4 + 4; ++nb) {
            const int n_base = n0 + nb * 32;
            if (nb * 32 >= n_tile) break;
            float v[32];
            tc::tmem_ld_32x32b_x32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(nb * 32), v);
            tc::tmem_ld_wait();
            tc_epilogue_block<EPI>(g, m0 + q * 32, n_base, v, T, lane);
        }
    }
    tc::tc_fence_before_sync();
    __syncthreads();
    if (warp == 0) tc::tmem_dealloc(tmem_base, 256);
}

template <int EPI>
int launch_gemm_bx3(const GemmArgs& g, cudaStream_t stream) {
    if (g.M <= 0 || g.N <= 0 || g.K <= 0) return HN_OK;
    GemmArgs a = g;
    HN_REQUIRE(g.A && g.Bp && g.C, "gemm_bx3: null operand");
    HN_REQUIRE(aligned16(g.A) && aligned16(g.Bp) && g.lda % 4 == 0,
               "gemm_bx3: A must be 16B aligned with a leading dimension that is a multiple of 4");
    auto ok = [](const void* p, int64_t ld) { return p == nullptr || (aligned16(p) && ld % 4 == 0); };
    a.vec_ok = ok(g.C, g.ldc) && ok(g.C2, g.ldc2) && ok(g.aux1, g.ldaux1) && ok(g.aux2, g.ldaux2);
    auto kern = gemm_bx3_kernel<EPI>;
    static bool configured = false;
    if (!configured) {
        HN_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, BX_SMEM_BYTES));
        configured = true;
    }
    dim3 grid((unsigned)ceil_div(g.M, TC_BM), (unsigned)ceil_div(g.N, TC_BN), 1);
    {
        TimingScope ts(stream);
        kern<<<grid, BX_THREADS, BX_SMEM_BYTES, stream>>>(a);
    }
    count_launch();
    HN_CHECK_LAUNCH();
    return HN_OK;
}

}  // namespace hn
