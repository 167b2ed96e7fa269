// tcgen05 tensor-core GEMMs with the same fused epilogues as gemm_simt.cuh (HN_TC_* precisions).
//
//   gemm_tc_kernel<B_TRANS, PASSES, EPI> : C[M,N] = epi( A[M,K] * B ), A = fp32 activations (rows =
//       points, K contiguous), B = fp32 packed weights used as x @ W^T (B_TRANS = false, B(k,n) =
//       W[n*ldb+k]) or as d @ W (B_TRANS = true, B(k,n) = W[k*ldb+n]).
//   gemm_tc_tn_kernel : C[M,N] += A^T * B over a K range (weight gradients: K = points, split over
//       CTAs, fp32 atomics), both operands MN-major straight from their row-major activations.
//
// Operands are converted to TF32 (cvt.rna) while being staged global -> registers -> shared memory
// in the canonical SWIZZLE_128B layouts; accumulation is fp32 in TMEM.  PASSES == 3 splits both
// operands into hi + lo TF32 parts (hi*hi + lo*hi + hi*lo), which restores ~fp32 accuracy
// (oracle/analytic.py measures the precisions).  One CTA owns a 128-row tile and up to 256 output
// columns; the K loop is double-buffered: the MMAs of chunk i overlap the staging of chunk i+1.
// The epilogue moves 32x32 blocks through shared memory so every global access is a coalesced
// 128-byte row segment.
#pragma once
#include "common.cuh"
#include "gemm_simt.cuh"
#include "tc_common.cuh"

namespace hn {

constexpr int TC_THREADS = 256;
constexpr int TC_BM = 128;
constexpr int TC_BN = 256;
constexpr int TC_BK = 32;                 // fp32/tf32 elements per 128-byte swizzle row
constexpr int TC_A_BYTES = TC_BM * 128;   // 16 KB per K chunk
constexpr int TC_B_BYTES = TC_BN * 128;   // 32 KB per K chunk
constexpr int TC_TPAD = 36;               // padded row of the epilogue transpose buffers (floats)

__device__ __forceinline__ float to_tf32(float x) {
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
    return __uint_as_float(r);
}

template <int PASSES>
__device__ __forceinline__ void split_store4(uint8_t* hi_base, uint8_t* lo_base, uint32_t off, float4 v) {
    float4 h = make_float4(to_tf32(v.x), to_tf32(v.y), to_tf32(v.z), to_tf32(v.w));
    *reinterpret_cast<float4*>(hi_base + off) = h;
    if (PASSES == 3) {
        float4 l = make_float4(to_tf32(v.x - h.x), to_tf32(v.y - h.y), to_tf32(v.z - h.z), to_tf32(v.w - h.w));
        *reinterpret_cast<float4*>(lo_base + off) = l;
    }
}

template <int PASSES>
__host__ __device__ constexpr int tc_stage_bytes() { return (PASSES == 3 ? 2 : 1) * (TC_A_BYTES + TC_B_BYTES); }
template <int PASSES>
__host__ __device__ constexpr int tc_smem_bytes() { return 2 * tc_stage_bytes<PASSES>() + 1024; }

// Epilogue shared by both kernels' callers: one warp, a 32-row x 32-column block whose accumulators
// sit in v[32] (thread = row).  T is this warp's [32][TC_TPAD] float buffer.
template <int EPI>
__device__ __forceinline__ void tc_epilogue_block(const GemmArgs& g, int64_t m_base, int n_base, float* v,
                                                  float* T, int lane) {
    const int nvalid = min(32, g.N - n_base);
    const int64_t row = m_base + lane;
    // blocks that straddle nsplit (193) or lie beyond it (C2 output), tails and unaligned outputs take
    // the per-element path
    const bool fast = g.vec_ok && nvalid == 32 && (n_base + 32 <= g.nsplit) && EPI != EPI_ATOMIC;
    if (!fast) {
        if (row < g.M) {
#pragma unroll 4
            for (int j = 0; j < 32; ++j)
                if (j < nvalid) epi_elem<EPI>(g, row, n_base + j, v[j]);
        }
        return;
    }
    auto load_tile = [&](const float* src, int64_t ld, float* dst) {
        // coalesced: lane -> (row = it*4 + lane/8, float4 column lane%8)
#pragma unroll
        for (int it = 0; it < 8; ++it) {
            int r = it * 4 + (lane >> 3), c4 = (lane & 7) * 4;
            float4 t = make_float4(0.f, 0.f, 0.f, 0.f);
            if (m_base + r < g.M) t = ld4(src + (m_base + r) * ld + n_base + c4);
            st4(T + r * TC_TPAD + c4, t);
        }
        __syncwarp();
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            float4 t = ld4(T + lane * TC_TPAD + j * 4);
            dst[j * 4] = t.x; dst[j * 4 + 1] = t.y; dst[j * 4 + 2] = t.z; dst[j * 4 + 3] = t.w;
        }
        __syncwarp();
    };
    auto store_tile = [&](float* dstp, int64_t ld, const float* src) {
#pragma unroll
        for (int j = 0; j < 8; ++j)
            st4(T + lane * TC_TPAD + j * 4, make_float4(src[j * 4], src[j * 4 + 1], src[j * 4 + 2], src[j * 4 + 3]));
        __syncwarp();
#pragma unroll
        for (int it = 0; it < 8; ++it) {
            int r = it * 4 + (lane >> 3), c4 = (lane & 7) * 4;
            if (m_base + r < g.M) st4(dstp + (m_base + r) * ld + n_base + c4, ld4(T + r * TC_TPAD + c4));
        }
        __syncwarp();
    };
    // results are formed in place: v <- primary output, a2 <- secondary output (EPI_TANGENT)
    if (EPI == EPI_STORE) {
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = g.alpha * v[j] + (g.bias ? g.bias[n_base + j] : 0.0f);
    } else if (EPI == EPI_BIAS_SOFTPLUS) {
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = softplus100(v[j] + g.bias[n_base + j]);
    } else if (EPI == EPI_BIAS_RELU) {
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = fmaxf(v[j] + g.bias[n_base + j], 0.0f);
    } else if (EPI == EPI_BIAS_SIGMOID) {
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = sigmoidf_(v[j] + g.bias[n_base + j]);
    } else if (EPI == EPI_MUL_SPRIME) {
        float a1[32];
        load_tile(g.aux1, g.ldaux1, a1);
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] *= sprime_from_h(a1[j]);
    } else if (EPI == EPI_TANGENT) {
        float a1[32], a2[32];
        load_tile(g.aux1, g.ldaux1, a1);
        load_tile(g.C2, g.ldc2, a2);
#pragma unroll
        for (int j = 0; j < 32; ++j) {
            float q = v[j];
            v[j] = sprime_from_h(a1[j]) * q;
            a2[j] = 100.0f * one_minus_sprime_from_h(a1[j]) * a2[j] * q;
        }
        store_tile(g.C2, g.ldc2, a2);
    } else if (EPI == EPI_REVERSE) {
        float a1[32], a2[32];
        load_tile(g.aux1, g.ldaux1, a1);
        load_tile(g.aux2, g.ldaux2, a2);
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = sprime_from_h(a1[j]) * v[j] + a2[j];
    } else if (EPI == EPI_RELU_BWD) {
        float a1[32];
        load_tile(g.aux1, g.ldaux1, a1);
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = a1[j] > 0.0f ? v[j] : 0.0f;
    } else if (EPI == EPI_ADD_AUX) {
        float a1[32];
        load_tile(g.aux1, g.ldaux1, a1);
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] += a1[j];
    }
    store_tile(g.C, g.ldc, v);
}

// -------------------------------------------------------------------------------------------------
template <bool B_TRANS, int PASSES, int EPI>
__global__ void __launch_bounds__(TC_THREADS, PASSES == 3 ? 1 : 2) gemm_tc_kernel(const GemmArgs g) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    __shared__ uint64_t bar_free[2];     // stage buffer may be overwritten (its MMAs completed)
    __shared__ uint64_t bar_done;        // all MMAs of the tile completed
    __shared__ uint32_t tmem_base_s;
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int64_t m0 = (int64_t)blockIdx.x * TC_BM;
    const int n0 = blockIdx.y * TC_BN;
    const int n_tile = min(TC_BN, g.N - n0);
    const int n_mma = (n_tile + 15) & ~15;            // UMMA N: multiple of 16
    const int kchunks = (g.K + TC_BK - 1) / TC_BK;
    constexpr int STAGE = tc_stage_bytes<PASSES>();

    if (warp == 0) tc::tmem_alloc(&tmem_base_s, 256);
    if (tid == 32) {
        tc::mbar_init(&bar_free[0], 1);
        tc::mbar_init(&bar_free[1], 1);
        tc::mbar_init(&bar_done, 1);
        tc::mbar_fence_init();
    }
    tc::tc_fence_before_sync();
    __syncthreads();
    tc::tc_fence_after_sync();
    const uint32_t tmem_base = tmem_base_s;
    const uint32_t idesc = tc::make_idesc(tc::FMT_TF32, 128, (uint32_t)n_mma);

    // Register-staged operands with one chunk of prefetch: the global loads of chunk kc+1 are issued
    // right after chunk kc has been written to shared memory, so their latency overlaps the fence,
    // the barrier, the MMA issue and the wait for the stage buffer.
    constexpr int A_IT = (TC_BM * 8) / TC_THREADS;   // 4 float4 per thread
    constexpr int B_IT = (TC_BN * 8) / TC_THREADS;   // 8 float4 per thread (n_mma <= 256)
    float4 ra[A_IT], rb[B_IT];
    const int nq_per_k = n_mma >> 2;
    auto load_chunk = [&](int kc) {
        const int k0 = kc * TC_BK;
#pragma unroll
        for (int it = 0; it < A_IT; ++it) {
            int idx = tid + it * TC_THREADS;
            int r = idx >> 3, c16 = idx & 7;
            int64_t gm = m0 + r;
            int gk = k0 + c16 * 4;
            // NOTE: nothing here may depend on the loaded VALUE (a tail mask would turn the prefetch
            // into a synchronous load); the k tail of A is masked in store_chunk instead.
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (gm < g.M && gk < g.K) v = ld4(g.A + gm * g.lda + gk);
            ra[it] = v;
        }
#pragma unroll
        for (int it = 0; it < B_IT; ++it) {
            int idx = tid + it * TC_THREADS;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (!B_TRANS) {
                int r = idx >> 3, c16 = idx & 7;
                int gn = n0 + r;
                int gk = k0 + c16 * 4;
                // weights are zero-padded up to their leading dimension, so the k tail needs no mask
                if (r < n_mma && gn < g.N && gk < g.K) v = ld4(g.B + (int64_t)gn * g.ldb + gk);
            } else {
                // B(k,n) = W[k*ldb + n]: float4 along n
                int kk = idx / nq_per_k, nq = idx - kk * nq_per_k;
                int gk = k0 + kk;
                int gn = n0 + nq * 4;
                // columns n >= N only feed accumulator columns that are never stored
                if (kk < TC_BK && gk < g.K && gn < g.N) v = ld4(g.B + (int64_t)gk * g.ldb + gn);
            }
            rb[it] = v;
        }
    };
    auto store_chunk = [&](int kc, uint8_t* sAh, uint8_t* sBh, uint8_t* sAl, uint8_t* sBl) {
#pragma unroll
        for (int it = 0; it < A_IT; ++it) {
            int idx = tid + it * TC_THREADS;
            float4 v = ra[it];
            int gk = kc * TC_BK + (idx & 7) * 4;
            if (gk + 1 >= g.K) v.y = 0.f;     // activations are not guaranteed finite past column K
            if (gk + 2 >= g.K) v.z = 0.f;
            if (gk + 3 >= g.K) v.w = 0.f;
            split_store4<PASSES>(sAh, sAl, tc::sw128_offset(idx >> 3, idx & 7), v);
        }
#pragma unroll
        for (int it = 0; it < B_IT; ++it) {
            int idx = tid + it * TC_THREADS;
            if (!B_TRANS) {
                if ((idx >> 3) < n_mma) split_store4<PASSES>(sBh, sBl, tc::sw128_offset(idx >> 3, idx & 7), rb[it]);
            } else {
                int kk = idx / nq_per_k, nq = idx - kk * nq_per_k;
                if (kk < TC_BK) {
                    const float vv[4] = {rb[it].x, rb[it].y, rb[it].z, rb[it].w};
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        uint32_t off = tc::sw128_offset(nq * 4 + j, kk >> 2) + (kk & 3) * 4;
                        float h = to_tf32(vv[j]);
                        *reinterpret_cast<float*>(sBh + off) = h;
                        if (PASSES == 3) *reinterpret_cast<float*>(sBl + off) = to_tf32(vv[j] - h);
                    }
                }
            }
        }
    };
    load_chunk(0);
    for (int kc = 0; kc < kchunks; ++kc) {
        const int s = kc & 1;
        uint8_t* sAh = smem + (size_t)s * STAGE;
        uint8_t* sBh = sAh + TC_A_BYTES;
        uint8_t* sAl = sBh + TC_B_BYTES;
        uint8_t* sBl = sAl + TC_A_BYTES;
        // wait until the MMAs that read this buffer two chunks ago are done
        if (kc >= 2) tc::mbar_wait(&bar_free[s], ((kc >> 1) - 1) & 1);
        store_chunk(kc, sAh, sBh, sAl, sBl);
        if (kc + 1 < kchunks) load_chunk(kc + 1);
        tc::fence_proxy_async_smem();
        __syncthreads();
        if (warp == 0 && tc::elect_one()) {
            tc::tc_fence_after_sync();
            const uint64_t dAh = tc::make_smem_desc_sw128(tc::smem_u32(sAh));
            const uint64_t dBh = tc::make_smem_desc_sw128(tc::smem_u32(sBh));
            const uint64_t dAl = tc::make_smem_desc_sw128(tc::smem_u32(sAl));
            const uint64_t dBl = tc::make_smem_desc_sw128(tc::smem_u32(sBl));
#pragma unroll
            for (int k = 0; k < TC_BK / 8; ++k) {
                // UMMA_K = 8 tf32 = 32 bytes = 2 units of 16 B inside the swizzle row
                if (PASSES == 3) {
                    tc::umma_tf32(tmem_base, dAl + 2 * k, dBh + 2 * k, idesc, (kc | k) != 0);
                    tc::umma_tf32(tmem_base, dAh + 2 * k, dBl + 2 * k, idesc, 1);
                    tc::umma_tf32(tmem_base, dAh + 2 * k, dBh + 2 * k, idesc, 1);
                } else {
                    tc::umma_tf32(tmem_base, dAh + 2 * k, dBh + 2 * k, idesc, (kc | k) != 0);
                }
            }
            tc::umma_commit(&bar_free[s]);
            if (kc == kchunks - 1) tc::umma_commit(&bar_done);
        }
        __syncwarp();
    }
    tc::mbar_wait(&bar_done, 0);
    tc::tc_fence_after_sync();
    // ---- epilogue: warp w -> TMEM lanes 32*(w%4).., column half (w/4) -------------------------------
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
    tc::tc_fence_before_sync();
    __syncthreads();
    if (warp == 0) tc::tmem_dealloc(tmem_base, 256);
}

// -------------------------------------------------------------------------------------------------
// Weight gradients: C[i,j] += sum_{p in K range} P[p,i] * Q[p,j]  (A = P^T, B = Q^T, K = points).
// TF32 operands only exist K-major with the plain SWIZZLE_128B layout (the MN-major form needs the
// 32-byte-base swizzle), so the transposition happens while staging: a lane owns one output row i
// and gathers 4 consecutive points with 4 coalesced scalar loads (a warp reads 128 contiguous bytes
// of one point row per load), then writes them as one 16-byte chunk of the K-major tile.
template <int PASSES>
__global__ void __launch_bounds__(TC_THREADS, PASSES == 3 ? 1 : 2) gemm_tc_tn_kernel(const GemmArgs g) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    __shared__ uint64_t bar_free[2];
    __shared__ uint64_t bar_done;
    __shared__ uint32_t tmem_base_s;
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int m0 = blockIdx.x * TC_BM;
    const int n0 = blockIdx.y * TC_BN;
    const int n_tile = min(TC_BN, g.N - n0);
    const int n_mma = (n_tile + 15) & ~15;
    const int64_t kbeg = (int64_t)blockIdx.z * g.k_chunk;
    const int64_t kend = min((int64_t)g.K, kbeg + g.k_chunk);
    if (kbeg >= kend) return;
    const int kchunks = (int)((kend - kbeg + TC_BK - 1) / TC_BK);
    constexpr int STAGE = tc_stage_bytes<PASSES>();

    if (warp == 0) tc::tmem_alloc(&tmem_base_s, 256);
    if (tid == 32) {
        tc::mbar_init(&bar_free[0], 1);
        tc::mbar_init(&bar_free[1], 1);
        tc::mbar_init(&bar_done, 1);
        tc::mbar_fence_init();
    }
    tc::tc_fence_before_sync();
    __syncthreads();
    tc::tc_fence_after_sync();
    const uint32_t tmem_base = tmem_base_s;
    const uint32_t idesc = tc::make_idesc(tc::FMT_TF32, 128, (uint32_t)n_mma);

    constexpr int A_IT = (TC_BM * 8) / TC_THREADS;   // 4 items x 4 scalars
    constexpr int B_IT = (TC_BN * 8) / TC_THREADS;   // 8 items x 4 scalars
    float4 ra[A_IT], rb[B_IT];
    auto load_chunk = [&](int kc) {
        const int64_t k0 = kbeg + (int64_t)kc * TC_BK;
#pragma unroll
        for (int it = 0; it < A_IT; ++it) {
            int idx = tid + it * TC_THREADS;
            int c16 = idx >> 7, r = idx & 127;
            int gi = m0 + r;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (gi < g.M) {
                int64_t gp = k0 + c16 * 4;
                const float* src = g.A + gp * g.lda + gi;
                if (gp + 0 < kend) v.x = src[0];
                if (gp + 1 < kend) v.y = src[g.lda];
                if (gp + 2 < kend) v.z = src[2 * g.lda];
                if (gp + 3 < kend) v.w = src[3 * g.lda];
            }
            ra[it] = v;
        }
#pragma unroll
        for (int it = 0; it < B_IT; ++it) {
            int idx = tid + it * TC_THREADS;
            int c16 = idx / n_mma, r = idx - c16 * n_mma;
            int gj = n0 + r;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (c16 < 8 && gj < g.N) {
                int64_t gp = k0 + c16 * 4;
                const float* src = g.B + gp * g.ldb + gj;
                if (gp + 0 < kend) v.x = src[0];
                if (gp + 1 < kend) v.y = src[g.ldb];
                if (gp + 2 < kend) v.z = src[2 * g.ldb];
                if (gp + 3 < kend) v.w = src[3 * g.ldb];
            }
            rb[it] = v;
        }
    };
    auto store_chunk = [&](uint8_t* sA, uint8_t* sB, uint8_t* sAl, uint8_t* sBl) {
#pragma unroll
        for (int it = 0; it < A_IT; ++it) {
            int idx = tid + it * TC_THREADS;
            split_store4<PASSES>(sA, sAl, tc::sw128_offset(idx & 127, idx >> 7), ra[it]);
        }
#pragma unroll
        for (int it = 0; it < B_IT; ++it) {
            int idx = tid + it * TC_THREADS;
            int c16 = idx / n_mma, r = idx - c16 * n_mma;
            if (c16 < 8) split_store4<PASSES>(sB, sBl, tc::sw128_offset(r, c16), rb[it]);
        }
    };
    load_chunk(0);
    for (int kc = 0; kc < kchunks; ++kc) {
        const int s = kc & 1;
        uint8_t* sA = smem + (size_t)s * STAGE;
        uint8_t* sB = sA + TC_A_BYTES;
        uint8_t* sAl = sB + TC_B_BYTES;
        uint8_t* sBl = sAl + TC_A_BYTES;
        if (kc >= 2) tc::mbar_wait(&bar_free[s], ((kc >> 1) - 1) & 1);
        store_chunk(sA, sB, sAl, sBl);
        if (kc + 1 < kchunks) load_chunk(kc + 1);
        tc::fence_proxy_async_smem();
        __syncthreads();
        if (warp == 0 && tc::elect_one()) {
            tc::tc_fence_after_sync();
            const uint64_t dA = tc::make_smem_desc_sw128(tc::smem_u32(sA));
            const uint64_t dB = tc::make_smem_desc_sw128(tc::smem_u32(sB));
            const uint64_t dAl = tc::make_smem_desc_sw128(tc::smem_u32(sAl));
            const uint64_t dBl = tc::make_smem_desc_sw128(tc::smem_u32(sBl));
#pragma unroll
            for (int k = 0; k < TC_BK / 8; ++k) {
                if (PASSES == 3) {
                    tc::umma_tf32(tmem_base, dAl + 2 * k, dB + 2 * k, idesc, (kc | k) != 0);
                    tc::umma_tf32(tmem_base, dA + 2 * k, dBl + 2 * k, idesc, 1);
                    tc::umma_tf32(tmem_base, dA + 2 * k, dB + 2 * k, idesc, 1);
                } else {
                    tc::umma_tf32(tmem_base, dA + 2 * k, dB + 2 * k, idesc, (kc | k) != 0);
                }
            }
            tc::umma_commit(&bar_free[s]);
            if (kc == kchunks - 1) tc::umma_commit(&bar_done);
        }
        __syncwarp();
    }
    tc::mbar_wait(&bar_done, 0);
    tc::tc_fence_after_sync();
    const int q = warp & 3, half = warp >> 2;
    const int64_t row = m0 + q * 32 + lane;
    for (int nb = half * 4; nb < half * 4 + 4; ++nb) {
        if (nb * 32 >= n_tile) break;
        float v[32];
        tc::tmem_ld_32x32b_x32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(nb * 32), v);
        tc::tmem_ld_wait();
        if (row < g.M) {
#pragma unroll
            for (int j = 0; j < 32; ++j) {
                int n = n0 + nb * 32 + j;
                if (n < g.N) atomicAdd(&g.C[row * g.ldc + n], g.alpha * v[j]);
            }
        }
    }
    tc::tc_fence_before_sync();
    __syncthreads();
    if (warp == 0) tc::tmem_dealloc(tmem_base, 256);
}

// -------------------------------------------------------------------------------------------------
template <bool B_TRANS, int PASSES, int EPI>
int launch_gemm_tc(const GemmArgs& g, cudaStream_t stream) {
    if (g.M <= 0 || g.N <= 0 || g.K <= 0) return HN_OK;
    GemmArgs a = g;
    HN_REQUIRE(g.A && g.B && g.C, "gemm_tc: null operand");
    HN_REQUIRE(aligned16(g.A) && aligned16(g.B) && g.lda % 4 == 0 && g.ldb % 4 == 0,
               "gemm_tc: operands must be 16B aligned with leading dimensions that are multiples of 4");
    auto ok = [](const void* p, int64_t ld) { return p == nullptr || (aligned16(p) && ld % 4 == 0); };
    a.vec_ok = ok(g.C, g.ldc) && ok(g.C2, g.ldc2) && ok(g.aux1, g.ldaux1) && ok(g.aux2, g.ldaux2);
    auto kern = gemm_tc_kernel<B_TRANS, PASSES, EPI>;
    constexpr int smem = tc_smem_bytes<PASSES>();
    static bool configured = false;
    if (!configured) {
        HN_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        configured = true;
    }
    dim3 grid((unsigned)ceil_div(g.M, TC_BM), (unsigned)ceil_div(g.N, TC_BN), 1);
    {
        TimingScope ts(stream);
        kern<<<grid, TC_THREADS, smem, stream>>>(a);
    }
    count_launch();
    HN_CHECK_LAUNCH();
    return HN_OK;
}

template <int PASSES>
int launch_gemm_tc_tn(const GemmArgs& g, cudaStream_t stream, int k_splits) {
    if (g.M <= 0 || g.N <= 0 || g.K <= 0) return HN_OK;
    GemmArgs a = g;
    HN_REQUIRE(g.A && g.B && g.C, "gemm_tc_tn: null operand");
    HN_REQUIRE(aligned16(g.A) && aligned16(g.B) && g.lda % 4 == 0 && g.ldb % 4 == 0,
               "gemm_tc_tn: operands must be 16B aligned with leading dimensions that are multiples of 4");
    constexpr int smem = tc_smem_bytes<PASSES>();
    static bool configured = false;
    if (!configured) {
        HN_CHECK_CUDA(cudaFuncSetAttribute(gemm_tc_tn_kernel<PASSES>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        configured = true;
    }
    int chunk = (int)round_up(ceil_div(g.K, std::max(1, k_splits)), TC_BK);
    a.k_chunk = chunk;
    dim3 grid((unsigned)ceil_div(g.M, TC_BM), (unsigned)ceil_div(g.N, TC_BN), (unsigned)ceil_div(g.K, chunk));
    {
        TimingScope ts(stream);
        gemm_tc_tn_kernel<PASSES><<<grid, TC_THREADS, smem, stream>>>(a);
    }
    count_launch();
    HN_CHECK_LAUNCH();
    return HN_OK;
}

}  // namespace hn
