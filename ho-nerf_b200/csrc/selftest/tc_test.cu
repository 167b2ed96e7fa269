// SELF-TEST library (libhonerf_b200_selftest.so, NOT part of the product library): Bring-up / self-test kernel for the tcgen05 path: C[M,N] = A[M,K] * B[N,K]^T with fp16 operands,
// fp32 accumulation in TMEM.  One CTA per 128-row tile, operands staged into shared memory with the
// canonical SWIZZLE_128B K-major layout by ordinary stores.  Exercises exactly the descriptors, TMEM
// allocation, MMA issue, commit/mbarrier and tcgen05.ld epilogue that the fused field kernels use,
// against torch.matmul in tests/test_gpu_tc.py.
#include "../common.cuh"
#include "../../../include/honerf_b200_selftest.h"
#include "../tc_common.cuh"

namespace hn {

// smem: A tile [K/64][128 rows x 128 B] then B [K/64][N rows x 128 B]
template <typename T>
__global__ void __launch_bounds__(128, 1) tc_gemm_test_kernel(const T* __restrict__ A, const T* __restrict__ B,
                                                              int M, int N, int K, float* __restrict__ C,
                                                              uint32_t fmt) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    __shared__ uint64_t mma_done;
    __shared__ uint32_t tmem_base_s;
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int kchunks = K / 64;
    uint8_t* sA = smem;
    uint8_t* sB = smem + (size_t)kchunks * 128 * 128;
    const int64_t m0 = (int64_t)blockIdx.x * 128;

    if (warp == 0) tc::tmem_alloc(&tmem_base_s, 256);
    if (tid == 0) {
        tc::mbar_init(&mma_done, 1);
        tc::mbar_fence_init();
    }
    // stage A: 128 rows x K halves; each thread copies 16-byte chunks
    for (int idx = tid; idx < 128 * (K / 8); idx += 128) {
        int r = idx / (K / 8), c = idx % (K / 8);
        int kc = c >> 3, c16 = c & 7;
        uint4 v = make_uint4(0, 0, 0, 0);
        if (m0 + r < M) v = *reinterpret_cast<const uint4*>(A + (m0 + r) * K + c * 8);
        *reinterpret_cast<uint4*>(sA + (size_t)kc * 128 * 128 + tc::sw128_offset(r, c16)) = v;
    }
    for (int idx = tid; idx < N * (K / 8); idx += 128) {
        int r = idx / (K / 8), c = idx % (K / 8);
        int kc = c >> 3, c16 = c & 7;
        uint4 v = *reinterpret_cast<const uint4*>(B + (int64_t)r * K + c * 8);
        *reinterpret_cast<uint4*>(sB + (size_t)kc * N * 128 + tc::sw128_offset(r, c16)) = v;
    }
    tc::fence_proxy_async_smem();
    tc::tc_fence_before_sync();
    __syncthreads();
    tc::tc_fence_after_sync();
    const uint32_t tmem_base = tmem_base_s;

    if (warp == 0 && tc::elect_one()) {
        const uint32_t idesc = tc::make_idesc(fmt, 128, N);
        for (int kc = 0; kc < kchunks; ++kc) {
            uint64_t adesc = tc::make_smem_desc_sw128(tc::smem_u32(sA + (size_t)kc * 128 * 128));
            uint64_t bdesc = tc::make_smem_desc_sw128(tc::smem_u32(sB + (size_t)kc * N * 128));
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                // one UMMA_K = 16 halves = 32 bytes = 2 units of 16 B inside the 128-byte swizzle row
                tc::umma_f16(tmem_base, adesc + 2 * k, bdesc + 2 * k, idesc, (kc | k) != 0);
            }
        }
        tc::umma_commit(&mma_done);
    }
    __syncwarp();
    tc::mbar_wait(&mma_done, 0);
    tc::tc_fence_after_sync();
    // epilogue: warp w owns TMEM lanes [32w, 32w+32) = rows m0 + 32w + lane
    const int64_t row = m0 + warp * 32 + lane;
    for (int n0 = 0; n0 < N; n0 += 32) {
        float v[32];
        tc::tmem_ld_32x32b_x32(tmem_base + ((uint32_t)(warp * 32) << 16) + n0, v);
        tc::tmem_ld_wait();
        if (row < M) {
#pragma unroll
            for (int j = 0; j < 32; ++j)
                if (n0 + j < N) C[row * N + n0 + j] = v[j];
        }
    }
    tc::tc_fence_before_sync();
    __syncthreads();
    if (warp == 0) tc::tmem_dealloc(tmem_base, 256);
}

// Same product with the A operand in TENSOR MEMORY (the `ts` MMA form): each thread packs its row of A (two 16-bit
// values per 32-bit word) and writes it with tcgen05.st; B as above.  Self-test for the layout the chain kernels
// would use to keep activations in TMEM.
template <typename T>
__global__ void __launch_bounds__(128, 1) tc_gemm_ts_test_kernel(const T* __restrict__ A, const T* __restrict__ B,
                                                                 int M, int N, int K, float* __restrict__ C,
                                                                 uint32_t fmt) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    __shared__ uint64_t mma_done;
    __shared__ uint32_t tmem_base_s;
    uint8_t* sB = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int kchunks = K / 64;
    const int64_t m0 = (int64_t)blockIdx.x * 128;
    if (warp == 0) tc::tmem_alloc(&tmem_base_s, 512);
    if (tid == 0) {
        tc::mbar_init(&mma_done, 1);
        tc::mbar_fence_init();
    }
    for (int idx = tid; idx < N * (K / 8); idx += 128) {
        int r = idx / (K / 8), c = idx % (K / 8);
        int kc = c >> 3, c16 = c & 7;
        uint4 v = *reinterpret_cast<const uint4*>(B + (int64_t)r * K + c * 8);
        *reinterpret_cast<uint4*>(sB + (size_t)kc * N * 128 + tc::sw128_offset(r, c16)) = v;
    }
    tc::fence_proxy_async_smem();
    tc::tc_fence_before_sync();
    __syncthreads();
    tc::tc_fence_after_sync();
    const uint32_t tmem_base = tmem_base_s;
    const uint32_t a_col = 256;                       // A occupies columns [256, 256 + K/2)
    {
        // thread = row (TMEM lane 32*warp + lane); 16 values -> 8 packed words per store
        const int64_t row = m0 + warp * 32 + lane;
        for (int k0 = 0; k0 < K; k0 += 16) {
            uint32_t w[8];
            if (row < M) {
                const uint4 a = *reinterpret_cast<const uint4*>(A + row * K + k0);
                const uint4 b = *reinterpret_cast<const uint4*>(A + row * K + k0 + 8);
                w[0] = a.x; w[1] = a.y; w[2] = a.z; w[3] = a.w; w[4] = b.x; w[5] = b.y; w[6] = b.z; w[7] = b.w;
            } else {
#pragma unroll
                for (int i = 0; i < 8; ++i) w[i] = 0;
            }
            tc::tmem_st_32x32b_x8(tmem_base + ((uint32_t)(warp * 32) << 16) + a_col + (uint32_t)(k0 / 2), w);
        }
        tc::tmem_st_wait();
    }
    tc::tc_fence_before_sync();
    __syncthreads();
    tc::tc_fence_after_sync();
    if (warp == 0 && tc::elect_one()) {
        const uint32_t idesc = tc::make_idesc(fmt, 128, N);
        for (int kc = 0; kc < kchunks; ++kc) {
            uint64_t bdesc = tc::make_smem_desc_sw128(tc::smem_u32(sB + (size_t)kc * N * 128));
#pragma unroll
            for (int k = 0; k < 4; ++k)
                tc::umma_f16_ts(tmem_base, tmem_base + a_col + (uint32_t)(kc * 32 + k * 8), bdesc + 2 * k, idesc, (kc | k) != 0);
        }
        tc::umma_commit(&mma_done);
    }
    __syncwarp();
    tc::mbar_wait(&mma_done, 0);
    tc::tc_fence_after_sync();
    const int64_t row = m0 + warp * 32 + lane;
    for (int n0 = 0; n0 < N; n0 += 32) {
        float v[32];
        tc::tmem_ld_32x32b_x32(tmem_base + ((uint32_t)(warp * 32) << 16) + n0, v);
        tc::tmem_ld_wait();
        if (row < M) {
#pragma unroll
            for (int j = 0; j < 32; ++j)
                if (n0 + j < N) C[row * N + n0 + j] = v[j];
        }
    }
    tc::tc_fence_before_sync();
    __syncthreads();
    if (warp == 0) tc::tmem_dealloc(tmem_base, 512);
}

}  // namespace hn

using namespace hn;

extern "C" int hn_tc_gemm_ts_test(const void* A, const void* B, int M, int N, int K, int is_bf16, float* C,
                                  hn_stream_t stream) {
    HN_REQUIRE(A && B && C, "hn_tc_gemm_ts_test: null pointer");
    HN_REQUIRE(M > 0 && N >= 16 && N <= 256 && N % 16 == 0 && K >= 64 && K % 64 == 0 && K <= 256,
               "hn_tc_gemm_ts_test: need 16 <= N <= 256 (N %% 16 == 0), K %% 64 == 0, K <= 256");
    size_t smem = (size_t)(K / 64) * N * 128 + 1024;
    auto kern = tc_gemm_ts_test_kernel<__half>;
    HN_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kern<<<(unsigned)ceil_div(M, 128), 128, smem, (cudaStream_t)stream>>>(
        (const __half*)A, (const __half*)B, M, N, K, C, is_bf16 ? tc::FMT_BF16 : tc::FMT_F16);
    count_launch();
    HN_CHECK_LAUNCH();
    return HN_OK;
}

extern "C" int hn_tc_gemm_test(const void* A, const void* B, int M, int N, int K, int is_bf16, float* C,
                               hn_stream_t stream) {
    HN_REQUIRE(A && B && C, "hn_tc_gemm_test: null pointer");
    HN_REQUIRE(M > 0 && N >= 16 && N <= 256 && N % 16 == 0 && K >= 64 && K % 64 == 0 && K <= 512,
               "hn_tc_gemm_test: need 16 <= N <= 256 (N %% 16 == 0), K %% 64 == 0, K <= 512");
    size_t smem = (size_t)(K / 64) * (128 + N) * 128 + 1024;
    HN_REQUIRE(smem <= 227 * 1024, "hn_tc_gemm_test: tile does not fit shared memory");
    auto kern = tc_gemm_test_kernel<__half>;
    HN_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kern<<<(unsigned)ceil_div(M, 128), 128, smem, (cudaStream_t)stream>>>(
        (const __half*)A, (const __half*)B, M, N, K, C, is_bf16 ? tc::FMT_BF16 : tc::FMT_F16);
    count_launch();
    HN_CHECK_LAUNCH();
    return HN_OK;
}

#include "../gemm_dispatch.cuh"

extern "C" int hn_gemm_test(int layout, int passes, int M, int N, int K, const float* A, int64_t lda,
                            const float* B, int64_t ldb, const float* bias, float* C, int64_t ldc,
                            hn_stream_t stream) {
    GemmArgs g;
    g.A = A; g.lda = lda; g.B = B; g.ldb = ldb; g.M = M; g.N = N; g.K = K; g.C = C; g.ldc = ldc; g.bias = bias;
    cudaStream_t s = (cudaStream_t)stream;
    const int splits = 37;
    if (layout == 0) {
        if (passes == 0) return launch_gemm<true, true, EPI_STORE>(g, s);
        if (passes == 1) return launch_gemm_tc<false, 1, EPI_STORE>(g, s);
        if (passes == 3) return launch_gemm_tc<false, 3, EPI_STORE>(g, s);
    } else if (layout == 1) {
        if (passes == 0) return launch_gemm<true, false, EPI_STORE>(g, s);
        if (passes == 1) return launch_gemm_tc<true, 1, EPI_STORE>(g, s);
        if (passes == 3) return launch_gemm_tc<true, 3, EPI_STORE>(g, s);
    } else if (layout == 2) {
        if (passes == 0) return launch_gemm<false, false, EPI_ATOMIC>(g, s, splits);
        if (passes == 1) return launch_gemm_tc_tn<1>(g, s, splits);
        if (passes == 3) return launch_gemm_tc_tn<3>(g, s, splits);
    }
    set_error("hn_gemm_test: unsupported layout/passes %d/%d", layout, passes);
    return HN_ERR_UNSUPPORTED;
}
