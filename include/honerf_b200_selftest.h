/* Self-test entry points of the tcgen05 plumbing (descriptors, TMEM, mbarriers) and of the per-layer contraction
 * kernels: built into libhonerf_b200_selftest.so from csrc/selftest/, loaded by tests/test_gpu_tc.py only.  NOT part of
 * the product library (libhonerf_b200.so exports nothing from here). */
#ifndef HONERF_B200_SELFTEST_H
#define HONERF_B200_SELFTEST_H
#include "honerf_b200.h"
#ifdef __cplusplus
extern "C" {
#endif

/* ---------------------------------------------------------------------------------------------
 * Self-test: C[M,N] (fp32) = A[M,K] * B[N,K]^T with fp16 (or bf16) operands on tcgen05 tensor cores
 * (fp32 accumulation in TMEM).  Self-test of the descriptors / TMEM / mbarrier plumbing shared by the
 * fused field kernels.  16 <= N <= 256, N % 16 == 0, K % 64 == 0.
 * ------------------------------------------------------------------------------------------- */
HN_API int hn_tc_gemm_test(const void* A, const void* B, int M, int N, int K, int is_bf16, float* C,
                           hn_stream_t stream);
/* Same product with the A operand staged in tensor memory (tcgen05.st + the `ts` MMA form); K <= 256. */
HN_API int hn_tc_gemm_ts_test(const void* A, const void* B, int M, int N, int K, int is_bf16, float* C,
                              hn_stream_t stream);
/* One dense contraction through the production kernels, for tests: C [M, ldc] (fp32).
 *   layout 0: C = A[M,lda] @ B[N,ldb]^T (+ bias[N])      layout 1: C = A[M,lda] @ B[K,ldb]
 *   layout 2: C += A[K,lda]^T @ B[K,ldb]  (K split over CTAs, atomics; caller zeroes C)
 *   passes 0: fp32 SIMT, 1: tcgen05 TF32, 3: tcgen05 split TF32. */
HN_API int hn_gemm_test(int layout, int passes, int M, int N, int K, const float* A, int64_t lda,
                        const float* B, int64_t ldb, const float* bias, float* C, int64_t ldc,
                        hn_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* HONERF_B200_SELFTEST_H */
