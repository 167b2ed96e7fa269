// Common helpers for the honerf_b200 CUDA library (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/honerf_b200.h"

namespace hn {

// Thread-local last error text, returned by hn_last_error().
void set_error(const char* fmt, ...);

#define HN_CHECK_CUDA(expr)                                                              \
    do {                                                                                 \
        cudaError_t _e = (expr);                                                         \
        if (_e != cudaSuccess) {                                                         \
            hn::set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #expr,                  \
                          cudaGetErrorString(_e));                                       \
            return HN_ERR_CUDA;                                                          \
        }                                                                                \
    } while (0)

#define HN_CHECK_LAUNCH() HN_CHECK_CUDA(cudaGetLastError())

#define HN_REQUIRE(cond, ...)                                                            \
    do {                                                                                 \
        if (!(cond)) {                                                                   \
            hn::set_error(__VA_ARGS__);                                                  \
            return HN_ERR_ARG;                                                           \
        }                                                                                \
    } while (0)

#define HN_PROPAGATE(expr)                                                               \
    do {                                                                                 \
        int _r = (expr);                                                                 \
        if (_r != HN_OK) return _r;                                                      \
    } while (0)

static inline int64_t ceil_div(int64_t a, int64_t b) { return (a + b - 1) / b; }
static inline int64_t round_up(int64_t a, int64_t b) { return ceil_div(a, b) * b; }
static inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

// nn.Softplus(beta=100), threshold 20 (utils/fields.py:125,310; SURVEY A-12)
__device__ __forceinline__ float softplus100(float z) {
    float t = z * 100.0f;
    return t > 20.0f ? z : log1pf(expf(t)) * 0.01f;
}
// softplus'(z) recovered from h = softplus(z):  sigmoid(100 z) = 1 - exp(-100 h)
__device__ __forceinline__ float sprime_from_h(float h) { return -expm1f(-100.0f * h); }
// 1 - softplus'(z) = exp(-100 h)
__device__ __forceinline__ float one_minus_sprime_from_h(float h) { return expf(-100.0f * h); }

__device__ __forceinline__ float sigmoidf_(float x) { return 1.0f / (1.0f + expf(-x)); }

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

int sm_count();
// bump the process-wide kernel-launch counter reported by hn_launch_count()
void count_launch(int n = 1);

// Opt-in device timing of the dense-contraction launches (hn_timing_* in the C ABI): when enabled,
// a pair of CUDA events is recorded on the launching stream around every GEMM launch.
enum TimingTag {
    TT_GEMM = 0,          // per-layer contractions (SIMT / TF32 paths, hand field)
    TT_SDF_ONLY = 1,      // chain: sampler's SDF queries
    TT_SDF_FWD = 2,       // chain: value + feature + normal sweep
    TT_SDF_BWD = 3,       // chain: tangent + reverse sweeps
    TT_DW = 4,            // chain: weight gradients
    TT_COLOR_FWD = 5,
    TT_COLOR_BWD = 6,
    TT_COUNT = 7
};
struct TimingScope {
    cudaStream_t stream;
    int slot;
    explicit TimingScope(cudaStream_t s, int tag = TT_GEMM);
    ~TimingScope();
};

}  // namespace hn
