// Shared device helpers for the field kernels: BARF-style encoding (utils/fields.py:13-20) of a
// 3-vector and its Jacobian products; column sums.
#pragma once
#include "common.cuh"

namespace hn {

// column j of [x(3), enc_L(x)(6L)]:  3 + c*2L + s*L + k  ->  (s ? cos : sin)(2^k x_c)
__device__ __forceinline__ float enc3_col(const float x[3], int L, int j) {
    if (j < 3) return x[j];
    int jj = j - 3;
    int c = jj / (2 * L);
    int r = jj - c * 2 * L;
    int s = r / L, k = r - s * L;
    float a = x[c] * (float)(1 << k);     // exact: power-of-two scaling, as input[...,None]*freq
    return s == 0 ? sinf(a) : cosf(a);
}

// (J_enc^T g)_c with the sin/cos values read back from the encoded vector e
// `stride`: distance between consecutive encoding columns of the point in e and g (1 = row-major rows)
__device__ __forceinline__ float enc3_jt_from_enc(const float* e, const float* g, int L, int c, int stride = 1) {
    float acc = g[c * stride];
    const float* es = e + (3 + c * 2 * L) * stride;
    const float* gs = g + (3 + c * 2 * L) * stride;
    float f = 1.0f;
    for (int k = 0; k < L; ++k) {
        acc += f * (es[(L + k) * stride] * gs[k * stride] - es[k * stride] * gs[(L + k) * stride]);
        f *= 2.0f;
    }
    return acc;
}

// out[c] += scale * sum_p X[p*ldx + c]   (atomic accumulation)
int launch_colsum(const float* X, int64_t ldx, int64_t rows, int cols, float scale, float* out,
                  cudaStream_t stream);

}  // namespace hn
