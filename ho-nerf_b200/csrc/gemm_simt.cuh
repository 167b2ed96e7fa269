// fp32 SIMT GEMM with fused epilogues -- the verification path (HN_SIMT_FP32) of every dense
// contraction in the field MLPs.  C[M,N] = epilogue( sum_k A(m,k) * B(k,n) ).
//
//   A_KMAJ : A(m,k) = A[m*lda + k]   (activations [points, features])   else A[k*lda + m]
//   B_KMAJ : B(k,n) = B[n*ldb + k]   (weights [out, in] used as x @ W^T) else B[k*ldb + n]
//
// 128x128x16 tiles, 256 threads, 8x8 outputs per thread, register-prefetched double buffering.
// Rows/cols/k beyond (M,N,K) are masked; every row must be readable up to round_up(len,4) floats
// (all buffers in this library have leading dimensions that are multiples of 4).
#pragma once
#include "common.cuh"

namespace hn {

enum Epi {
    EPI_STORE = 0,          // C = alpha*acc (+ bias[n])
    EPI_BIAS_SOFTPLUS = 1,  // C = softplus100(acc + bias[n])
    EPI_BIAS_RELU = 2,      // C = relu(acc + bias[n])
    EPI_BIAS_SIGMOID = 3,   // C = sigmoid(acc + bias[n])
    EPI_MUL_SPRIME = 4,     // n <  nsplit: C = acc * s'(aux1[m,n]);  n >= nsplit: C2[m,n-nsplit] = acc
    EPI_TANGENT = 5,        // q = acc; s1 = s'(aux1); C = s1*q; aux2/C2 (in place) = 100*(1-s1)*C2*q
    EPI_REVERSE = 6,        // n < nsplit: C = s'(aux1)*acc + aux2[m,n];  n >= nsplit: C2[m,n-nsplit] = acc
    EPI_RELU_BWD = 7,       // C = aux1[m,n] > 0 ? acc : 0
    EPI_ADD_AUX = 8,        // C = acc + aux1[m,n]
    EPI_ATOMIC = 9,         // atomicAdd(C, alpha*acc)   (split-K weight gradients)
};

struct GemmArgs {
    const float* A = nullptr; int64_t lda = 0;
    const float* B = nullptr; int64_t ldb = 0;
    const float* BT = nullptr; int64_t ldbt = 0;   // optional transposed copy of B (d @ W on tensor cores)
    // HN_TC_BF16X3 without a chain kernel: B (resp. BT) pre-packed as bf16 hi/lo tcgen05 tiles (gemm_bx3.cuh), 256-row
    // tiles `*_tile_bytes` apart, starting at k-block bp_kb0
    const uint8_t* Bp = nullptr; int64_t bp_tile_bytes = 0; int bp_kb0 = 0;
    const uint8_t* BTp = nullptr; int64_t btp_tile_bytes = 0;
    int M = 0, N = 0, K = 0;
    float* C = nullptr; int64_t ldc = 0;
    float* C2 = nullptr; int64_t ldc2 = 0;
    const float* bias = nullptr;
    const float* aux1 = nullptr; int64_t ldaux1 = 0;
    const float* aux2 = nullptr; int64_t ldaux2 = 0;
    float alpha = 1.0f;
    int nsplit = 1 << 30;   // column at which the epilogue switches to the C2 output
    int k_chunk = 0;        // split-K: K range per blockIdx.z (0 = whole K)
    int vec_ok = 0;         // set by launch_gemm: every epilogue pointer 16B aligned, every ld % 4 == 0
};

constexpr int GBM = 128, GBN = 128, GBK = 16, GPAD = 4, GTHREADS = 256;

template <int EPI>
__device__ __forceinline__ void epi_elem(const GemmArgs& g, int64_t m, int n, float acc) {
    if (EPI == EPI_STORE) {
        float v = g.alpha * acc;
        if (g.bias) v += g.bias[n];
        g.C[m * g.ldc + n] = v;
    } else if (EPI == EPI_BIAS_SOFTPLUS) {
        g.C[m * g.ldc + n] = softplus100(acc + g.bias[n]);
    } else if (EPI == EPI_BIAS_RELU) {
        g.C[m * g.ldc + n] = fmaxf(acc + g.bias[n], 0.0f);
    } else if (EPI == EPI_BIAS_SIGMOID) {
        g.C[m * g.ldc + n] = sigmoidf_(acc + g.bias[n]);
    } else if (EPI == EPI_MUL_SPRIME) {
        if (n < g.nsplit) g.C[m * g.ldc + n] = acc * sprime_from_h(g.aux1[m * g.ldaux1 + n]);
        else g.C2[m * g.ldc2 + (n - g.nsplit)] = acc;
    } else if (EPI == EPI_TANGENT) {
        float h = g.aux1[m * g.ldaux1 + n];
        float om = one_minus_sprime_from_h(h);
        float s1 = sprime_from_h(h);
        float d = g.C2[m * g.ldc2 + n];
        g.C[m * g.ldc + n] = s1 * acc;
        g.C2[m * g.ldc2 + n] = 100.0f * om * d * acc;
    } else if (EPI == EPI_REVERSE) {
        if (n < g.nsplit)
            g.C[m * g.ldc + n] = sprime_from_h(g.aux1[m * g.ldaux1 + n]) * acc + g.aux2[m * g.ldaux2 + n];
        else
            g.C2[m * g.ldc2 + (n - g.nsplit)] = acc;
    } else if (EPI == EPI_RELU_BWD) {
        g.C[m * g.ldc + n] = g.aux1[m * g.ldaux1 + n] > 0.0f ? acc : 0.0f;
    } else if (EPI == EPI_ADD_AUX) {
        g.C[m * g.ldc + n] = acc + g.aux1[m * g.ldaux1 + n];
    } else if (EPI == EPI_ATOMIC) {
        atomicAdd(&g.C[m * g.ldc + n], g.alpha * acc);
    }
}

__device__ __forceinline__ float4 ld4(const float* p) { return *reinterpret_cast<const float4*>(p); }
__device__ __forceinline__ void st4(float* p, float4 v) { *reinterpret_cast<float4*>(p) = v; }

// Vectorised epilogue on 4 consecutive columns n..n+3 (all < N, n % 4 == 0, not straddling nsplit).
template <int EPI>
__device__ __forceinline__ bool epi_vec4(const GemmArgs& g, int64_t m, int n, const float* a) {
    if (EPI == EPI_ATOMIC) return false;
    if ((EPI == EPI_MUL_SPRIME || EPI == EPI_REVERSE) && n + 3 >= g.nsplit) return false;
    float4 o;
    if (EPI == EPI_STORE) {
        o = make_float4(g.alpha * a[0], g.alpha * a[1], g.alpha * a[2], g.alpha * a[3]);
        if (g.bias) { o.x += g.bias[n]; o.y += g.bias[n + 1]; o.z += g.bias[n + 2]; o.w += g.bias[n + 3]; }
    } else if (EPI == EPI_BIAS_SOFTPLUS) {
        o = make_float4(softplus100(a[0] + g.bias[n]), softplus100(a[1] + g.bias[n + 1]),
                        softplus100(a[2] + g.bias[n + 2]), softplus100(a[3] + g.bias[n + 3]));
    } else if (EPI == EPI_BIAS_RELU) {
        o = make_float4(fmaxf(a[0] + g.bias[n], 0.f), fmaxf(a[1] + g.bias[n + 1], 0.f),
                        fmaxf(a[2] + g.bias[n + 2], 0.f), fmaxf(a[3] + g.bias[n + 3], 0.f));
    } else if (EPI == EPI_BIAS_SIGMOID) {
        o = make_float4(sigmoidf_(a[0] + g.bias[n]), sigmoidf_(a[1] + g.bias[n + 1]),
                        sigmoidf_(a[2] + g.bias[n + 2]), sigmoidf_(a[3] + g.bias[n + 3]));
    } else if (EPI == EPI_MUL_SPRIME) {
        float4 h = ld4(g.aux1 + m * g.ldaux1 + n);
        o = make_float4(a[0] * sprime_from_h(h.x), a[1] * sprime_from_h(h.y),
                        a[2] * sprime_from_h(h.z), a[3] * sprime_from_h(h.w));
    } else if (EPI == EPI_TANGENT) {
        float4 h = ld4(g.aux1 + m * g.ldaux1 + n);
        float4 d = ld4(g.C2 + m * g.ldc2 + n);
        o = make_float4(sprime_from_h(h.x) * a[0], sprime_from_h(h.y) * a[1],
                        sprime_from_h(h.z) * a[2], sprime_from_h(h.w) * a[3]);
        float4 x = make_float4(100.f * one_minus_sprime_from_h(h.x) * d.x * a[0],
                               100.f * one_minus_sprime_from_h(h.y) * d.y * a[1],
                               100.f * one_minus_sprime_from_h(h.z) * d.z * a[2],
                               100.f * one_minus_sprime_from_h(h.w) * d.w * a[3]);
        st4(g.C2 + m * g.ldc2 + n, x);
    } else if (EPI == EPI_REVERSE) {
        float4 h = ld4(g.aux1 + m * g.ldaux1 + n);
        float4 x = ld4(g.aux2 + m * g.ldaux2 + n);
        o = make_float4(sprime_from_h(h.x) * a[0] + x.x, sprime_from_h(h.y) * a[1] + x.y,
                        sprime_from_h(h.z) * a[2] + x.z, sprime_from_h(h.w) * a[3] + x.w);
    } else if (EPI == EPI_RELU_BWD) {
        float4 h = ld4(g.aux1 + m * g.ldaux1 + n);
        o = make_float4(h.x > 0.f ? a[0] : 0.f, h.y > 0.f ? a[1] : 0.f, h.z > 0.f ? a[2] : 0.f,
                        h.w > 0.f ? a[3] : 0.f);
    } else if (EPI == EPI_ADD_AUX) {
        float4 h = ld4(g.aux1 + m * g.ldaux1 + n);
        o = make_float4(a[0] + h.x, a[1] + h.y, a[2] + h.z, a[3] + h.w);
    }
    st4(g.C + m * g.ldc + n, o);
    return true;
}

template <bool A_KMAJ, bool B_KMAJ, int EPI>
__global__ void __launch_bounds__(GTHREADS, 2) gemm_simt_kernel(const GemmArgs g) {
    __shared__ __align__(16) float As[2][GBK][GBM + GPAD];
    __shared__ __align__(16) float Bs[2][GBK][GBN + GPAD];
    const int tid = threadIdx.x;
    const int tx = tid & 15, ty = tid >> 4;
    const int64_t m0 = (int64_t)blockIdx.x * GBM;
    const int n0 = blockIdx.y * GBN;
    int kbeg = 0, kend = g.K;
    if (g.k_chunk > 0) {
        kbeg = blockIdx.z * g.k_chunk;
        kend = min(g.K, kbeg + g.k_chunk);
        if (kbeg >= kend) return;
    }
    float acc[8][8];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = 0.0f;

    float4 ra[2], rb[2];
    auto load_tiles = [&](int k0) {
#pragma unroll
        for (int it = 0; it < 2; ++it) {
            int idx = tid + it * GTHREADS;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (A_KMAJ) {
                int mm = idx >> 2, kq = (idx & 3) * 4;
                int64_t gm = m0 + mm;
                int gk = k0 + kq;
                if (gm < g.M && gk < kend) {
                    v = ld4(g.A + gm * g.lda + gk);
                    if (gk + 1 >= kend) v.y = 0.f;
                    if (gk + 2 >= kend) v.z = 0.f;
                    if (gk + 3 >= kend) v.w = 0.f;
                }
            } else {
                int kk = idx >> 5, mq = (idx & 31) * 4;
                int64_t gm = m0 + mq;
                int gk = k0 + kk;
                if (gk < kend && gm < g.M) {
                    v = ld4(g.A + (int64_t)gk * g.lda + gm);
                    if (gm + 1 >= g.M) v.y = 0.f;
                    if (gm + 2 >= g.M) v.z = 0.f;
                    if (gm + 3 >= g.M) v.w = 0.f;
                }
            }
            ra[it] = v;
            v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (B_KMAJ) {
                int nn = idx >> 2, kq = (idx & 3) * 4;
                int gn = n0 + nn;
                int gk = k0 + kq;
                if (gn < g.N && gk < kend) {
                    v = ld4(g.B + (int64_t)gn * g.ldb + gk);
                    if (gk + 1 >= kend) v.y = 0.f;
                    if (gk + 2 >= kend) v.z = 0.f;
                    if (gk + 3 >= kend) v.w = 0.f;
                }
            } else {
                int kk = idx >> 5, nq = (idx & 31) * 4;
                int gn = n0 + nq;
                int gk = k0 + kk;
                if (gk < kend && gn < g.N) {
                    v = ld4(g.B + (int64_t)gk * g.ldb + gn);
                    if (gn + 1 >= g.N) v.y = 0.f;
                    if (gn + 2 >= g.N) v.z = 0.f;
                    if (gn + 3 >= g.N) v.w = 0.f;
                }
            }
            rb[it] = v;
        }
    };
    auto store_tiles = [&](int buf) {
#pragma unroll
        for (int it = 0; it < 2; ++it) {
            int idx = tid + it * GTHREADS;
            if (A_KMAJ) {
                int mm = idx >> 2, kq = (idx & 3) * 4;
                As[buf][kq + 0][mm] = ra[it].x; As[buf][kq + 1][mm] = ra[it].y;
                As[buf][kq + 2][mm] = ra[it].z; As[buf][kq + 3][mm] = ra[it].w;
            } else {
                int kk = idx >> 5, mq = (idx & 31) * 4;
                st4(&As[buf][kk][mq], ra[it]);
            }
            if (B_KMAJ) {
                int nn = idx >> 2, kq = (idx & 3) * 4;
                Bs[buf][kq + 0][nn] = rb[it].x; Bs[buf][kq + 1][nn] = rb[it].y;
                Bs[buf][kq + 2][nn] = rb[it].z; Bs[buf][kq + 3][nn] = rb[it].w;
            } else {
                int kk = idx >> 5, nq = (idx & 31) * 4;
                st4(&Bs[buf][kk][nq], rb[it]);
            }
        }
    };

    // k tiles start at a multiple of 4 (kbeg is a multiple of GBK by construction)
    load_tiles(kbeg);
    store_tiles(0);
    __syncthreads();
    int buf = 0;
    for (int k0 = kbeg; k0 < kend; k0 += GBK) {
        const bool has_next = k0 + GBK < kend;
        if (has_next) load_tiles(k0 + GBK);
#pragma unroll
        for (int kk = 0; kk < GBK; ++kk) {
            float4 a0 = ld4(&As[buf][kk][ty * 4]);
            float4 a1 = ld4(&As[buf][kk][64 + ty * 4]);
            float4 b0 = ld4(&Bs[buf][kk][tx * 4]);
            float4 b1 = ld4(&Bs[buf][kk][64 + tx * 4]);
            const float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
            const float b[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
                for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
        }
        if (has_next) {
            store_tiles(buf ^ 1);
            __syncthreads();
            buf ^= 1;
        }
    }

    // epilogue
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        int64_t m = m0 + (i < 4 ? ty * 4 + i : 64 + ty * 4 + (i - 4));
        if (m >= g.M) continue;
#pragma unroll
        for (int jh = 0; jh < 2; ++jh) {
            int n = n0 + jh * 64 + tx * 4;
            if (n >= g.N) continue;
            const float* a = &acc[i][jh * 4];
            bool done = false;
            if (g.vec_ok && n + 3 < g.N) done = epi_vec4<EPI>(g, m, n, a);
            if (!done) {
#pragma unroll
                for (int j = 0; j < 4; ++j)
                    if (n + j < g.N) epi_elem<EPI>(g, m, n + j, a[j]);
            }
        }
    }
}


template <bool A_KMAJ, bool B_KMAJ, int EPI>
int launch_gemm(const GemmArgs& g, cudaStream_t stream, int k_splits = 1) {
    if (g.M <= 0 || g.N <= 0 || g.K <= 0) return HN_OK;
    GemmArgs a = g;
    HN_REQUIRE(g.A && g.B && g.C, "gemm: null operand");
    HN_REQUIRE(aligned16(g.A) && aligned16(g.B) && g.lda % 4 == 0 && g.ldb % 4 == 0,
               "gemm: operands must be 16B aligned with leading dimensions that are multiples of 4");
    auto ok = [](const void* p, int64_t ld) { return p == nullptr || (aligned16(p) && ld % 4 == 0); };
    a.vec_ok = ok(g.C, g.ldc) && ok(g.C2, g.ldc2) && ok(g.aux1, g.ldaux1) && ok(g.aux2, g.ldaux2) &&
               (g.bias == nullptr || aligned16(g.bias) || true);
    dim3 grid((unsigned)ceil_div(g.M, GBM), (unsigned)ceil_div(g.N, GBN), 1);
    if (k_splits > 1) {
        int chunk = (int)round_up(ceil_div(g.K, k_splits), GBK);
        a.k_chunk = chunk;
        grid.z = (unsigned)ceil_div(g.K, chunk);
    }
    {
        TimingScope ts(stream);
        gemm_simt_kernel<A_KMAJ, B_KMAJ, EPI><<<grid, GTHREADS, 0, stream>>>(a);
    }
    count_launch();
    HN_CHECK_LAUNCH();
    return HN_OK;
}

}  // namespace hn
