// Object SDF field as fused tile-chain kernels on tcgen05 (HN_TC_BF16X3), see chain_common.cuh.
// Replaces, per 128-point tile and without leaving the SM between layers,
//   SDFNetwork_OBJ.sdf       (utils/fields.py:316-331)          -> sdf_only_kernel
// The positional encoding (utils/fields.py:13-20) is computed by the epilogue warps straight into the
// first layer's A operand; the skip connection's [h3 | e] / sqrt2 (utils/fields.py:323-324) is formed in
// place (1/sqrt2 is folded into W_4 at pack time, SURVEY A-8).
#include <algorithm>
#include <cstdlib>

#include "chain_common.cuh"
#include "chain_dw.cuh"
#include "chain_obj_layout.cuh"
#include "fields_common.cuh"

namespace hn {
namespace chain {

// ------------------------------------------------------------------------------------------------
// epilogue helpers
// ------------------------------------------------------------------------------------------------

// [x(3), sin/cos(2^k x_c)] of one point written as columns shift + j of the A operand; the column
// groups of a row share the 30 (coordinate, frequency) pairs.  Column shift+63 (the K padding of the
// first layer) is zeroed when shift == 0.
// `g` (may be NULL): fp32 copy for the stash: the point's entry of a column-major [64][128] tile (g[128 j] = e_j,
// see eoff) when !g_tiled, else the base of a tiled [128, 256] tile whose columns shift + j receive e_j.
template <bool LO16 = false>
__device__ __forceinline__ void write_encoding(uint8_t* smem, int row, int cg, const float x[3], int shift,
                                               float* __restrict__ g = nullptr, bool g_tiled = false) {
    auto gput = [&](int j, float v) {
        if (g) g[g_tiled ? toff(row, shift + j) : j * TILE_M] = v;
    };
    if (cg == 0) {
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            a_store1<LO16>(smem, row, shift + c, x[c]);
            gput(c, x[c]);
        }
        if (shift == 0) {
            a_store1<LO16>(smem, row, 63, 0.0f);
            gput(63, 0.0f);
        }
    }
    for (int idx = cg; idx < 30; idx += EPI_CGROUPS) {
        const int c = idx / 10, k = idx - c * 10;
        float s, co;
        sincosf(x[c] * (float)(1 << k), &s, &co);
        a_store1<LO16>(smem, row, shift + 3 + c * 20 + k, s);
        a_store1<LO16>(smem, row, shift + 3 + c * 20 + 10 + k, co);
        gput(3 + c * 20 + k, s);
        gput(3 + c * 20 + 10 + k, co);
    }
}

struct SdfOnlyParams {
    const float* pts;
    int64_t n;
    float inv_scale;
    float* sdf;
    const uint8_t* chain;
    const float* bias[9];
    const float* w_out0;      // fp32 row 0 of the packed output layer (the sdf row)
    int n_tiles;
    long long* prof;          // optional cycle counters [grid][4]
};

__global__ void __launch_bounds__(THREADS, 1)
sdf_only_kernel(const __grid_constant__ SdfOnlyParams p, const __grid_constant__ Program prog) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    __shared__ Barriers bar;
    uint8_t* smem = chain_setup(smem_raw, &bar);
    // per-column-group partial sums of the sdf head live in the last A k-block, which nobody touches between
    // the last layer's MMAs and the next tile's first epilogue
    float (*s_head)[TILE_M] = reinterpret_cast<float (*)[TILE_M]>(smem + 3 * KB_BYTES);
    const int warp = threadIdx.x >> 5;
    const int n_my_tiles = p.n_tiles > (int)blockIdx.x ? (p.n_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x : 0;
    if (warp == 0) {
        producer_loop(prog, p.chain, smem, &bar, n_my_tiles);
    } else if (warp == 1) {
        mma_loop(prog, smem, &bar, n_my_tiles, p.prof);
    } else {
        int row, cg;
        epi_coords(row, cg);
        const uint32_t tmem = bar.tmem_base;
        uint32_t acc_par = 0;
        long long t_acc = 0;
        for (int t = 0; t < n_my_tiles; ++t) {
            const int64_t tile = (int64_t)blockIdx.x + (int64_t)t * gridDim.x;
            const int64_t gp = tile * TILE_M + row;
            float x[3] = {0.f, 0.f, 0.f};
            if (gp < p.n) { x[0] = p.pts[gp * 3]; x[1] = p.pts[gp * 3 + 1]; x[2] = p.pts[gp * 3 + 2]; }
            write_encoding<true>(smem, row, cg, x, 0);
            epi_publish_a(&bar);
            float head = 0.0f;
            for (int l = 0; l < 8; ++l) {
                const long long tw = clock64();
                epi_wait_acc(&bar, acc_par);
                t_acc += clock64() - tw;
                const float* __restrict__ bias = p.bias[l];
                const bool skip_tail = l == 3 && cg == EPI_CGROUPS - 1 && EPI_COLS <= 64;   // columns >= 192: below
                if (!skip_tail) {
#pragma unroll
                    for (int blk = 0; blk < EPI_COLS / 32; ++blk) {
                        const int col0 = cg * EPI_COLS + blk * 32;
                        float v[32];
                        acc_load32(tmem, row, col0, v);
#pragma unroll
                        for (int j = 0; j < 32; j += 4) {
                            const float4 b = __ldg(reinterpret_cast<const float4*>(bias + col0 + j));
                            v[j] = softplus100_fast(v[j] + b.x);
                            v[j + 1] = softplus100_fast(v[j + 1] + b.y);
                            v[j + 2] = softplus100_fast(v[j + 2] + b.z);
                            v[j + 3] = softplus100_fast(v[j + 3] + b.w);
                        }
                        if (l < 7) {
#pragma unroll
                            for (int j = 0; j < 32; j += 8) a_store8<true>(smem, row, col0 + j, v + j);
                        } else {
#pragma unroll
                            for (int j = 0; j < 32; j += 4) {
                                const float4 w = __ldg(reinterpret_cast<const float4*>(p.w_out0 + col0 + j));
                                head += v[j] * w.x + v[j + 1] * w.y + v[j + 2] * w.z + v[j + 3] * w.w;
                            }
                        }
                    }
                }
                if (l == 3) {
                    // skip input columns 192..255 = [h3[192], e(63)]
                    if (cg == EPI_CGROUPS - 1) {
                        float v[32];
                        acc_load32(tmem, row, 192, v);
                        a_store1<true>(smem, row, 192, softplus100_fast(v[0] + __ldg(bias + 192)));
                    }
                    write_encoding<true>(smem, row, cg, x, 193);
                }
                if (l < 7) epi_publish_a(&bar);
            }
            // sdf = (h7 . W_out[0] + b_out[0]) / scale
            s_head[cg][row] = head;
            tc::named_bar_sync(1, EPI_THREADS);
            if (cg == 0 && gp < p.n) {
                float acc = 0.0f;
#pragma unroll
                for (int g = 0; g < EPI_CGROUPS; ++g) acc += s_head[g][row];
                p.sdf[gp] = (acc + __ldg(p.bias[8])) * p.inv_scale;
            }
        }
        if (p.prof && threadIdx.x == 64) p.prof[blockIdx.x * 4 + 3] = t_acc;
    }
    chain_teardown(&bar);
}

// ------------------------------------------------------------------------------------------------
// value + feature + analytic normal (SDFNetwork_OBJ.forward + .gradient, utils/fields.py:316-347) with
// the stash of hn_sdf_obj_bwd: 17 chained layers per tile -- value trunk (8), feature head (1), normal
// sweep (7 + the 64-wide encoding layer).
// ------------------------------------------------------------------------------------------------
struct FwdParams {
    const float* pts;
    int64_t n;
    float inv_scale;
    float* sdf;
    float* feat;
    int64_t ld_feat;
    float* normal;
    float* E;          // stash pieces, same layout as the per-layer path (fields_obj.cu: ObjSdfStash)
    float* H[8];
    float* D[8];
    float* EB;
    const uint8_t* chain;
    const float* bias[9];
    const float* w_out0;
    int n_tiles;
    int stagger;
    long long* prof;
};

__device__ __forceinline__ float sprime_fast(float h) { return 1.0f - ex2_approx(-144.26950408889634f * h); }

__global__ void __launch_bounds__(THREADS, 1)
sdf_fwd_kernel(const __grid_constant__ FwdParams p, const __grid_constant__ Program prog) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    __shared__ Barriers bar;
    uint8_t* smem = chain_setup(smem_raw, &bar);
    const int warp = threadIdx.x >> 5;
    const int n_my_tiles = p.n_tiles > (int)blockIdx.x ? (p.n_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x : 0;
    if (warp == 0) {
        producer_loop(prog, p.chain, smem, &bar, n_my_tiles);
    } else if (warp == 1) {
        mma_loop(prog, smem, &bar, n_my_tiles, p.prof);
    } else {
        int row, cg;
        epi_coords(row, cg);
        const uint32_t tmem = bar.tmem_base;
        uint32_t acc_par = 0;
        stagger_start(p.stagger, 4);
        for (int t = 0; t < n_my_tiles; ++t) {
            const int64_t tile = (int64_t)blockIdx.x + (int64_t)t * gridDim.x;
            const int64_t gp = tile * TILE_M + row;
            const bool live = gp < p.n;
            float x[3] = {0.f, 0.f, 0.f};
            if (live) { x[0] = p.pts[gp * 3]; x[1] = p.pts[gp * 3 + 1]; x[2] = p.pts[gp * 3 + 2]; }
            write_encoding<true>(smem, row, cg, x, 0, live ? p.E + eoff(gp) : nullptr);
            epi_publish_a(&bar);
            // ---- value trunk ----------------------------------------------------------------------
            float head = 0.0f;
            for (int l = 0; l < 8; ++l) {
                epi_wait_acc(&bar, acc_par);
                const float* __restrict__ bias = p.bias[l];
                float* __restrict__ ht = p.H[l] + tile * TILE_FLOATS;       // tiled [128, 256]
                const bool skip_tail = l == 3 && cg == EPI_CGROUPS - 1;      // columns >= 192: below
                if (!skip_tail) {
#pragma unroll
                    for (int blk = 0; blk < EPI_COLS / 32; ++blk) {
                        const int col0 = cg * EPI_COLS + blk * 32;
                        float v[32];
                        acc_load32(tmem, row, col0, v);
#pragma unroll
                        for (int j = 0; j < 32; j += 4) {
                            const float4 b = __ldg(reinterpret_cast<const float4*>(bias + col0 + j));
                            v[j] = softplus100_fast(v[j] + b.x);
                            v[j + 1] = softplus100_fast(v[j + 1] + b.y);
                            v[j + 2] = softplus100_fast(v[j + 2] + b.z);
                            v[j + 3] = softplus100_fast(v[j + 3] + b.w);
                            if (live) st4(ht + toff(row, col0 + j), make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]));
                        }
#pragma unroll
                        for (int j = 0; j < 32; j += 8) a_store8<true>(smem, row, col0 + j, v + j);
                        if (l == 7) {
#pragma unroll
                            for (int j = 0; j < 32; j += 4) {
                                const float4 w = __ldg(reinterpret_cast<const float4*>(p.w_out0 + col0 + j));
                                head += v[j] * w.x + v[j + 1] * w.y + v[j + 2] * w.z + v[j + 3] * w.w;
                            }
                        }
                    }
                }
                if (l == 3) {
                    // skip input, columns 192..255 = [h3[192], e_0 .. e_62]: 16 per column group, vector stores; e comes
                    // back from the fp32 encoding stash written at the start of the tile (no second sincosf)
                    const int c0 = 192 + cg * 16;
                    float tl[16];
#pragma unroll
                    for (int j = 0; j < 16; ++j) tl[j] = (live && c0 + j >= 193) ? p.E[eoff(gp, c0 + j - 193)] : 0.0f;
                    if (cg == 0) {
                        float v[32];
                        acc_load32(tmem, row, 192, v);
                        tl[0] = softplus100_fast(v[0] + __ldg(bias + 192));
                    }
                    a_store8<true>(smem, row, c0, tl);
                    a_store8<true>(smem, row, c0 + 8, tl + 8);
                    if (live) {
#pragma unroll
                        for (int j = 0; j < 16; j += 4) st4(ht + toff(row, c0 + j), make_float4(tl[j], tl[j + 1], tl[j + 2], tl[j + 3]));
                    }
                }
                epi_publish_a(&bar);
            }
            // sdf = (h7 . W_out[0] + b_out[0]) / scale; the 4 column groups meet through the (still unused) EB rows
            if (live) p.EB[eoff(gp, cg)] = head;
            tc::named_bar_sync(1, EPI_THREADS);
            if (cg == 0 && live) {
                const float* __restrict__ h4 = p.EB + eoff(gp);
                p.sdf[gp] = (h4[0] + h4[TILE_M] + h4[2 * TILE_M] + h4[3 * TILE_M] + __ldg(p.bias[8])) * p.inv_scale;
            }
            // ---- feature head; seed of the normal sweep D7 = s'(h7) * W_out[0] / scale ----------------
            epi_wait_acc(&bar, acc_par);
            {
                const float* __restrict__ bias = p.bias[8] + 1;
#pragma unroll
                for (int blk = 0; blk < EPI_COLS / 32; ++blk) {
                    const int col0 = cg * EPI_COLS + blk * 32;
                    float v[32];
                    acc_load32(tmem, row, col0, v);
                    if (live) {
#pragma unroll
                        for (int j = 0; j < 32; j += 4)
                            st4(p.feat + gp * p.ld_feat + col0 + j,
                                make_float4(v[j] + __ldg(bias + col0 + j), v[j + 1] + __ldg(bias + col0 + j + 1),
                                            v[j + 2] + __ldg(bias + col0 + j + 2), v[j + 3] + __ldg(bias + col0 + j + 3)));
                    }
#pragma unroll
                    for (int j = 0; j < 32; j += 8) {
                        float h[8];
                        a_load8<true>(smem, row, col0 + j, h);
                        const float4 w0 = __ldg(reinterpret_cast<const float4*>(p.w_out0 + col0 + j));
                        const float4 w1 = __ldg(reinterpret_cast<const float4*>(p.w_out0 + col0 + j + 4));
                        const float w[8] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w};
#pragma unroll
                        for (int i = 0; i < 8; ++i) h[i] = sprime_fast(h[i]) * w[i] * p.inv_scale;
                        a_store8(smem, row, col0 + j, h);
                        if (live) {
                            float* __restrict__ dt = p.D[7] + tile * TILE_FLOATS;
                            st4(dt + toff(row, col0 + j), make_float4(h[0], h[1], h[2], h[3]));
                            st4(dt + toff(row, col0 + j + 4), make_float4(h[4], h[5], h[6], h[7]));
                        }
                    }
                }
            }
            epi_publish_a(&bar);
            // ---- normal sweep: D_{l-1} = s'(h_{l-1}) * (D_l W_l), l = 7..1 ------------------------------
            for (int l = 7; l >= 1; --l) {
                const float* __restrict__ ht = p.H[l - 1] + tile * TILE_FLOATS;
                float* __restrict__ dt = p.D[l - 1] + tile * TILE_FLOATS;
                epi_stream<1, 8>(&bar, acc_par, tmem, row, cg, live, true, ht, ht, [&](int col0, float* v, float4 (*aux)[2]) {
#pragma unroll
                    for (int q = 0; q < 2; ++q) {
                        const int j = q * 4;
                        const float4 h = aux[0][q];
                        float d[4] = {v[j] * sprime_fast(h.x), v[j + 1] * sprime_fast(h.y), v[j + 2] * sprime_fast(h.z),
                                      v[j + 3] * sprime_fast(h.w)};
                        float g4[4] = {d[0], d[1], d[2], d[3]};
                        if (l == 4 && col0 + j + 3 > 192) {
                            // columns 193..255 of the skip layer's input are the encoding: their cotangent (the raw product)
                            // rides in the otherwise unused columns 193..255 of the D_3 tile until the last step
#pragma unroll
                            for (int i = 0; i < 4; ++i) {
                                if (col0 + j + i > 192) {
                                    g4[i] = v[j + i];
                                    d[i] = 0.0f;
                                }
                            }
                        }
                        v[j] = d[0]; v[j + 1] = d[1]; v[j + 2] = d[2]; v[j + 3] = d[3];
                        if (live) st4(dt + toff(row, col0 + j), make_float4(g4[0], g4[1], g4[2], g4[3]));
                    }
                    a_store8(smem, row, col0, v);
                });
                epi_publish_a(&bar);
            }
            // ---- encoding layer: eb = D_0 W_0 + (skip part); normal = J_e^T eb ----------------------------
            epi_wait_acc(&bar, acc_par);
            {
                float v[16];
                tc::tmem_ld_32x32b_x16(tmem + ((uint32_t)(row & ~31) << 16) + (uint32_t)(cg * 16), v);
                tc::tmem_ld_wait();
                if (live) {
                    const float* __restrict__ d3 = p.D[3] + tile * TILE_FLOATS;      // skip part: columns 193 + j
#pragma unroll
                    for (int j = 0; j < 16; j += 4) {
                        const int col = cg * 16 + j;
                        float4 e;
                        e.x = v[j] + d3[toff(row, 193 + col)];
                        e.y = v[j + 1] + d3[toff(row, 194 + col)];
                        e.z = v[j + 2] + d3[toff(row, 195 + col)];
                        e.w = col + 3 == 63 ? 0.0f : v[j + 3] + d3[toff(row, 196 + col)];
                        float* __restrict__ eb = p.EB + eoff(gp, col);
                        eb[0] = e.x; eb[TILE_M] = e.y; eb[2 * TILE_M] = e.z; eb[3 * TILE_M] = e.w;
                    }
                }
            }
            tc::tc_fence_before_sync();
            tc::named_bar_sync(1, EPI_THREADS);
            if (cg < 3 && live) {
                const float* __restrict__ e = p.E + eoff(gp, 3 + cg * 20);
                const float* __restrict__ g = p.EB + eoff(gp, 3 + cg * 20);
                float acc = p.EB[eoff(gp, cg)];
                float f = 1.0f;
#pragma unroll
                for (int k = 0; k < 10; ++k) {
                    acc += f * (e[(10 + k) * TILE_M] * g[k * TILE_M] - e[k * TILE_M] * g[(10 + k) * TILE_M]);
                    f *= 2.0f;
                }
                p.normal[gp * 3 + cg] = acc;
            }
        }
    }
    chain_teardown(&bar);
}

// ------------------------------------------------------------------------------------------------
// Second-order backward of (sdf, feature, normal) (the Hessian-vector products the reference gets from
// autograd.grad(create_graph=True), utils/fields.py:336-347): tangent sweep along d_normal (8 layers) then
// reverse sweep (output layer + 7 hidden + the encoding layer), 17 chained layers per tile.  Produces
// d_pts and, in the workspace, the operands of the weight-gradient contractions
//   dW_l = DZ_l^T a_{l-1} + D_l^T u_{l-1}    (a = value activations H, u = tangent activations U).
// ------------------------------------------------------------------------------------------------
struct BwdParams {
    int64_t n;
    float inv_scale;
    const float* E;       // stash of the forward
    const float* H[8];
    const float* D[8];
    const float* EB;
    const float* d_sdf;   // may be NULL
    const float* d_feat;  // may be NULL
    int64_t ld_dfeat;
    const float* d_normal;
    float* d_pts;         // may be NULL
    float* UE;            // workspace
    float* U[8];
    float* X[8];
    float* DZ[8];
    float* DE;
    const uint8_t* chain;
    const float* w_out0;
    int n_tiles;
    int stagger;
    long long* prof;
};

__global__ void __launch_bounds__(THREADS, 1)
sdf_bwd_kernel(const __grid_constant__ BwdParams p, const __grid_constant__ Program prog) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    __shared__ Barriers bar;
    uint8_t* smem = chain_setup(smem_raw, &bar);
    const int warp = threadIdx.x >> 5;
    const int n_my_tiles = p.n_tiles > (int)blockIdx.x ? (p.n_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x : 0;
    if (warp == 0) {
        producer_loop(prog, p.chain, smem, &bar, n_my_tiles);
    } else if (warp == 1) {
        mma_loop(prog, smem, &bar, n_my_tiles, p.prof);
    } else {
        int row, cg;
        epi_coords(row, cg);
        const uint32_t tmem = bar.tmem_base;
        uint32_t acc_par = 0;
        stagger_start(p.stagger, 4);
        for (int t = 0; t < n_my_tiles; ++t) {
            const int64_t tile = (int64_t)blockIdx.x + (int64_t)t * gridDim.x;
            const int64_t gp = tile * TILE_M + row;
            const bool live = gp < p.n;
            float dn[3] = {0.f, 0.f, 0.f};
            if (live) { dn[0] = p.d_normal[gp * 3]; dn[1] = p.d_normal[gp * 3 + 1]; dn[2] = p.d_normal[gp * 3 + 2]; }
            // ue = J_e(x) dn: tangent of the encoding, columns shift + j of the A operand (and of `g`)
            auto write_ue = [&](int shift, float* __restrict__ g, bool gt) {
                const float* __restrict__ e = p.E + eoff(gp);
                if (cg == 0) {
#pragma unroll
                    for (int c = 0; c < 3; ++c) {
                        a_store1(smem, row, shift + c, dn[c]);
                        if (live) g[gt ? toff(row, shift + c) : c * TILE_M] = dn[c];
                    }
                    if (shift == 0) {
                        a_store1(smem, row, 63, 0.0f);
                        if (live) g[63 * TILE_M] = 0.0f;
                    }
                }
                for (int idx = cg; idx < 30; idx += EPI_CGROUPS) {
                    const int c = idx / 10, k = idx - c * 10;
                    const float f = (float)(1 << k);
                    float sn = 0.f, cs = 0.f;
                    if (live) { sn = e[(3 + c * 20 + k) * TILE_M]; cs = e[(3 + c * 20 + 10 + k) * TILE_M]; }
                    const float us = f * cs * dn[c], uc = -f * sn * dn[c];
                    a_store1(smem, row, shift + 3 + c * 20 + k, us);
                    a_store1(smem, row, shift + 3 + c * 20 + 10 + k, uc);
                    if (live) {
                        g[gt ? toff(row, shift + 3 + c * 20 + k) : (3 + c * 20 + k) * TILE_M] = us;
                        g[gt ? toff(row, shift + 3 + c * 20 + 10 + k) : (3 + c * 20 + 10 + k) * TILE_M] = uc;
                    }
                }
            };
            write_ue(0, p.UE + eoff(gp), false);
            epi_publish_a(&bar);
            // ---- tangent sweep: q_l = W_l u_{l-1}; u_l = s'(h_l) q_l; X_l = 100 (1 - s') D_l q_l -----------
            for (int l = 0; l < 8; ++l) {
                const float* __restrict__ ht = p.H[l] + tile * TILE_FLOATS;
                const float* __restrict__ dt = p.D[l] + tile * TILE_FLOATS;
                float* __restrict__ ut = p.U[l] + tile * TILE_FLOATS;
                float* __restrict__ xt = p.X[l] + tile * TILE_FLOATS;
                const bool skip_tail = l == 3 && cg == EPI_CGROUPS - 1;
                epi_stream<2, 8>(&bar, acc_par, tmem, row, cg, live, !skip_tail, ht, dt, [&](int col0, float* v, float4 (*aux)[2]) {
#pragma unroll
                    for (int q = 0; q < 2; ++q) {
                        const int j = q * 4;
                        const float4 h = aux[0][q], d = aux[1][q];
                        const float hh[4] = {h.x, h.y, h.z, h.w}, dd[4] = {d.x, d.y, d.z, d.w};
                        float xx[4];
#pragma unroll
                        for (int i = 0; i < 4; ++i) {
                            const float qv = v[j + i];
                            const float em = ex2_approx(-144.26950408889634f * hh[i]);       // 1 - s'
                            v[j + i] = (1.0f - em) * qv;
                            xx[i] = 100.0f * em * dd[i] * qv;
                        }
                        if (live) {
                            st4(ut + toff(row, col0 + j), make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]));
                            st4(xt + toff(row, col0 + j), make_float4(xx[0], xx[1], xx[2], xx[3]));
                        }
                    }
                    if (l < 7) a_store8(smem, row, col0, v);
                });
                if (l == 3) {
                    // tangent of the skip input, columns 192..255 = [u3[192], ue_0 .. ue_62]: 16 per column group; ue comes
                    // back from the workspace written at the start of the tile
                    const int c0 = 192 + cg * 16;
                    float tl[16];
#pragma unroll
                    for (int j = 0; j < 16; ++j) tl[j] = (live && c0 + j >= 193) ? p.UE[eoff(gp, c0 + j - 193)] : 0.0f;
                    if (cg == 0) {
                        float v[32];
                        acc_load32(tmem, row, 192, v);
                        float u = 0.f;
                        if (live) {
                            const float em = ex2_approx(-144.26950408889634f * ht[toff(row, 192)]);
                            u = (1.0f - em) * v[0];
                            xt[toff(row, 192)] = 100.0f * em * dt[toff(row, 192)] * v[0];
                        }
                        tl[0] = u;
                    }
                    a_store8(smem, row, c0, tl);
                    a_store8(smem, row, c0 + 8, tl + 8);
                    if (live) {
#pragma unroll
                        for (int j = 0; j < 16; j += 4) st4(ut + toff(row, c0 + j), make_float4(tl[j], tl[j + 1], tl[j + 2], tl[j + 3]));
                    }
                }
                if (l == 7) {
                    // A operand of the output layer's reverse step: the tile's rows of d_feat (row-major).  A plain copy, so
                    // the threads are re-mapped for coalescing: consecutive threads take consecutive float4 of a row.
                    tc::named_bar_sync(1, EPI_THREADS);        // every thread is done with its own A stores of this step
                    const int et = threadIdx.x - 64;
#pragma unroll 4
                    for (int it = 0; it < TILE_M * 64 / EPI_THREADS; ++it) {
                        const int idx = it * EPI_THREADS + et;
                        const int r = idx >> 6, c = (idx & 63) * 4;
                        const int64_t g = tile * TILE_M + r;
                        float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
                        if (g < p.n && p.d_feat) a = ld4(p.d_feat + g * p.ld_dfeat + c);
                        uint2 hi, lo;
                        split2(a.x, a.y, hi.x, lo.x);
                        split2(a.z, a.w, hi.y, lo.y);
                        const uint32_t off = (uint32_t)(c >> 6) * KB_BYTES + tc::sw128_offset((uint32_t)r, (uint32_t)((c & 63) >> 3)) +
                                             (uint32_t)(c & 4) * 2u;
                        *reinterpret_cast<uint2*>(smem + off) = hi;
                        *reinterpret_cast<uint2*>(smem + A_LO_OFF + off) = lo;
                    }
                }
                epi_publish_a(&bar);
            }
            // ---- reverse sweep: dz_{l-1} = s'(h_{l-1}) (dz_l W_l) + X_{l-1}, l = 8..1 ------------------------
            const float gs = (live && p.d_sdf) ? p.d_sdf[gp] * p.inv_scale : 0.0f;
            for (int l = 8; l >= 1; --l) {
                const float* __restrict__ ht = p.H[l - 1] + tile * TILE_FLOATS;
                const float* __restrict__ xt = p.X[l - 1] + tile * TILE_FLOATS;
                float* __restrict__ zt = p.DZ[l - 1] + tile * TILE_FLOATS;
                epi_stream<2, 8>(&bar, acc_par, tmem, row, cg, live, true, ht, xt, [&](int col0, float* v, float4 (*aux)[2]) {
#pragma unroll
                    for (int q = 0; q < 2; ++q) {
                        const int j = q * 4;
                        const float4 h = aux[0][q], xq = aux[1][q];
                        float da[4] = {v[j], v[j + 1], v[j + 2], v[j + 3]};
                        if (l == 8) {
                            const float4 w = __ldg(reinterpret_cast<const float4*>(p.w_out0 + col0 + j));
                            da[0] += gs * w.x; da[1] += gs * w.y; da[2] += gs * w.z; da[3] += gs * w.w;
                        }
                        float dz[4] = {sprime_fast(h.x) * da[0] + xq.x, sprime_fast(h.y) * da[1] + xq.y,
                                       sprime_fast(h.z) * da[2] + xq.z, sprime_fast(h.w) * da[3] + xq.w};
                        float g4[4] = {dz[0], dz[1], dz[2], dz[3]};
                        if (l == 4 && col0 + j + 3 > 192) {
                            // cotangent of the encoding part of the skip input: kept in the unused columns of the DZ_3 tile
#pragma unroll
                            for (int i = 0; i < 4; ++i) {
                                if (col0 + j + i > 192) {
                                    g4[i] = da[i];
                                    dz[i] = 0.0f;
                                }
                            }
                        }
                        v[j] = dz[0]; v[j + 1] = dz[1]; v[j + 2] = dz[2]; v[j + 3] = dz[3];
                        if (live) st4(zt + toff(row, col0 + j), make_float4(g4[0], g4[1], g4[2], g4[3]));
                    }
                    a_store8(smem, row, col0, v);
                });
                epi_publish_a(&bar);
            }
            // ---- encoding layer: de = dz_0 W_0 + (skip part); d_x = J_e^T de + Hessian term ---------------------
            epi_wait_acc(&bar, acc_par);
            if (p.d_pts) {
                float v[16];
                tc::tmem_ld_32x32b_x16(tmem + ((uint32_t)(row & ~31) << 16) + (uint32_t)(cg * 16), v);
                tc::tmem_ld_wait();
                if (live) {
                    const float* __restrict__ z3 = p.DZ[3] + tile * TILE_FLOATS;     // skip part: columns 193 + j
#pragma unroll
                    for (int j = 0; j < 16; j += 4) {
                        const int col = cg * 16 + j;
                        float4 e;
                        e.x = v[j] + z3[toff(row, 193 + col)];
                        e.y = v[j + 1] + z3[toff(row, 194 + col)];
                        e.z = v[j + 2] + z3[toff(row, 195 + col)];
                        e.w = col + 3 == 63 ? 0.0f : v[j + 3] + z3[toff(row, 196 + col)];
                        float* __restrict__ de = p.DE + eoff(gp, col);
                        de[0] = e.x; de[TILE_M] = e.y; de[2 * TILE_M] = e.z; de[3 * TILE_M] = e.w;
                    }
                }
                tc::tc_fence_before_sync();
                tc::named_bar_sync(1, EPI_THREADS);
                if (cg < 3 && live) {
                    const float* __restrict__ e = p.E + eoff(gp, 3 + cg * 20);
                    const float* __restrict__ g = p.DE + eoff(gp, 3 + cg * 20);
                    const float* __restrict__ b = p.EB + eoff(gp, 3 + cg * 20);
                    float acc = p.DE[eoff(gp, cg)], hess = 0.0f, f = 1.0f;
#pragma unroll
                    for (int k = 0; k < 10; ++k) {
                        const float sn = e[k * TILE_M], cs = e[(10 + k) * TILE_M];
                        acc += f * (cs * g[k * TILE_M] - sn * g[(10 + k) * TILE_M]);
                        hess -= f * f * (sn * b[k * TILE_M] + cs * b[(10 + k) * TILE_M]);
                        f *= 2.0f;
                    }
                    p.d_pts[gp * 3 + cg] = acc + dn[cg] * hess;
                }
            }
        }
    }
    chain_teardown(&bar);
}

static int check_chain_mlp(const hn_mlp_t* m) {
    HN_REQUIRE(m && m->n_layers == 9, "object SDF mlp must have 9 layers");
    HN_REQUIRE(m->chain && m->chain_bytes >= (int64_t)obj_layout().total && aligned16(m->chain),
               "HN_TC_BF16X3 needs the packed chain operands (hn_sdf_obj_chain_pack)");
    return HN_OK;
}

static long long* g_prof = nullptr;   // set by hn_chain_set_prof (diagnostics)
static int g_stagger_fwd = 0, g_stagger_bwd = 0;   // cycles per stagger slot (hn_chain_set_stagger); measured: no effect

int launch_sdf_only_ts(const hn_mlp_t* m, const float* pts, int64_t n, float inv_scale, float* sdf, cudaStream_t s,
                       const float* xs = nullptr, const float* ys = nullptr, const float* zs = nullptr, int ny = 0, int nz = 0);   // chain_ts.cu
void set_prof_ts(long long* p);

int launch_sdf_only(const hn_mlp_t* m, const float* pts, int64_t n, float inv_scale, float* sdf, cudaStream_t s) {
    HN_PROPAGATE(check_chain_mlp(m));
    static const bool use_ts = getenv("HONERF_SDF_TS") ? atoi(getenv("HONERF_SDF_TS")) != 0 : true;   // default: activations in TMEM
    if (use_ts) return launch_sdf_only_ts(m, pts, n, inv_scale, sdf, s);
    const ObjLayout L = obj_layout();
    SdfOnlyParams p;
    p.pts = pts; p.n = n; p.inv_scale = inv_scale; p.sdf = sdf;
    p.chain = reinterpret_cast<const uint8_t*>(m->chain);
    for (int l = 0; l < 9; ++l) p.bias[l] = m->b[l];
    p.w_out0 = m->W[8];
    p.n_tiles = (int)ceil_div(n, TILE_M);
    p.prof = g_prof;
    Program prog = {};
    prog.n_steps = 8;
    for (int l = 0; l < 8; ++l) {
        prog.step[l].b_off = L.nt16_off[l];
        prog.step[l].n_mma = L.nt_n[l];
        prog.step[l].kblocks = L.nt_kb[l];
        prog.step[l].a_kb0 = 0;
        prog.step[l].f16 = 1;
    }
    static bool configured = false;
    if (!configured) {
        HN_CHECK_CUDA(cudaFuncSetAttribute(sdf_only_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
        configured = true;
    }
    const int grid = std::min(p.n_tiles, sm_count());
    {
        TimingScope ts(s, TT_SDF_ONLY);
        sdf_only_kernel<<<grid, THREADS, SMEM_BYTES, s>>>(p, prog);
    }
    count_launch();
    HN_CHECK_LAUNCH();
    return HN_OK;
}

int launch_sdf_fwd(const hn_mlp_t* m, const float* pts, int64_t n, float inv_scale, float* sdf, float* feat,
                   int64_t ld_feat, float* normal, float* stash, cudaStream_t s) {
    HN_PROPAGATE(check_chain_mlp(m));
    const ObjLayout L = obj_layout();
    FwdParams p;
    p.pts = pts; p.n = n; p.inv_scale = inv_scale; p.sdf = sdf; p.feat = feat; p.ld_feat = ld_feat; p.normal = normal;
    const int64_t np = round_up(n, TILE_M);      // the tiled arrays hold whole tiles
    p.E = stash;
    for (int l = 0; l < 8; ++l) {
        p.H[l] = stash + np * 64 + (int64_t)l * np * 256;
        p.D[l] = stash + np * 64 + (int64_t)(8 + l) * np * 256;
    }
    p.EB = stash + np * 64 + 16 * np * 256;
    p.chain = reinterpret_cast<const uint8_t*>(m->chain);
    for (int l = 0; l < 9; ++l) p.bias[l] = m->b[l];
    p.w_out0 = m->W[8];
    p.n_tiles = (int)ceil_div(n, TILE_M);
    p.prof = g_prof;
    p.stagger = g_stagger_fwd;
    Program prog = {};
    int k = 0;
    for (int l = 0; l < 9; ++l, ++k) {           // value trunk + feature head: a @ W_l^T, fp16 pairs
        prog.step[k].f16 = 1;
        prog.step[k].b_off = L.nt16_off[l];
        prog.step[k].n_mma = L.nt_n[l];
        prog.step[k].kblocks = L.nt_kb[l];
        prog.step[k].a_kb0 = 0;
    }
    for (int l = 7; l >= 0; --l, ++k) {          // normal sweep: d @ W_l
        prog.step[k].b_off = L.nn_off[l];
        prog.step[k].n_mma = L.nn_n[l];
        prog.step[k].kblocks = L.nn_kb[l];
        prog.step[k].a_kb0 = 0;
    }
    prog.n_steps = k;
    static bool configured = false;
    if (!configured) {
        HN_CHECK_CUDA(cudaFuncSetAttribute(sdf_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
        configured = true;
    }
    const int grid = std::min(p.n_tiles, sm_count());
    {
        TimingScope ts(s, TT_SDF_FWD);
        sdf_fwd_kernel<<<grid, THREADS, SMEM_BYTES, s>>>(p, prog);
    }
    count_launch();
    HN_CHECK_LAUNCH();
    return HN_OK;
}

// backward workspace (n padded to whole tiles): UE | U[8] | X[8] | DZ[8] | DE | partial sums of the dW kernel
int64_t bwd_ws_floats(int64_t n) { return round_up(n, TILE_M) * (64 + 3 * 8 * 256 + 64) + dw_part_floats(9); }

int launch_sdf_bwd(const hn_mlp_t* m, int64_t n, float inv_scale, const float* stash, const float* d_sdf,
                   const float* d_feat, int64_t ld_dfeat, const float* d_normal, float* d_pts, float* ws,
                   cudaStream_t s) {
    HN_PROPAGATE(check_chain_mlp(m));
    const ObjLayout L = obj_layout();
    BwdParams p;
    p.n = n; p.inv_scale = inv_scale;
    const int64_t np = round_up(n, TILE_M);
    p.E = stash;
    for (int l = 0; l < 8; ++l) {
        p.H[l] = stash + np * 64 + (int64_t)l * np * 256;
        p.D[l] = stash + np * 64 + (int64_t)(8 + l) * np * 256;
    }
    p.EB = stash + np * 64 + 16 * np * 256;
    p.d_sdf = d_sdf; p.d_feat = d_feat; p.ld_dfeat = ld_dfeat; p.d_normal = d_normal; p.d_pts = d_pts;
    float* q = ws;
    p.UE = q; q += np * 64;
    for (int l = 0; l < 8; ++l) { p.U[l] = q; q += np * 256; }
    for (int l = 0; l < 8; ++l) { p.X[l] = q; q += np * 256; }
    for (int l = 0; l < 8; ++l) { p.DZ[l] = q; q += np * 256; }
    p.DE = q;
    p.chain = reinterpret_cast<const uint8_t*>(m->chain);
    p.w_out0 = m->W[8];
    p.n_tiles = (int)ceil_div(n, TILE_M);
    p.prof = g_prof;
    p.stagger = g_stagger_bwd;
    Program prog = {};
    int k = 0;
    for (int l = 0; l < 8; ++l, ++k) {           // tangent sweep: u @ W_l^T
        prog.step[k].b_off = L.nt_off[l];
        prog.step[k].n_mma = L.nt_n[l];
        prog.step[k].kblocks = L.nt_kb[l];
        prog.step[k].a_kb0 = 0;
    }
    for (int l = 8; l >= 0; --l, ++k) {          // reverse sweep: dz @ W_l
        prog.step[k].b_off = L.nn_off[l];
        prog.step[k].n_mma = L.nn_n[l];
        prog.step[k].kblocks = L.nn_kb[l];
        prog.step[k].a_kb0 = 0;
    }
    prog.n_steps = k;
    static bool configured = false;
    if (!configured) {
        HN_CHECK_CUDA(cudaFuncSetAttribute(sdf_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
        configured = true;
    }
    const int grid = std::min(p.n_tiles, sm_count());
    {
        TimingScope ts(s, TT_SDF_BWD);
        sdf_bwd_kernel<<<grid, THREADS, SMEM_BYTES, s>>>(p, prog);
    }
    count_launch();
    HN_CHECK_LAUNCH();
    return HN_OK;
}

// dW_out[0, :] += inv_scale * sum_p (d_sdf[p] * h7[p, :] + u7[p, :]);  db_out[0] += inv_scale * sum_p d_sdf[p]
// (row 0 of the output layer: the sdf value uses it directly, the normal sweep is seeded with it).
__global__ void __launch_bounds__(128) out_row0_grad_kernel(const float* __restrict__ H7, const float* __restrict__ U7,
                                                            const float* __restrict__ d_sdf, int64_t n, int n_tiles,
                                                            float inv_scale, float* __restrict__ dW_row0,
                                                            float* __restrict__ db0) {
    __shared__ float red[4][5];
    const int c4 = blockIdx.x, row = threadIdx.x;
    const int t0 = (int)((int64_t)n_tiles * blockIdx.y / gridDim.y), t1 = (int)((int64_t)n_tiles * (blockIdx.y + 1) / gridDim.y);
    float a[5] = {0.f, 0.f, 0.f, 0.f, 0.f};
    for (int t = t0; t < t1; ++t) {
        const int64_t pnt = (int64_t)t * TILE_M + row;
        if (pnt >= n) continue;
        const float w = d_sdf ? d_sdf[pnt] : 0.0f;
        const float4 h = ld4(H7 + t * TILE_FLOATS + c4 * 512 + row * 4);
        const float4 u = ld4(U7 + t * TILE_FLOATS + c4 * 512 + row * 4);
        a[0] += w * h.x + u.x; a[1] += w * h.y + u.y; a[2] += w * h.z + u.z; a[3] += w * h.w + u.w;
        a[4] += w;
    }
#pragma unroll
    for (int i = 0; i < 5; ++i) {
        a[i] = warp_sum(a[i]);
        if ((row & 31) == 0) red[row >> 5][i] = a[i];
    }
    __syncthreads();
    if (row < 5) {
        const float v = (red[0][row] + red[1][row] + red[2][row] + red[3][row]) * inv_scale;
        if (row < 4) {
            if (dW_row0) atomicAdd(dW_row0 + c4 * 4 + row, v);
        } else if (c4 == 0 && db0) {
            atomicAdd(db0, v);
        }
    }
}

// all weight / bias gradients of the object SDF net from the operands the backward chain kernel left in HBM
int launch_sdf_bwd_weights(const hn_mlp_t* m, int64_t n, float inv_scale, const float* stash, const float* d_sdf,
                           const float* d_feat, int64_t ld_dfeat, float* ws, const hn_mlp_grad_t* grad,
                           cudaStream_t s) {
    const int64_t np = round_up(n, TILE_M);
    const float* E = stash;
    const float* H = stash + np * 64;
    const float* D = H + 8 * np * 256;
    const float* UE = ws;
    const float* U = UE + np * 64;
    const float* DZ = U + 16 * np * 256;
    float* part = ws + np * (64 + 3 * 8 * 256 + 64);
    DwParams p;
    DwReduceParams r;
    p.n = n; p.n_tiles = (int)(np / TILE_M); p.part = part;
    r.part = part;
    int k = 0;
    for (int l = 0; l < 8; ++l, ++k) {
        DwJob& j = p.job[k];
        const int out = m->out_dim[l], in = m->in_dim[l];
        j.P[0] = {DZ + (int64_t)l * np * 256, 0, out, 1};
        j.P[1] = {D + (int64_t)l * np * 256, 0, out, 1};
        if (l == 0) {
            j.Q[0] = {E, 64, in, 2};          // column-major [64][128] tiles
            j.Q[1] = {UE, 64, in, 2};
        } else {
            j.Q[0] = {H + (int64_t)(l - 1) * np * 256, 0, in, 1};       // H[3] holds the skip input [h3 | e]
            j.Q[1] = {U + (int64_t)(l - 1) * np * 256, 0, in, 1};       // U[3] holds [u3 | ue]
        }
        j.n_pairs = 2;
        j.n_mma = (int)round_up(in, 16);
        j.db = grad->db[l]; j.db_scale = 1.0f;
        r.job[k] = reduce_job(grad->dW[l], m->ld[l], 0, out, in);
    }
    if (d_feat) {
        DwJob& j = p.job[k];
        j.P[0] = {d_feat, ld_dfeat, 256, 0};
        j.Q[0] = {H + (int64_t)7 * np * 256, 0, 256, 1};
        j.P[1] = j.P[0]; j.Q[1] = j.Q[0];
        j.n_pairs = 1;
        j.n_mma = 256;
        j.db = grad->db[8] ? grad->db[8] + 1 : nullptr; j.db_scale = 1.0f;
        r.job[k] = reduce_job(grad->dW[8], m->ld[8], 1, 256, 256);
        ++k;
    }
    p.n_jobs = k;
    HN_PROPAGATE(launch_dw(p, r, s));
    if (grad->dW[8] || grad->db[8]) {
        const int splits = std::max(1, std::min(p.n_tiles, 16));
        out_row0_grad_kernel<<<dim3(64, splits), 128, 0, s>>>(H + (int64_t)7 * np * 256, U + (int64_t)7 * np * 256, d_sdf, n,
                                                            p.n_tiles, inv_scale, grad->dW[8], grad->db[8]);
        count_launch();
        HN_CHECK_LAUNCH();
    }
    return HN_OK;
}

int64_t stash_floats(int64_t n) { return round_up(n, TILE_M) * (64 + 16 * 256 + 64); }

}  // namespace chain
}  // namespace hn

using namespace hn;

extern "C" {

int hn_chain_set_stagger(int fwd_cycles, int bwd_cycles) {
    chain::g_stagger_fwd = fwd_cycles;
    chain::g_stagger_bwd = bwd_cycles;
    return HN_OK;
}

int hn_chain_set_prof(void* buf) {
    chain::g_prof = reinterpret_cast<long long*>(buf);
    chain::set_prof_ts(chain::g_prof);
    return HN_OK;
}

int64_t hn_sdf_obj_chain_bytes(void) { return (int64_t)chain::obj_layout().total; }

int hn_sdf_obj_grid(const hn_mlp_t* mlp, const float* xs, int nx, const float* ys, int ny, const float* zs, int nz,
                    float inv_scale, float* u, hn_stream_t stream) {
    HN_REQUIRE(mlp && xs && ys && zs && u && nx >= 1 && ny >= 1 && nz >= 1, "hn_sdf_obj_grid: bad arguments");
    HN_PROPAGATE(chain::check_chain_mlp(mlp));
    const int64_t n = (int64_t)nx * ny * nz;
    HN_REQUIRE(n < ((int64_t)1 << 31) * 64, "hn_sdf_obj_grid: lattice too large");
    return chain::launch_sdf_only_ts(mlp, nullptr, n, inv_scale, u, (cudaStream_t)stream, xs, ys, zs, ny, nz);
}

int hn_sdf_obj_chain_pack(const hn_mlp_t* m, void* chain_buf, int64_t chain_bytes, hn_stream_t stream) {
    HN_REQUIRE(m && m->n_layers == 9, "hn_sdf_obj_chain_pack: object SDF mlp must have 9 layers");
    const chain::ObjLayout L = chain::obj_layout();
    HN_REQUIRE(chain_buf && chain_bytes >= (int64_t)L.total && aligned16(chain_buf),
               "hn_sdf_obj_chain_pack: buffer too small or misaligned (need %u bytes)", L.total);
    static const int in_d[9] = {63, 256, 256, 256, 256, 256, 256, 256, 256};
    static const int out_d[9] = {256, 256, 256, 193, 256, 256, 256, 256, 257};
    cudaStream_t s = (cudaStream_t)stream;
    uint8_t* dst = reinterpret_cast<uint8_t*>(chain_buf);
    chain::pack_batch_begin();
    for (int l = 0; l < 9; ++l) {
        HN_REQUIRE(m->in_dim[l] == in_d[l] && m->out_dim[l] == out_d[l] && m->W[l] && m->WT[l],
                   "hn_sdf_obj_chain_pack: layer %d has the wrong shape or no transposed copy", l);
        const int row0 = l == 8 ? 1 : 0;
        const int rows = l == 8 ? 256 : out_d[l];
        // NT: B(n = out, k = in) from W [out, ld]
        HN_PROPAGATE(chain::launch_pack_b(m->W[l], m->ld[l], row0, 0, rows, in_d[l], L.nt_n[l], L.nt_kb[l],
                                          dst + L.nt_off[l], s));
        // NN: B(n = in, k = out) from WT [in, ldT]
        HN_PROPAGATE(chain::launch_pack_b(m->WT[l], m->ldT[l], 0, row0, in_d[l], rows, L.nn_n[l], L.nn_kb[l],
                                          dst + L.nn_off[l], s));
        HN_PROPAGATE(chain::launch_pack_b(m->W[l], m->ld[l], chain::pack_map(row0, 0), rows, in_d[l], L.nt_n[l], L.nt_kb[l],
                                          dst + L.nt16_off[l], s, true));
        if (l < 8)
            for (int h = 0; h < 2; ++h)
                HN_PROPAGATE(chain::launch_pack_b(m->W[l], m->ld[l], chain::pack_map(128 * h, 0),
                                                  std::max(0, std::min(128, rows - 128 * h)), in_d[l], 128, L.nt_kb[l],
                                                  dst + L.nth_off[l][h], s, true));
    }
    // HN_TC_MIXED16 operands (chain16_obj.cu): the feature head as 128-row halves, the normal sweep's weights as fp16 pairs
    for (int h = 0; h < 2; ++h)
        HN_PROPAGATE(chain::launch_pack_b(m->W[8], m->ld[8], chain::pack_map(1 + 128 * h, 0), 128, 256, 128, 4,
                                          dst + L.nth8_off[h], s, true));
    for (int l = 0; l < 8; ++l)
        HN_PROPAGATE(chain::launch_pack_b(m->WT[l], m->ldT[l], chain::pack_map(0, 0), in_d[l], out_d[l], L.nn_n[l], L.nn_kb[l],
                                          dst + L.nn16_off[l], s, true));
    return chain::pack_batch_flush(s);
}

}  // extern "C"
