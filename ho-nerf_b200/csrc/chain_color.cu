// Object colour field (RenderingNetwork_OBJ.forward, utils/fields.py:387-405) as fused tile-chain kernels
// on tcgen05 (HN_TC_BF16X3), forward and backward.  Input row of the first layer:
//   [pts + enc10 (63) | dirs + enc4 (27) | feature (256) | normal + enc4 (27)] = 373
// The 373-wide first layer is two chained steps on one accumulator: the 256 feature columns, then the 117
// encoding columns (computed by the epilogue warps straight into the A operand).  ReLU between layers,
// sigmoid on the 3 outputs (the last layer runs as an N = 16 MMA).
//
// S16 (HN_TC_MIXED16): same arithmetic in the chain, but everything the backward and the weight gradients read back is a
// pair of 16-bit "dW-ready" T16 tiles (chain16.cuh), the hi and lo halves the chain's A operand holds anyway: R16[l] = relu
// output (the sign of its hi half is the backward's mask), FEAT16, ENC16 (padded to 256 columns) and DZ16[l] -- the weight
// gradients then run on dw16_kernel with bulk-copied operands and no conversion, three bf16 MMAs per product (Ph Qh + Pl Qh +
// Ph Ql).  One MMA per product on the hi halves alone was measured and rejected: the colour net's weight gradients are
// cancelling sums, a single bf16 rounding of the operands moved them from 8e-3 to 1.2e-2 relative (the SDF net's stay <= 4e-3).
#include <algorithm>

#include "chain16.cuh"
#include "chain_dw.cuh"
#include "fields_common.cuh"

namespace hn {
namespace chain {

// entry of point `point`, column `col` in the column-major [tile][128 columns][128 rows] encoding stash
__host__ __device__ __forceinline__ int64_t coff(int64_t point, int col = 0) {
    return ((point >> 7) << 14) + (int64_t)col * TILE_M + (point & 127);
}
constexpr int ENC_LD = 128;      // [enc10(pts) 63 | enc4(dirs) 27 | enc4(normal) 27 | 0 x 11]
constexpr int ENC_DIRS = 63, ENC_NRM = 90, ENC_DIM = 117;
constexpr int COLOR16_STASH_FLOATS = ENC_LD + 12 * 128;     // S16 stash per point: ENC fp32 + hi / lo T16 tiles of six arrays
constexpr int CIN_FEAT0 = 90, CIN_NRM0 = 346;     // column offsets inside the reference's 373-wide input

struct ColorLayout {
    uint32_t nt0a, nt0b, nt[5], nn[5], nn0a, nn0b, total;   // nt[1..4], nn[1..4] used
};
static ColorLayout color_layout() {
    ColorLayout L;
    uint32_t off = 0;
    L.nt0a = off; off += b_operand_bytes(256, 4);
    L.nt0b = off; off += b_operand_bytes(256, 2);
    for (int l = 1; l <= 3; ++l) { L.nt[l] = off; off += b_operand_bytes(256, 4); }
    L.nt[4] = off; off += b_operand_bytes(16, 4);
    L.nn[4] = off; off += b_operand_bytes(256, 1);
    for (int l = 3; l >= 1; --l) { L.nn[l] = off; off += b_operand_bytes(256, 4); }
    L.nn0a = off; off += b_operand_bytes(256, 4);
    L.nn0b = off; off += b_operand_bytes(128, 4);
    L.nt[0] = L.nn[0] = 0;
    L.total = off;
    return L;
}

// [x(3), sin/cos(2^k x_c), k < L] of a 3-vector -> columns col_base + j of the A operand and of the point's entry g
// (may be NULL) of a column-major [128][128] stash tile (column j at g[128 j], see coff); the (coordinate, frequency)
// pairs are shared by the 4 column groups of a row
__device__ __forceinline__ void write_enc3(uint8_t* smem, int row, int cg, const float x[3], int L, int col_base,
                                           float* __restrict__ g) {
    if (cg == 0) {
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            a_store1(smem, row, col_base + c, x[c]);
            if (g) g[(col_base + c) * TILE_M] = x[c];
        }
    }
    for (int idx = cg; idx < 3 * L; idx += EPI_CGROUPS) {
        const int c = idx / L, k = idx - c * L;
        float s, co;
        sincosf(x[c] * (float)(1 << k), &s, &co);
        const int js = col_base + 3 + c * 2 * L + k, jc = js + L;
        a_store1(smem, row, js, s);
        a_store1(smem, row, jc, co);
        if (g) { g[js * TILE_M] = s; g[jc * TILE_M] = co; }
    }
}

struct ColorFwdParams {
    const float* pts;
    const float* dirs;
    const float* feat;
    int64_t ld_feat;
    const float* feat2;     // TAIL: second addend of the first layer's pre-activation rows (may be NULL), same ld
    const float* normal;
    int64_t n;
    float* rgb;
    float* ENC;       // stash: column-major tiles [tile][128 columns][128 rows]
    float* FEAT;      // stash: tiled copy of the feature input (operand of the first layer's weight gradient)
    float* R[4];      // stash: tiled ReLU outputs
    // S16: bf16 T16 tiles instead of FEAT / R (ENC stays fp32 for the backward's J^T; ENC16 is the weight-gradient operand)
    uint8_t* FEAT16;  // hi tiles; the lo tile of each array follows at + lo_off bytes
    uint8_t* R16[4];
    uint8_t* ENC16;
    size_t lo_off;
    int store_lo;
    const uint8_t* chain;
    const float* bias[5];
    int n_tiles;
};

// TAIL (hand colour net, forward-only rendering): the first layer's ReLU output arrives as fp32 rows in p.feat (its 1669-wide
// contraction stays a per-layer kernel); the chain runs layers 1..3 and the sigmoid output layer, nothing is stashed.
template <bool S16, bool TAIL = false>
__global__ void __launch_bounds__(THREADS, 1)
color_fwd_kernel(const __grid_constant__ ColorFwdParams p, const __grid_constant__ Program prog) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    __shared__ Barriers bar;
    uint8_t* smem = chain_setup(smem_raw, &bar);
    const int warp = threadIdx.x >> 5;
    const int n_my_tiles = p.n_tiles > (int)blockIdx.x ? (p.n_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x : 0;
    if (warp == 0) {
        producer_loop(prog, p.chain, smem, &bar, n_my_tiles);
    } else if (warp == 1) {
        mma_loop(prog, smem, &bar, n_my_tiles);
    } else {
        int row, cg;
        epi_coords(row, cg);
        const uint32_t tmem = bar.tmem_base;
        uint32_t acc_par = 0;
        for (int t = 0; t < n_my_tiles; ++t) {
            const int64_t tile = (int64_t)blockIdx.x + (int64_t)t * gridDim.x;
            const int64_t gp = tile * TILE_M + row;
            const bool live = gp < p.n;
            // ---- A <- feature columns (and their tiled copy for the backward).  A plain copy of row-major rows, so the
            // threads are mapped for coalescing: consecutive threads take consecutive float4 of a row.
            {
                float* __restrict__ ft = p.FEAT + tile * TILE_FLOATS;
                // a warp takes 8 rows x 4 float4: the row-major loads are 64-byte segments, the tiled stores 128-byte ones
                const int ew = (threadIdx.x - 64) >> 5, ln = threadIdx.x & 31;
#pragma unroll 4
                for (int it = 0; it < 16; ++it) {
                    const int b = it * EPI_WARPS + ew;           // 256 blocks of 8 rows x 16 columns
                    const int r = (b >> 4) * 8 + (ln & 7), c = ((b & 15) * 4 + (ln >> 3)) * 4;
                    const int64_t g = tile * TILE_M + r;
                    float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (g < p.n) {
                        a = ld4(p.feat + g * p.ld_feat + c);
                        if (!S16 && !TAIL) st4(ft + toff(r, c), a);
                        if (TAIL) {      // rows are PRE-activations of the first layer (two partial contractions): + bias, ReLU
                            if (p.feat2) {
                                const float4 b2 = ld4(p.feat2 + g * p.ld_feat + c);
                                a.x += b2.x; a.y += b2.y; a.z += b2.z; a.w += b2.w;
                            }
                            const float4 b0 = __ldg(reinterpret_cast<const float4*>(p.bias[0] + c));
                            a.x = fmaxf(a.x + b0.x, 0.0f); a.y = fmaxf(a.y + b0.y, 0.0f);
                            a.z = fmaxf(a.z + b0.z, 0.0f); a.w = fmaxf(a.w + b0.w, 0.0f);
                        }
                    }
                    uint2 hi, lo;
                    split2(a.x, a.y, hi.x, lo.x);
                    split2(a.z, a.w, hi.y, lo.y);
                    if (S16 && !TAIL) {    // rows past n are stored too (zeros): the weight-gradient MMAs read whole tiles
                        uint8_t* f16 = p.FEAT16 + (size_t)tile * T16_TILE_BYTES + t16_off(r, c >> 3) + (uint32_t)(c & 4) * 2u;
                        *reinterpret_cast<uint2*>(f16) = hi;
                        if (p.store_lo) *reinterpret_cast<uint2*>(f16 + p.lo_off) = lo;
                    }
                    const uint32_t off = (uint32_t)(c >> 6) * KB_BYTES + tc::sw128_offset((uint32_t)r, (uint32_t)((c & 63) >> 3)) +
                                         (uint32_t)(c & 4) * 2u;
                    *reinterpret_cast<uint2*>(smem + off) = hi;
                    *reinterpret_cast<uint2*>(smem + A_LO_OFF + off) = lo;
                }
            }
            epi_publish_a(&bar);
            // ---- A <- encodings of pts, dirs, normal (columns 0..127), accumulated onto the feature part ------
            if (!TAIL) {
            epi_wait_acc(&bar, acc_par);
            {
                float x[3] = {0.f, 0.f, 0.f}, d[3] = {0.f, 0.f, 0.f}, nr[3] = {0.f, 0.f, 0.f};
                if (live) {
#pragma unroll
                    for (int c = 0; c < 3; ++c) { x[c] = p.pts[gp * 3 + c]; d[c] = p.dirs[gp * 3 + c]; nr[c] = p.normal[gp * 3 + c]; }
                }
                float* __restrict__ g = live ? p.ENC + coff(gp) : nullptr;
                write_enc3(smem, row, cg, x, 10, 0, g);
                write_enc3(smem, row, cg, d, 4, ENC_DIRS, g);
                write_enc3(smem, row, cg, nr, 4, ENC_NRM, g);
                if (cg == 0) {
                    for (int j = ENC_DIM; j < ENC_LD; ++j) {
                        a_store1(smem, row, j, 0.0f);
                        if (g) g[j * TILE_M] = 0.0f;
                    }
                }
                if (S16) {
                    // ENC16: the row's 128 encoding columns as bf16 chunks, read back from the A operand the four column
                    // groups of the row just wrote (hi half = the bf16 rounding of the value)
                    tc::named_bar_sync(1, EPI_THREADS);
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        const int c = (cg * 4 + q) * 8;
                        const uint32_t off = (uint32_t)(c >> 6) * KB_BYTES + tc::sw128_offset((uint32_t)row, (uint32_t)((c & 63) >> 3));
                        uint8_t* e16 = p.ENC16 + (size_t)tile * T16_TILE_BYTES + t16_off(row, c >> 3);
                        stg16(e16, *reinterpret_cast<const uint4*>(smem + off));
                        if (p.store_lo) stg16(e16 + p.lo_off, *reinterpret_cast<const uint4*>(smem + A_LO_OFF + off));
                    }
                }
            }
            epi_publish_a(&bar);
            }
            // ---- hidden layers: ReLU ----------------------------------------------------------------------------
            for (int l = TAIL ? 1 : 0; l < 4; ++l) {
                epi_wait_acc(&bar, acc_par);
                const float* __restrict__ bias = p.bias[l];
                float* __restrict__ rt = p.R[l] + tile * TILE_FLOATS;
#pragma unroll
                for (int blk = 0; blk < EPI_COLS / 32; ++blk) {
                    const int col0 = cg * EPI_COLS + blk * 32;
                    float v[32];
                    acc_load32(tmem, row, col0, v);
#pragma unroll
                    for (int j = 0; j < 32; j += 4) {
                        const float4 b = __ldg(reinterpret_cast<const float4*>(bias + col0 + j));
                        v[j] = fmaxf(v[j] + b.x, 0.0f);
                        v[j + 1] = fmaxf(v[j + 1] + b.y, 0.0f);
                        v[j + 2] = fmaxf(v[j + 2] + b.z, 0.0f);
                        v[j + 3] = fmaxf(v[j + 3] + b.w, 0.0f);
                        if (!S16 && !TAIL && live) st4(rt + toff(row, col0 + j), make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]));
                    }
#pragma unroll
                    for (int j = 0; j < 32; j += 8) {
                        a_store8(smem, row, col0 + j, v + j);
                        if (S16 && !TAIL) {
                            uint4 qh, ql;
                            split2(v[j], v[j + 1], qh.x, ql.x); split2(v[j + 2], v[j + 3], qh.y, ql.y);
                            split2(v[j + 4], v[j + 5], qh.z, ql.z); split2(v[j + 6], v[j + 7], qh.w, ql.w);
                            uint8_t* r16 = p.R16[l] + (size_t)tile * T16_TILE_BYTES + t16_off(row, (col0 + j) >> 3);
                            stg16(r16, qh);
                            if (p.store_lo) stg16(r16 + p.lo_off, ql);
                        }
                    }
                }
                epi_publish_a(&bar);
            }
            // ---- output layer: sigmoid ---------------------------------------------------------------------------
            epi_wait_acc(&bar, acc_par);
            if (cg == 0) {
                float v[16];
                tc::tmem_ld_32x32b_x16(tmem + ((uint32_t)(row & ~31) << 16), v);
                tc::tmem_ld_wait();
                if (live) {
#pragma unroll
                    for (int c = 0; c < 3; ++c) p.rgb[gp * 3 + c] = 1.0f / (1.0f + __expf(-(v[c] + __ldg(p.bias[4] + c))));
                }
            }
        }
    }
    chain_teardown(&bar);
}

struct ColorBwdParams {
    int64_t n;
    const float* ENC;
    const float* R[4];
    const float* rgb;
    const float* d_rgb;
    float* d_pts;      // any of the four input cotangents may be NULL
    float* d_dirs;
    float* d_feat;
    int64_t ld_dfeat;
    float* d_normal;
    float* DZ4;        // workspace: [np, 4] row-major
    float* DZ[4];      // tiled
    float* DENC;       // column-major tiles like ENC
    const uint8_t* R16[4];   // S16: bf16 T16 tiles (mask = sign of the stored relu output), DZ16 hi / lo written for dw16_kernel
    uint8_t* DZ16[4];
    size_t dz_lo_off;
    int store_lo;
    const uint8_t* chain;
    int n_tiles;
};

template <bool S16>
__global__ void __launch_bounds__(THREADS, 1)
color_bwd_kernel(const __grid_constant__ ColorBwdParams p, const __grid_constant__ Program prog) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    __shared__ Barriers bar;
    uint8_t* smem = chain_setup(smem_raw, &bar);
    const int warp = threadIdx.x >> 5;
    const int n_my_tiles = p.n_tiles > (int)blockIdx.x ? (p.n_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x : 0;
    if (warp == 0) {
        producer_loop(prog, p.chain, smem, &bar, n_my_tiles);
    } else if (warp == 1) {
        mma_loop(prog, smem, &bar, n_my_tiles);
    } else {
        int row, cg;
        epi_coords(row, cg);
        const uint32_t tmem = bar.tmem_base;
        uint32_t acc_par = 0;
        for (int t = 0; t < n_my_tiles; ++t) {
            const int64_t tile = (int64_t)blockIdx.x + (int64_t)t * gridDim.x;
            const int64_t gp = tile * TILE_M + row;
            const bool live = gp < p.n;
            // ---- A <- dz_4 = d_rgb * rgb (1 - rgb), K padded to 64 ------------------------------------------
            if (cg == 0) {
                float z[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
                if (live) {
#pragma unroll
                    for (int c = 0; c < 3; ++c) {
                        const float y = p.rgb[gp * 3 + c];
                        z[c] = p.d_rgb[gp * 3 + c] * y * (1.0f - y);
                    }
                    st4(p.DZ4 + gp * 4, make_float4(z[0], z[1], z[2], 0.0f));
                }
                a_store8(smem, row, 0, z);
                const float zero[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll
                for (int j = 8; j < 64; j += 8) a_store8(smem, row, j, zero);
            }
            epi_publish_a(&bar);
            // ---- hidden layers: dz_{l-1} = [r_{l-1} > 0] (dz_l W_l), l = 4..1 ----------------------------------------
            for (int l = 4; l >= 1; --l) {
                if (S16) {
                    // the eight 8-column chunks of this thread's R16 row segment, in flight while the MMAs run
                    const size_t tb = (size_t)tile * T16_TILE_BYTES;
                    uint4 rq[EPI_COLS / 8];
#pragma unroll
                    for (int i = 0; i < EPI_COLS / 8; ++i) rq[i] = ldg16(p.R16[l - 1] + tb + t16_off(row, (cg * EPI_COLS >> 3) + i));
                    epi_wait_acc(&bar, acc_par);
#pragma unroll
                    for (int i = 0; i < EPI_COLS / 8; ++i) {
                        const int col0 = cg * EPI_COLS + i * 8;
                        float v[8];
                        tc::tmem_ld_32x32b_x8(tmem + ((uint32_t)(row & ~31) << 16) + (uint32_t)col0, v);
                        tc::tmem_ld_wait();
                        const uint32_t w[4] = {rq[i].x, rq[i].y, rq[i].z, rq[i].w};
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            v[2 * j] = bf16_lo(w[j]) > 0.0f ? v[2 * j] : 0.0f;
                            v[2 * j + 1] = bf16_hi(w[j]) > 0.0f ? v[2 * j + 1] : 0.0f;
                        }
                        uint4 qh, ql;
                        split2(v[0], v[1], qh.x, ql.x); split2(v[2], v[3], qh.y, ql.y); split2(v[4], v[5], qh.z, ql.z); split2(v[6], v[7], qh.w, ql.w);
                        uint8_t* z16 = p.DZ16[l - 1] + tb + t16_off(row, col0 >> 3);
                        stg16(z16, qh);
                        if (p.store_lo) stg16(z16 + p.dz_lo_off, ql);
                        a_store8(smem, row, col0, v);
                    }
                    epi_publish_a(&bar);
                    continue;
                }
                const float* __restrict__ rt = p.R[l - 1] + tile * TILE_FLOATS;
                float* __restrict__ zt = p.DZ[l - 1] + tile * TILE_FLOATS;
                epi_stream<1, 8>(&bar, acc_par, tmem, row, cg, live, true, rt, rt, [&](int col0, float* v, float4 (*aux)[2]) {
#pragma unroll
                    for (int q = 0; q < 2; ++q) {
                        const int j = q * 4;
                        const float4 r = aux[0][q];
                        v[j] = r.x > 0.0f ? v[j] : 0.0f;
                        v[j + 1] = r.y > 0.0f ? v[j + 1] : 0.0f;
                        v[j + 2] = r.z > 0.0f ? v[j + 2] : 0.0f;
                        v[j + 3] = r.w > 0.0f ? v[j + 3] : 0.0f;
                        if (live) st4(zt + toff(row, col0 + j), make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]));
                    }
                    a_store8(smem, row, col0, v);
                });
                epi_publish_a(&bar);
            }
            // ---- input cotangent, feature columns: d_feat = dz_0 W_0[:, 90:346] ------------------------------------------
            epi_wait_acc(&bar, acc_par);
            if (p.d_feat) {
#pragma unroll
                for (int blk = 0; blk < EPI_COLS / 32; ++blk) {
                    const int col0 = cg * EPI_COLS + blk * 32;
                    float v[32];
                    acc_load32(tmem, row, col0, v);
                    if (live) {
#pragma unroll
                        for (int j = 0; j < 32; j += 4)
                            st4(p.d_feat + gp * p.ld_dfeat + col0 + j, make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]));
                    }
                }
            }
            epi_publish_a(&bar);      // A (dz_0) is unchanged; this only releases the accumulator
            // ---- input cotangent, encoding columns -> d_pts, d_dirs, d_normal through J_enc^T ----------------------------
            epi_wait_acc(&bar, acc_par);
            if (p.d_pts || p.d_dirs || p.d_normal) {
                float v[32];
                acc_load32(tmem, row, cg * 32, v);
                if (live) {
#pragma unroll
                    for (int j = 0; j < 32; ++j) p.DENC[coff(gp, cg * 32 + j)] = v[j];
                }
                tc::tc_fence_before_sync();
                tc::named_bar_sync(1, EPI_THREADS);
                if (live && cg < 3) {
                    const float* __restrict__ e = p.ENC + coff(gp);
                    const float* __restrict__ g = p.DENC + coff(gp);
                    float* out = cg == 0 ? p.d_pts : cg == 1 ? p.d_dirs : p.d_normal;
                    const int base = cg == 0 ? 0 : cg == 1 ? ENC_DIRS : ENC_NRM;
                    const int L = cg == 0 ? 10 : 4;
                    if (out) {
#pragma unroll
                        for (int c = 0; c < 3; ++c) out[gp * 3 + c] = enc3_jt_from_enc(e + base * TILE_M, g + base * TILE_M, L, c, TILE_M);
                    }
                }
            }
        }
    }
    chain_teardown(&bar);
}

// S16: dW_4[c, :] += sum_p dz4[p, c] r3[p, :], db_4[c] += sum_p dz4[p, c]  (3 output rows: too thin for an MMA tile).
// grid (32 chunks of 8 features, splits), 128 threads = the points of a tile.
__global__ void __launch_bounds__(128) color_out_grad16_kernel(const uint8_t* __restrict__ R3, const uint8_t* __restrict__ R3L,
                                                              const float* __restrict__ DZ4, int64_t n, int n_tiles,
                                                              float* __restrict__ dW4, int ld, float* __restrict__ db4) {
    __shared__ float red[4][27];
    const int f8 = blockIdx.x, p = threadIdx.x;
    const int t0 = (int)((int64_t)n_tiles * blockIdx.y / gridDim.y), t1 = (int)((int64_t)n_tiles * (blockIdx.y + 1) / gridDim.y);
    float a[27];
#pragma unroll
    for (int i = 0; i < 27; ++i) a[i] = 0.0f;
    for (int t = t0; t < t1; ++t) {
        const int64_t pnt = (int64_t)t * TILE_M + p;
        if (pnt >= n) continue;
        const float4 z = ld4(DZ4 + pnt * 4);
        const uint4 h = ldg16(R3 + (size_t)t * T16_TILE_BYTES + t16_off(p, f8));
        float r[8] = {bf16_lo(h.x), bf16_hi(h.x), bf16_lo(h.y), bf16_hi(h.y), bf16_lo(h.z), bf16_hi(h.z), bf16_lo(h.w), bf16_hi(h.w)};
        if (R3L) {        // lo halves of the pair (may be absent)
            const uint4 l = ldg16(R3L + (size_t)t * T16_TILE_BYTES + t16_off(p, f8));
            r[0] += bf16_lo(l.x); r[1] += bf16_hi(l.x); r[2] += bf16_lo(l.y); r[3] += bf16_hi(l.y);
            r[4] += bf16_lo(l.z); r[5] += bf16_hi(l.z); r[6] += bf16_lo(l.w); r[7] += bf16_hi(l.w);
        }
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            a[i] += z.x * r[i];
            a[8 + i] += z.y * r[i];
            a[16 + i] += z.z * r[i];
        }
        a[24] += z.x; a[25] += z.y; a[26] += z.z;
    }
#pragma unroll
    for (int i = 0; i < 27; ++i) {
        a[i] = warp_sum(a[i]);
        if ((p & 31) == 0) red[p >> 5][i] = a[i];
    }
    __syncthreads();
    if (p < 27) {
        const float v = red[0][p] + red[1][p] + red[2][p] + red[3][p];
        if (p < 24) {
            if (dW4) atomicAdd(dW4 + (p >> 3) * ld + f8 * 8 + (p & 7), v);
        } else if (f8 == 0 && db4) {
            atomicAdd(db4 + (p - 24), v);
        }
    }
}

// S16 weight gradients: 1 = three bf16 MMAs per product on hi / lo tile pairs (the arithmetic of the fp32-stash path, agrees with
// it to 4e-7), 0 = one MMA on the hi tiles alone (lo tiles not written: half the stash traffic; weight gradients move by ~2e-3
// relative like the SDF net's).  HONERF_COLOR_DW_X3 overrides the default.
static bool color_dw_x3() {
    static const bool v = getenv("HONERF_COLOR_DW_X3") ? atoi(getenv("HONERF_COLOR_DW_X3")) != 0 : true;
    return v;
}

static int check_color_chain(const hn_mlp_t* m) {
    HN_REQUIRE(m && m->n_layers == 5, "object colour mlp must have 5 layers");
    HN_REQUIRE(m->chain && m->chain_bytes >= (int64_t)color_layout().total && aligned16(m->chain),
               "HN_TC_BF16X3 needs the packed chain operands (hn_color_obj_chain_pack)");
    return HN_OK;
}
static void set_step(Step& s, uint32_t off, int n_mma, int kb, int acc_in = 0) {
    s.b_off = off; s.n_mma = (uint16_t)n_mma; s.kblocks = (uint8_t)kb; s.a_kb0 = 0; s.acc_in = (uint8_t)acc_in;
}

int64_t color_stash_floats(int64_t n) { return round_up(n, TILE_M) * std::max(ENC_LD + 5 * 256, COLOR16_STASH_FLOATS); }
int64_t color_bwd_ws_floats(int64_t n) { return round_up(n, TILE_M) * (4 + 4 * 256 + ENC_LD) + dw_part_floats(6); }

// S16 stash (floats per padded point): ENC fp32 128 | hi tiles: ENC16 (256-column T16 tile) 128 | FEAT16 128 | R16[4] 4 x 128 |
// the lo tiles of the same six arrays, lo_off bytes after their hi tiles
struct Color16Stash {
    float* ENC;
    uint8_t *ENC16, *FEAT16, *R16[4];
    size_t lo_off;
    Color16Stash(float* stash, int64_t np) {
        ENC = stash;
        uint8_t* b = reinterpret_cast<uint8_t*>(stash + np * ENC_LD);
        ENC16 = b; b += np * 512;
        FEAT16 = b; b += np * 512;
        for (int l = 0; l < 4; ++l) { R16[l] = b; b += np * 512; }
        lo_off = (size_t)np * 512 * 6;
    }
};

int launch_color_fwd(const hn_mlp_t* m, const float* pts, const float* dirs, const float* feat, int64_t ld_feat,
                     const float* normal, int64_t n, float* rgb, float* stash, cudaStream_t s, bool s16) {
    HN_PROPAGATE(check_color_chain(m));
    HN_REQUIRE(ld_feat % 4 == 0 && aligned16(feat), "feature input must be 16-byte aligned with ld %% 4 == 0");
    const ColorLayout L = color_layout();
    const int64_t np = round_up(n, TILE_M);
    ColorFwdParams p;
    p.pts = pts; p.dirs = dirs; p.feat = feat; p.ld_feat = ld_feat; p.normal = normal; p.n = n; p.rgb = rgb;
    p.ENC = stash;
    p.FEAT = stash + np * ENC_LD;
    for (int l = 0; l < 4; ++l) p.R[l] = stash + np * ENC_LD + (int64_t)(1 + l) * np * 256;
    const Color16Stash S(stash, np);
    p.FEAT16 = S.FEAT16; p.ENC16 = S.ENC16; p.lo_off = S.lo_off; p.store_lo = color_dw_x3() ? 1 : 0;
    for (int l = 0; l < 4; ++l) p.R16[l] = S.R16[l];
    p.chain = reinterpret_cast<const uint8_t*>(m->chain);
    for (int l = 0; l < 5; ++l) p.bias[l] = m->b[l];
    p.n_tiles = (int)(np / TILE_M);
    Program prog = {};
    set_step(prog.step[0], L.nt0a, 256, 4);
    set_step(prog.step[1], L.nt0b, 256, 2, 1);
    for (int l = 1; l <= 3; ++l) set_step(prog.step[1 + l], L.nt[l], 256, 4);
    set_step(prog.step[5], L.nt[4], 16, 4);
    prog.n_steps = 6;
    static bool configured = false;
    if (!configured) {
        HN_CHECK_CUDA(cudaFuncSetAttribute(color_fwd_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
        HN_CHECK_CUDA(cudaFuncSetAttribute(color_fwd_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
        configured = true;
    }
    {
        TimingScope ts(s, TT_COLOR_FWD);
        if (s16) color_fwd_kernel<true><<<std::min(p.n_tiles, sm_count()), THREADS, SMEM_BYTES, s>>>(p, prog);
        else color_fwd_kernel<false><<<std::min(p.n_tiles, sm_count()), THREADS, SMEM_BYTES, s>>>(p, prog);
    }
    count_launch();
    HN_CHECK_LAUNCH();
    return HN_OK;
}

// Hand colour net, forward-only: layers 1..3 + output of a 5-layer colour mlp (same shapes as the object colour net's) on the
// chain kernel; `ops` = operands packed by color_tail_pack at the ColorLayout offsets; Z0a (+ Z0b) = the first layer's
// pre-activation rows WITHOUT bias (one or two partial contractions), bias and ReLU are applied while they are loaded
int launch_color_tail_fwd(const hn_mlp_t* m, const uint8_t* ops, const float* Z0a, const float* Z0b, int64_t ld_z, int64_t n,
                          float* rgb, cudaStream_t s) {
    HN_REQUIRE(ld_z % 4 == 0 && aligned16(Z0a) && aligned16(Z0b), "colour tail: first-layer rows must be 16-byte aligned with ld %% 4 == 0");
    const ColorLayout L = color_layout();
    ColorFwdParams p = {};
    p.feat = Z0a; p.feat2 = Z0b; p.ld_feat = ld_z; p.n = n; p.rgb = rgb;
    p.chain = ops;
    for (int l = 0; l < 5; ++l) p.bias[l] = m->b[l];
    p.n_tiles = (int)ceil_div(n, TILE_M);
    Program prog = {};
    for (int l = 1; l <= 3; ++l) set_step(prog.step[l - 1], L.nt[l], 256, 4);
    set_step(prog.step[3], L.nt[4], 16, 4);
    prog.n_steps = 4;
    static bool configured = false;
    if (!configured) {
        HN_CHECK_CUDA(cudaFuncSetAttribute(color_fwd_kernel<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
        configured = true;
    }
    color_fwd_kernel<false, true><<<std::min(p.n_tiles, sm_count()), THREADS, SMEM_BYTES, s>>>(p, prog);
    count_launch();
    HN_CHECK_LAUNCH();
    return HN_OK;
}
int64_t color_tail_bytes() { return (int64_t)color_layout().total; }
int color_tail_pack(const hn_mlp_t* m, uint8_t* dst, cudaStream_t s) {
    const ColorLayout L = color_layout();
    pack_batch_begin();
    for (int l = 1; l <= 3; ++l) HN_PROPAGATE(launch_pack_b(m->W[l], m->ld[l], 0, 0, 256, 256, 256, 4, dst + L.nt[l], s));
    HN_PROPAGATE(launch_pack_b(m->W[4], m->ld[4], 0, 0, 3, 256, 16, 4, dst + L.nt[4], s));
    return pack_batch_flush(s);
}

int launch_color_bwd(const hn_mlp_t* m, int64_t n, const float* stash, const float* rgb, const float* d_rgb, float* d_pts,
                     float* d_dirs, float* d_feat, int64_t ld_dfeat, float* d_normal, const hn_mlp_grad_t* grad, float* ws,
                     cudaStream_t s, bool s16) {
    HN_PROPAGATE(check_color_chain(m));
    HN_REQUIRE(!d_feat || (ld_dfeat % 4 == 0 && aligned16(d_feat)), "d_feat must be 16-byte aligned with ld %% 4 == 0");
    const ColorLayout L = color_layout();
    const int64_t np = round_up(n, TILE_M);
    ColorBwdParams p;
    p.n = n;
    p.ENC = stash;
    const float* FEAT = stash + np * ENC_LD;
    for (int l = 0; l < 4; ++l) p.R[l] = stash + np * ENC_LD + (int64_t)(1 + l) * np * 256;
    p.rgb = rgb; p.d_rgb = d_rgb; p.d_pts = d_pts; p.d_dirs = d_dirs; p.d_feat = d_feat; p.ld_dfeat = ld_dfeat; p.d_normal = d_normal;
    p.DZ4 = ws;
    for (int l = 0; l < 4; ++l) p.DZ[l] = ws + np * 4 + (int64_t)l * np * 256;
    p.DENC = ws + np * 4 + 4 * np * 256;
    float* part = p.DENC + np * ENC_LD;
    // S16 views of the same buffers: R16 in the stash, DZ16 hi tiles (4 x 512 B per point) where the fp32 DZ tiles start, their lo
    // tiles in the second half of that region
    const Color16Stash S(const_cast<float*>(stash), np);
    p.dz_lo_off = (size_t)np * 512 * 4;
    p.store_lo = (grad && color_dw_x3()) ? 1 : 0;
    for (int l = 0; l < 4; ++l) {
        p.R16[l] = S.R16[l];
        p.DZ16[l] = reinterpret_cast<uint8_t*>(ws + np * 4) + (size_t)l * np * 512;
    }
    p.chain = reinterpret_cast<const uint8_t*>(m->chain);
    p.n_tiles = (int)(np / TILE_M);
    Program prog = {};
    set_step(prog.step[0], L.nn[4], 256, 1);
    for (int l = 3; l >= 1; --l) set_step(prog.step[4 - l], L.nn[l], 256, 4);
    set_step(prog.step[4], L.nn0a, 256, 4);
    set_step(prog.step[5], L.nn0b, 128, 4);
    prog.n_steps = 6;
    static bool configured = false;
    if (!configured) {
        HN_CHECK_CUDA(cudaFuncSetAttribute(color_bwd_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
        HN_CHECK_CUDA(cudaFuncSetAttribute(color_bwd_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
        configured = true;
    }
    {
        TimingScope ts(s, TT_COLOR_BWD);
        if (s16) color_bwd_kernel<true><<<std::min(p.n_tiles, sm_count()), THREADS, SMEM_BYTES, s>>>(p, prog);
        else color_bwd_kernel<false><<<std::min(p.n_tiles, sm_count()), THREADS, SMEM_BYTES, s>>>(p, prog);
    }
    count_launch();
    HN_CHECK_LAUNCH();
    if (!grad) return HN_OK;
    if (s16) {
        // ---- weight gradients from the 16-bit tiles: layers 0 (feature and encoding columns), 1..3 on dw16_kernel, the
        //      3-row output layer on its own small kernel ----------------------------------------------------------------
        Dw16Params dp = {};
        DwReduceParams rp;
        dp.n_tiles = p.n_tiles; dp.part = part; rp.part = part;
        int k = 0;
        auto job16 = [&](const uint8_t* P, const uint8_t* Q, int n_mma, float* db, DwReduceJob r) {
            Dw16Job& j = dp.job[k];
            j.P[0] = P; j.Q[0] = Q;                               // Ph Qh
            j.P[1] = P + p.dz_lo_off; j.Q[1] = Q;                 // Pl Qh
            j.P[2] = P; j.Q[2] = Q + S.lo_off;                    // Ph Ql
            j.n_pairs = color_dw_x3() ? 3 : 1; j.db_mask = color_dw_x3() ? 3 : 1;      // bias gradient = column sums of Ph (+ Pl)
            j.q_chunks = 32; j.n_mma = n_mma; j.db = db; j.p_cols = 256;
            rp.job[k] = r;
            ++k;
        };
        const int ld0 = m->ld[0];
        job16(p.DZ16[0], S.FEAT16, 256, grad->db[0], reduce_job(grad->dW[0], ld0, 0, 256, 256, CIN_FEAT0));
        job16(p.DZ16[0], S.ENC16, 128, nullptr, reduce_job(grad->dW[0], ld0, 0, 256, ENC_DIM, 0, CIN_FEAT0, CIN_NRM0));
        for (int l = 1; l <= 3; ++l)
            job16(p.DZ16[l], S.R16[l - 1], 256, grad->db[l], reduce_job(grad->dW[l], m->ld[l], 0, 256, 256));
        dp.n_jobs = k;
        HN_PROPAGATE(launch_dw16(dp, rp, s));
        if (grad->dW[4] || grad->db[4]) {
            const int splits = std::max(1, std::min(p.n_tiles, 64));
            color_out_grad16_kernel<<<dim3(32, splits), 128, 0, s>>>(S.R16[3], color_dw_x3() ? S.R16[3] + S.lo_off : nullptr, p.DZ4, n, p.n_tiles,
                                                                     grad->dW[4], m->ld[4], grad->db[4]);
            count_launch();
            HN_CHECK_LAUNCH();
        }
        return HN_OK;
    }
    // ---- weight gradients: dW_l = DZ_l^T a_{l-1} for all five layers in one launch -------------------------------
    DwParams dp;
    DwReduceParams rp;
    dp.n = n; dp.n_tiles = p.n_tiles; dp.part = part; rp.part = part;
    int k = 0;
    auto job = [&](DwOperand P, DwOperand Q, float* db, DwReduceJob r) {
        DwJob& j = dp.job[k];
        j.P[0] = P; j.Q[0] = Q; j.P[1] = P; j.Q[1] = Q;
        j.n_pairs = 1;
        j.n_mma = (int)round_up(Q.cols, 16);
        j.db = db; j.db_scale = 1.0f;
        rp.job[k] = r;
        ++k;
    };
    const int ld0 = m->ld[0];
    job({p.DZ[0], 0, 256, 1}, {FEAT, 0, 256, 1}, grad->db[0], reduce_job(grad->dW[0], ld0, 0, 256, 256, CIN_FEAT0));
    job({p.DZ[0], 0, 256, 1}, {p.ENC, ENC_LD, ENC_DIM, 2}, nullptr,
        reduce_job(grad->dW[0], ld0, 0, 256, ENC_DIM, 0, CIN_FEAT0, CIN_NRM0));
    for (int l = 1; l <= 3; ++l)
        job({p.DZ[l], 0, 256, 1}, {p.R[l - 1], 0, 256, 1}, grad->db[l], reduce_job(grad->dW[l], m->ld[l], 0, 256, 256));
    job({p.DZ4, 4, 3, 0}, {p.R[3], 0, 256, 1}, grad->db[4], reduce_job(grad->dW[4], m->ld[4], 0, 3, 256));
    dp.n_jobs = k;
    return launch_dw(dp, rp, s);
}

}  // namespace chain
}  // namespace hn

using namespace hn;

extern "C" {

int64_t hn_color_obj_chain_bytes(void) { return (int64_t)chain::color_layout().total; }

int hn_color_obj_chain_pack(const hn_mlp_t* m, void* chain_buf, int64_t chain_bytes, hn_stream_t stream) {
    HN_REQUIRE(m && m->n_layers == 5, "hn_color_obj_chain_pack: object colour mlp must have 5 layers");
    const chain::ColorLayout L = chain::color_layout();
    HN_REQUIRE(chain_buf && chain_bytes >= (int64_t)L.total && aligned16(chain_buf),
               "hn_color_obj_chain_pack: buffer too small or misaligned (need %u bytes)", L.total);
    static const int in_d[5] = {373, 256, 256, 256, 256};
    static const int out_d[5] = {256, 256, 256, 256, 3};
    for (int l = 0; l < 5; ++l)
        HN_REQUIRE(m->in_dim[l] == in_d[l] && m->out_dim[l] == out_d[l] && m->W[l] && m->WT[l],
                   "hn_color_obj_chain_pack: layer %d has the wrong shape or no transposed copy", l);
    cudaStream_t s = (cudaStream_t)stream;
    uint8_t* dst = reinterpret_cast<uint8_t*>(chain_buf);
    using chain::PackMap;
    using chain::launch_pack_b;
    const int big = 1 << 30;
    chain::pack_batch_begin();
    // a @ W^T operands: B(n = out, k = in)
    HN_PROPAGATE(launch_pack_b(m->W[0], m->ld[0], 0, chain::CIN_FEAT0, 256, 256, 256, 4, dst + L.nt0a, s));
    HN_PROPAGATE(launch_pack_b(m->W[0], m->ld[0], PackMap{0, big, 0, 0, chain::CIN_FEAT0, chain::CIN_NRM0}, 256, chain::ENC_DIM, 256, 2,
                               dst + L.nt0b, s));
    for (int l = 1; l <= 3; ++l) HN_PROPAGATE(launch_pack_b(m->W[l], m->ld[l], 0, 0, 256, 256, 256, 4, dst + L.nt[l], s));
    HN_PROPAGATE(launch_pack_b(m->W[4], m->ld[4], 0, 0, 3, 256, 16, 4, dst + L.nt[4], s));
    // d @ W operands: B(n = in, k = out) from the transposed copies
    HN_PROPAGATE(launch_pack_b(m->WT[4], m->ldT[4], 0, 0, 256, 3, 256, 1, dst + L.nn[4], s));
    for (int l = 3; l >= 1; --l) HN_PROPAGATE(launch_pack_b(m->WT[l], m->ldT[l], 0, 0, 256, 256, 256, 4, dst + L.nn[l], s));
    HN_PROPAGATE(launch_pack_b(m->WT[0], m->ldT[0], chain::CIN_FEAT0, 0, 256, 256, 256, 4, dst + L.nn0a, s));
    HN_PROPAGATE(launch_pack_b(m->WT[0], m->ldT[0], PackMap{0, chain::CIN_FEAT0, chain::CIN_NRM0, 0, big, 0}, chain::ENC_DIM, 256, 128, 4,
                               dst + L.nn0b, s));
    return chain::pack_batch_flush(s);
}

}  // extern "C"
