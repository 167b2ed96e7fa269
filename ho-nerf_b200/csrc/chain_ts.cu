// SDFNetwork_OBJ.sdf (utils/fields.py:316-331) as a tile-chain kernel whose ACTIVATIONS LIVE IN TENSOR MEMORY:
// the A operand of every layer is written by the epilogue warps with tcgen05.st (two fp16 values per 32-bit TMEM
// column, lane = point) and consumed by the `ts` form of tcgen05.mma.  Compared with chain_obj.cu's sdf_only_kernel
// (A in shared memory) this
//   * removes the A operand's shared-memory reads, so a layer can be issued as two 128-column halves at full MMA
//     rate: while the tensor core works on the second half, the epilogue already runs the softplus of the first half
//     and holds the packed result in registers (A may only be overwritten once all MMAs of the layer have read it);
//   * frees the 128 KB activation buffer: a 10-stage weight ring and an fp32 copy of the point's encoding (reused by
//     the skip connection) take its place.
// TMEM columns: [0,256) accumulator, [256,384) A_hi, [384,512) A_lo.
#include <algorithm>

#include "chain_obj_layout.cuh"

namespace hn {
namespace chain {

constexpr int TS_STAGE_BYTES = 128 * 128;        // one half operand k-block: [128 rows x 64 k] fp16
constexpr int TS_STAGES = 10;
constexpr int TS_ENC_LD = 65;                    // fp32 encoding scratch [128][65] (padded: conflict-free columns)
constexpr int TS_ENC_OFF = TS_STAGES * TS_STAGE_BYTES;
constexpr int TS_HEAD_OFF = TS_ENC_OFF + TILE_M * TS_ENC_LD * 4;
constexpr int TS_SMEM_BYTES = TS_HEAD_OFF + EPI_CGROUPS * TILE_M * 4 + 1024;
constexpr uint32_t TS_A_HI = 256, TS_A_LO = 384;

struct BarriersTs {
    uint64_t full[TS_STAGES];
    uint64_t empty[TS_STAGES];
    uint64_t a_ready;
    uint64_t acc_full[2];    // one per N-half: a single barrier completing twice before a slow thread looks would alias its parity
    uint32_t tmem_base;
};

struct SdfTsParams {
    const float* pts;         // [n,3] points, or NULL: lattice mode
    // lattice mode (extract_geometry, utils/renderer.py:262-278): point i = (xs[i / (ny nz)], ys[(i / nz) % ny], zs[i % nz]),
    // the ij-meshgrid order of the reference; the axes are the caller's torch.linspace values, so the coordinates are the
    // reference's bits and no [n,3] point tensor is ever written or read
    const float* xs;
    const float* ys;
    const float* zs;
    int ny, nz;
    int64_t n;
    float inv_scale;
    float* sdf;
    const uint8_t* chain;
    const float* bias[9];
    const float* w_out0;
    int n_tiles;
    long long* prof;
};

// fp16 (hi, lo) words of a column pair
__device__ __forceinline__ void pack_pair(float a, float b, uint32_t& hi, uint32_t& lo) { split2_lo16(a, b, hi, lo); }

__global__ void __launch_bounds__(THREADS, 1)
sdf_only_ts_kernel(const __grid_constant__ SdfTsParams p, const __grid_constant__ Program prog) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    __shared__ BarriersTs bar;
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    float* s_enc = reinterpret_cast<float*>(smem + TS_ENC_OFF);
    float* s_head = reinterpret_cast<float*>(smem + TS_HEAD_OFF);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (warp == 0) tc::tmem_alloc(&bar.tmem_base, 512);
    if (threadIdx.x == 32) {
        for (int s = 0; s < TS_STAGES; ++s) {
            tc::mbar_init(&bar.full[s], 1);
            tc::mbar_init(&bar.empty[s], 1);
        }
        tc::mbar_init(&bar.a_ready, EPI_THREADS);
        tc::mbar_init(&bar.acc_full[0], 1);
        tc::mbar_init(&bar.acc_full[1], 1);
        tc::mbar_fence_init();
    }
    tc::tc_fence_before_sync();
    __syncthreads();
    tc::tc_fence_after_sync();
    const uint32_t tmem = bar.tmem_base;
    const int n_my_tiles = p.n_tiles > (int)blockIdx.x ? (p.n_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x : 0;

    if (warp == 0) {
        // ---- weight producer ---------------------------------------------------------------------------------
        if (lane == 0) {
            uint32_t stage = 0, phase = 0;
            for (int t = 0; t < n_my_tiles; ++t)
                for (int s = 0; s < prog.n_steps; ++s) {
                    const Step st = prog.step[s];
                    const uint8_t* src = p.chain + st.b_off;
                    for (int c = 0; c < 2 * st.kblocks; ++c) {
                        tc::mbar_wait(&bar.empty[stage], phase ^ 1u);
                        tc::mbar_arrive_expect_tx(&bar.full[stage], TS_STAGE_BYTES);
                        tc::bulk_g2s(smem + stage * TS_STAGE_BYTES, src + (size_t)c * TS_STAGE_BYTES, TS_STAGE_BYTES, &bar.full[stage]);
                        if (++stage == TS_STAGES) { stage = 0; phase ^= 1u; }
                    }
                }
        }
    } else if (warp == 1) {
        // ---- MMA issuer: A from tensor memory ---------------------------------------------------------------------
        if (lane == 0) {
            long long t_a = 0, t_w = 0, t0 = clock64(), tt;
            const uint32_t ring = tc::smem_u32(smem);
            const uint32_t idesc = tc::make_idesc(tc::FMT_F16, 128, 128);
            uint32_t stage = 0, phase = 0, a_par = 0;
            for (int t = 0; t < n_my_tiles; ++t)
                for (int s = 0; s < prog.n_steps; ++s) {
                    const Step st = prog.step[s];
                    const uint32_t d = tmem + st.acc_col;
                    if (!st.no_wait) {
                        tt = clock64();
                        tc::mbar_wait(&bar.a_ready, a_par);
                        t_a += clock64() - tt;
                        a_par ^= 1u;
                        tc::tc_fence_after_sync();
                    }
                    for (int kb = 0; kb < st.kblocks; ++kb) {
                        const uint32_t ah = tmem + TS_A_HI + (uint32_t)kb * 32, al = tmem + TS_A_LO + (uint32_t)kb * 32;
                        tt = clock64();
                        tc::mbar_wait(&bar.full[stage], phase);
                        t_w += clock64() - tt;
                        tc::tc_fence_after_sync();
                        uint64_t dB = tc::make_smem_desc_sw128(ring + stage * TS_STAGE_BYTES);
#pragma unroll
                        for (int k = 0; k < 4; ++k) {
                            tc::umma_f16_ts(d, al + 8 * k, dB + 2 * k, idesc, (kb | k) != 0);
                            tc::umma_f16_ts(d, ah + 8 * k, dB + 2 * k, idesc, 1);
                        }
                        tc::umma_commit(&bar.empty[stage]);
                        if (++stage == TS_STAGES) { stage = 0; phase ^= 1u; }
                        tt = clock64();
                        tc::mbar_wait(&bar.full[stage], phase);
                        t_w += clock64() - tt;
                        tc::tc_fence_after_sync();
                        dB = tc::make_smem_desc_sw128(ring + stage * TS_STAGE_BYTES);
#pragma unroll
                        for (int k = 0; k < 4; ++k) tc::umma_f16_ts(d, ah + 8 * k, dB + 2 * k, idesc, 1);
                        tc::umma_commit(&bar.empty[stage]);
                        if (++stage == TS_STAGES) { stage = 0; phase ^= 1u; }
                    }
                    tc::umma_commit(&bar.acc_full[st.acc_col ? 1 : 0]);
                }
            if (p.prof) {
                p.prof[blockIdx.x * 4 + 0] = t_a;
                p.prof[blockIdx.x * 4 + 1] = t_w;
                p.prof[blockIdx.x * 4 + 2] = clock64() - t0;
            }
        }
    } else {
        // ---- epilogue warps: thread = (row, 32-column group of each half) ---------------------------------------------
        const int row = (warp & 3) * 32 + lane, cg = (warp - 2) >> 2;
        const uint32_t lane_base = (uint32_t)(row & ~31) << 16;
        uint32_t acc_par[2] = {0u, 0u};
        auto publish = [&]() {
            tc::tmem_st_wait();
            tc::tc_fence_before_sync();
            tc::mbar_arrive(&bar.a_ready);
        };
        auto wait_acc = [&](int hf) {
            tc::mbar_wait(&bar.acc_full[hf], acc_par[hf]);
            acc_par[hf] ^= 1u;
            tc::tc_fence_after_sync();
        };
        // 32 activation values of columns [col0, col0+32) -> 16 (hi, lo) words
        auto pack32 = [&](const float* v, uint32_t* hi, uint32_t* lo) {
#pragma unroll
            for (int i = 0; i < 16; ++i) pack_pair(v[2 * i], v[2 * i + 1], hi[i], lo[i]);
        };
        auto store_a = [&](int col0, const uint32_t* hi, const uint32_t* lo) {
            const uint32_t c = (uint32_t)(col0 >> 1);
            tc::tmem_st_32x32b_x8(tmem + lane_base + TS_A_HI + c, hi);
            tc::tmem_st_32x32b_x8(tmem + lane_base + TS_A_HI + c + 8, hi + 8);
            tc::tmem_st_32x32b_x8(tmem + lane_base + TS_A_LO + c, lo);
            tc::tmem_st_32x32b_x8(tmem + lane_base + TS_A_LO + c + 8, lo + 8);
        };
        auto activate32 = [&](int acc_col0, int col0, const float* __restrict__ bias, float* v) {
            acc_load32(tmem, row, acc_col0, v);
#pragma unroll
            for (int j = 0; j < 32; j += 4) {
                const float4 b = __ldg(reinterpret_cast<const float4*>(bias + col0 + j));
                v[j] = softplus100_fast(v[j] + b.x);
                v[j + 1] = softplus100_fast(v[j + 1] + b.y);
                v[j + 2] = softplus100_fast(v[j + 2] + b.z);
                v[j + 3] = softplus100_fast(v[j + 3] + b.w);
            }
        };
        for (int t = 0; t < n_my_tiles; ++t) {
            const int64_t tile = (int64_t)blockIdx.x + (int64_t)t * gridDim.x;
            const int64_t gp = tile * TILE_M + row;
            // ---- encoding -> shared scratch (fp32, kept for the skip connection) -> packed first-layer operand --------
            {
                float x[3] = {0.f, 0.f, 0.f};
                if (gp < p.n) {
                    if (p.pts) {
                        x[0] = p.pts[gp * 3]; x[1] = p.pts[gp * 3 + 1]; x[2] = p.pts[gp * 3 + 2];
                    } else {
                        const int64_t iyz = gp / p.nz;
                        x[2] = __ldg(p.zs + (gp - iyz * p.nz));
                        const int64_t ix = iyz / p.ny;
                        x[1] = __ldg(p.ys + (iyz - ix * p.ny));
                        x[0] = __ldg(p.xs + ix);
                    }
                }
                float* e = s_enc + row * TS_ENC_LD;
                if (cg == 0) {
                    e[0] = x[0]; e[1] = x[1]; e[2] = x[2]; e[63] = 0.0f;
                }
                for (int idx = cg; idx < 30; idx += EPI_CGROUPS) {
                    const int c = idx / 10, k = idx - c * 10;
                    float s, co;
                    sincosf(x[c] * (float)(1 << k), &s, &co);
                    e[3 + c * 20 + k] = s;
                    e[3 + c * 20 + 10 + k] = co;
                }
                tc::named_bar_sync(1, EPI_THREADS);
                uint32_t hi[8], lo[8];
#pragma unroll
                for (int i = 0; i < 8; ++i) pack_pair(e[cg * 16 + 2 * i], e[cg * 16 + 2 * i + 1], hi[i], lo[i]);
                tc::tmem_st_32x32b_x8(tmem + lane_base + TS_A_HI + (uint32_t)(cg * 8), hi);
                tc::tmem_st_32x32b_x8(tmem + lane_base + TS_A_LO + (uint32_t)(cg * 8), lo);
            }
            publish();
            float head = 0.0f;
            for (int l = 0; l < 8; ++l) {
                const float* __restrict__ bias = p.bias[l];
                float v[32];
                uint32_t hh[16], hl[16];
                // ---- first half (columns cg*32 ..), under the second half's MMAs ----------------------------------------
                wait_acc(0);
                activate32(cg * 32, cg * 32, bias, v);
                if (l < 7) {
                    pack32(v, hh, hl);
                } else {
#pragma unroll
                    for (int j = 0; j < 32; j += 4) {
                        const float4 w = __ldg(reinterpret_cast<const float4*>(p.w_out0 + cg * 32 + j));
                        head += v[j] * w.x + v[j + 1] * w.y + v[j + 2] * w.z + v[j + 3] * w.w;
                    }
                }
                // ---- second half: every MMA of the layer has read A, it may be overwritten ---------------------------------
                wait_acc(1);
                if (l < 7) store_a(cg * 32, hh, hl);
                const int col0 = 128 + cg * 32;
                if (l == 3 && col0 >= 192) {
                    // skip input, columns 192..255 = [h3[192], e_0 .. e_62]
                    const float* e = s_enc + row * TS_ENC_LD;
                    if (col0 == 192) {
                        float a[32];
                        acc_load32(tmem, row, 192, a);
                        v[0] = softplus100_fast(a[0] + __ldg(bias + 192));
#pragma unroll
                        for (int j = 1; j < 32; ++j) v[j] = e[j - 1];
                    } else {
#pragma unroll
                        for (int j = 0; j < 32; ++j) v[j] = e[31 + j];
                    }
                } else {
                    activate32(col0, col0, bias, v);
                }
                if (l < 7) {
                    pack32(v, hh, hl);
                    store_a(col0, hh, hl);
                    publish();
                } else {
#pragma unroll
                    for (int j = 0; j < 32; j += 4) {
                        const float4 w = __ldg(reinterpret_cast<const float4*>(p.w_out0 + col0 + j));
                        head += v[j] * w.x + v[j + 1] * w.y + v[j + 2] * w.z + v[j + 3] * w.w;
                    }
                }
            }
            s_head[cg * TILE_M + row] = head;
            tc::tc_fence_before_sync();
            tc::named_bar_sync(1, EPI_THREADS);
            if (cg == 0 && gp < p.n) {
                float acc = 0.0f;
#pragma unroll
                for (int g = 0; g < EPI_CGROUPS; ++g) acc += s_head[g * TILE_M + row];
                p.sdf[gp] = (acc + __ldg(p.bias[8])) * p.inv_scale;
            }
        }
    }
    tc::tc_fence_before_sync();
    __syncthreads();
    if (warp == 0) tc::tmem_dealloc(tmem, 512);
}

static long long* g_prof_ts = nullptr;
void set_prof_ts(long long* p) { g_prof_ts = p; }

int launch_sdf_only_ts(const hn_mlp_t* m, const float* pts, int64_t n, float inv_scale, float* sdf, cudaStream_t s,
                       const float* xs, const float* ys, const float* zs, int ny, int nz) {
    const ObjLayout L = obj_layout();
    SdfTsParams p;
    p.pts = pts; p.n = n; p.inv_scale = inv_scale; p.sdf = sdf;
    p.xs = xs; p.ys = ys; p.zs = zs; p.ny = ny; p.nz = nz;
    p.chain = reinterpret_cast<const uint8_t*>(m->chain);
    for (int l = 0; l < 9; ++l) p.bias[l] = m->b[l];
    p.w_out0 = m->W[8];
    p.n_tiles = (int)ceil_div(n, TILE_M);
    p.prof = g_prof_ts;
    Program prog = {};
    prog.n_steps = 16;
    for (int l = 0; l < 8; ++l)
        for (int h = 0; h < 2; ++h) {
            Step& st = prog.step[2 * l + h];
            st.b_off = L.nth_off[l][h];
            st.n_mma = 128;
            st.kblocks = L.nt_kb[l];
            st.f16 = 1;
            st.no_wait = (uint8_t)h;
            st.acc_col = (uint16_t)(128 * h);
        }
    static bool configured = false;
    if (!configured) {
        HN_CHECK_CUDA(cudaFuncSetAttribute(sdf_only_ts_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, TS_SMEM_BYTES));
        configured = true;
    }
    const int grid = std::min(p.n_tiles, sm_count());
    {
        TimingScope ts(s, TT_SDF_ONLY);
        sdf_only_ts_kernel<<<grid, THREADS, TS_SMEM_BYTES, s>>>(p, prog);
    }
    count_launch();
    HN_CHECK_LAUNCH();
    return HN_OK;
}

}  // namespace chain
}  // namespace hn
