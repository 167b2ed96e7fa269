// Object SDF field, HN_TC_MIXED16 (see chain16.cuh):
//   trunk16_kernel   SDFNetwork_OBJ.forward (utils/fields.py:316-331): value trunk + feature head, three fp16 MMAs per
//                    product, activations in tensor memory (the design of chain_ts.cu); leaves EM = exp(-100 h) (fp16),
//                    A16 = h (bf16), the encoding E (fp32) / E16 (bf16) in HBM
//   nsweep16_kernel  SDFNetwork_OBJ.gradient (utils/fields.py:336-347) as the analytic normal sweep
//                    D_{l-1} = s'(h_{l-1}) * (D_l W_l), normal = J_e^T (D_0 W_0 + skip part): fp16 A operand in tensor
//                    memory (double-buffered), two MMAs per product; leaves D16 (bf16) and EB (fp32)
//   bwd16_kernel     the second-order backward of (sdf, feature, normal): tangent sweep + reverse sweep, bf16 A operand
//                    in tensor memory, two MMAs per product; leaves U16 / X16 / DZ16 / DF16 / UE16 (bf16) and d_pts
//   + chain16_dw.cu  all weight / bias gradients from those tiles in one launch
#include <algorithm>

#include "chain16.cuh"
#include "chain_dw.cuh"
#include "chain_obj_layout.cuh"

namespace hn {
namespace chain {

// ------------------------------------------------------------------------------------------------------------------
// trunk: activations (fp16 hi + lo) in tensor memory, TMEM columns [0,256) accumulator, [256,384) A_hi, [384,512) A_lo
// ------------------------------------------------------------------------------------------------------------------
constexpr int TR_STAGE_BYTES = 128 * 128;
constexpr int TR_STAGES = 10;
constexpr int TR_ENC_LD = 65;
constexpr int TR_ENC_OFF = TR_STAGES * TR_STAGE_BYTES;
constexpr int TR_HEAD_OFF = TR_ENC_OFF + TILE_M * TR_ENC_LD * 4;
constexpr int TR_SMEM_BYTES = TR_HEAD_OFF + EPI_CGROUPS * TILE_M * 4 + 1024;
constexpr uint32_t TR_A_HI = 256, TR_A_LO = 384;

struct TrBarriers {
    uint64_t full[TR_STAGES];
    uint64_t empty[TR_STAGES];
    uint64_t a_ready;
    uint64_t acc_full[2];    // one per N-half (see SwBarriers)
    uint32_t tmem_base;
};

struct Trunk16Params {
    uint32_t* dbg;
    const float* pts;
    int64_t n;
    float inv_scale;
    float* sdf;
    float* feat;
    int64_t ld_feat;
    float* E;            // fp32 column-major [64][128] tiles (eoff)
    uint8_t* E16;        // bf16 T16N tiles
    uint8_t* EM[8];      // fp16 T16 tiles: em = exp(-100 h) rounded to fp16
    uint8_t* EML[8];     // fp16 T16 tiles: fp16(em - fp16(em)), read by the normal sweep only
    uint8_t* A16[8];     // bf16 T16 tiles
    const uint8_t* chain;
    const float* bias[9];
    const float* w_out0;
    int n_tiles;
};

__global__ void __launch_bounds__(THREADS, 1)
trunk16_kernel(const __grid_constant__ Trunk16Params p, const __grid_constant__ Program prog) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    __shared__ TrBarriers bar;
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    float* s_enc = reinterpret_cast<float*>(smem + TR_ENC_OFF);
    float* s_head = reinterpret_cast<float*>(smem + TR_HEAD_OFF);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (warp == 0) tc::tmem_alloc(&bar.tmem_base, 512);
    if (threadIdx.x == 32) {
        for (int s = 0; s < TR_STAGES; ++s) {
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
        if (lane == 0) {
            uint32_t stage = 0, phase = 0;
            for (int t = 0; t < n_my_tiles; ++t)
                for (int s = 0; s < prog.n_steps; ++s) {
                    const Step st = prog.step[s];
                    const uint8_t* src = p.chain + st.b_off;
                    for (int c = 0; c < 2 * st.kblocks; ++c) {
                        tc::mbar_wait(&bar.empty[stage], phase ^ 1u);
                        tc::mbar_arrive_expect_tx(&bar.full[stage], TR_STAGE_BYTES);
                        tc::bulk_g2s(smem + stage * TR_STAGE_BYTES, src + (size_t)c * TR_STAGE_BYTES, TR_STAGE_BYTES, &bar.full[stage]);
                        if (++stage == TR_STAGES) { stage = 0; phase ^= 1u; }
                    }
                }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            const uint32_t ring = tc::smem_u32(smem);
            const uint32_t idesc = tc::make_idesc(tc::FMT_F16, 128, 128);
            uint32_t stage = 0, phase = 0, a_par = 0;
            for (int t = 0; t < n_my_tiles; ++t)
                for (int s = 0; s < prog.n_steps; ++s) {
                    dbg_mark(p.dbg, 1, (uint32_t)(t << 8 | s));
                    const Step st = prog.step[s];
                    const uint32_t d = tmem + st.acc_col;
                    if (!st.no_wait) {
                        tc::mbar_wait(&bar.a_ready, a_par);
                        a_par ^= 1u;
                        tc::tc_fence_after_sync();
                    }
                    for (int kb = 0; kb < st.kblocks; ++kb) {
                        const uint32_t ah = tmem + TR_A_HI + (uint32_t)kb * 32, al = tmem + TR_A_LO + (uint32_t)kb * 32;
                        tc::mbar_wait(&bar.full[stage], phase);
                        tc::tc_fence_after_sync();
                        uint64_t dB = tc::make_smem_desc_sw128(ring + stage * TR_STAGE_BYTES);
#pragma unroll
                        for (int k = 0; k < 4; ++k) {
                            tc::umma_f16_ts(d, al + 8 * k, dB + 2 * k, idesc, (kb | k) != 0);
                            tc::umma_f16_ts(d, ah + 8 * k, dB + 2 * k, idesc, 1);
                        }
                        tc::umma_commit(&bar.empty[stage]);
                        if (++stage == TR_STAGES) { stage = 0; phase ^= 1u; }
                        tc::mbar_wait(&bar.full[stage], phase);
                        tc::tc_fence_after_sync();
                        dB = tc::make_smem_desc_sw128(ring + stage * TR_STAGE_BYTES);
#pragma unroll
                        for (int k = 0; k < 4; ++k) tc::umma_f16_ts(d, ah + 8 * k, dB + 2 * k, idesc, 1);
                        tc::umma_commit(&bar.empty[stage]);
                        if (++stage == TR_STAGES) { stage = 0; phase ^= 1u; }
                    }
                    tc::umma_commit(&bar.acc_full[st.acc_col ? 1 : 0]);
                }
            dbg_mark(p.dbg, 1, 0xffffffffu);
        }
    } else {
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
        auto store_a = [&](int col0, const uint32_t* hi, const uint32_t* lo) {
            const uint32_t c = (uint32_t)(col0 >> 1);
            tc::tmem_st_32x32b_x8(tmem + lane_base + TR_A_HI + c, hi);
            tc::tmem_st_32x32b_x8(tmem + lane_base + TR_A_HI + c + 8, hi + 8);
            tc::tmem_st_32x32b_x8(tmem + lane_base + TR_A_LO + c, lo);
            tc::tmem_st_32x32b_x8(tmem + lane_base + TR_A_LO + c + 8, lo + 8);
        };
        for (int t = 0; t < n_my_tiles; ++t) {
            const int64_t tile = (int64_t)blockIdx.x + (int64_t)t * gridDim.x;
            const int64_t gp = tile * TILE_M + row;
            // 8 columns [c, c + 8) of layer l: stash chunks (EM fp16, A16 bf16) + fp16 hi / lo words of the next A operand
            auto emit8 = [&](int l, int c, const float* h, const float* em, uint32_t* hi4, uint32_t* lo4) {
                const uint32_t off = t16_off(row, c >> 3);
                uint4 q, ql;
                split2_lo16(em[0], em[1], q.x, ql.x); split2_lo16(em[2], em[3], q.y, ql.y);
                split2_lo16(em[4], em[5], q.z, ql.z); split2_lo16(em[6], em[7], q.w, ql.w);
                stg16(p.EM[l] + (size_t)tile * T16_TILE_BYTES + off, q);
                stg16(p.EML[l] + (size_t)tile * T16_TILE_BYTES + off, ql);
                q.x = pack_bf16x2(h[0], h[1]); q.y = pack_bf16x2(h[2], h[3]); q.z = pack_bf16x2(h[4], h[5]); q.w = pack_bf16x2(h[6], h[7]);
                stg16(p.A16[l] + (size_t)tile * T16_TILE_BYTES + off, q);
#pragma unroll
                for (int i = 0; i < 4; ++i) split2_lo16(h[2 * i], h[2 * i + 1], hi4[i], lo4[i]);
            };
            // softplus of accumulator columns [acc_col0, +32) = layer columns [col0, +32)
            auto act_half = [&](int l, int col0, uint32_t* hh, uint32_t* hl, float& head) {
                const float* __restrict__ bias = p.bias[l];
                float v[32];
                acc_load32(tmem, row, col0, v);
#pragma unroll
                for (int j = 0; j < 32; j += 8) {
                    const float4 b0 = __ldg(reinterpret_cast<const float4*>(bias + col0 + j));
                    const float4 b1 = __ldg(reinterpret_cast<const float4*>(bias + col0 + j + 4));
                    const float bb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
                    float h[8], em[8];
#pragma unroll
                    for (int i = 0; i < 8; ++i) h[i] = softplus100_em(v[j + i] + bb[i], em[i]);
                    emit8(l, col0 + j, h, em, hh + (j >> 1), hl + (j >> 1));
                    if (l == 7) {
                        const float4 w0 = __ldg(reinterpret_cast<const float4*>(p.w_out0 + col0 + j));
                        const float4 w1 = __ldg(reinterpret_cast<const float4*>(p.w_out0 + col0 + j + 4));
                        head += h[0] * w0.x + h[1] * w0.y + h[2] * w0.z + h[3] * w0.w + h[4] * w1.x + h[5] * w1.y + h[6] * w1.z + h[7] * w1.w;
                    }
                }
            };
            // ---- encoding -> shared scratch (fp32, kept for the skip connection), E / E16 stash, first-layer operand ------
            {
                float x[3] = {0.f, 0.f, 0.f};
                if (gp < p.n) { x[0] = p.pts[gp * 3]; x[1] = p.pts[gp * 3 + 1]; x[2] = p.pts[gp * 3 + 2]; }
                float* e = s_enc + row * TR_ENC_LD;
                float* __restrict__ ge = p.E + eoff(gp);
                if (cg == 0) {
                    e[0] = x[0]; e[1] = x[1]; e[2] = x[2]; e[63] = 0.0f;
                    ge[0] = x[0]; ge[TILE_M] = x[1]; ge[2 * TILE_M] = x[2]; ge[63 * TILE_M] = 0.0f;
                }
                for (int idx = cg; idx < 30; idx += EPI_CGROUPS) {
                    const int c = idx / 10, k = idx - c * 10;
                    float s, co;
                    sincosf(x[c] * (float)(1 << k), &s, &co);
                    e[3 + c * 20 + k] = s;
                    e[3 + c * 20 + 10 + k] = co;
                    ge[(3 + c * 20 + k) * TILE_M] = s;
                    ge[(3 + c * 20 + 10 + k) * TILE_M] = co;
                }
                tc::named_bar_sync(1, EPI_THREADS);
                uint32_t hi[8], lo[8];
#pragma unroll
                for (int i = 0; i < 8; ++i) split2_lo16(e[cg * 16 + 2 * i], e[cg * 16 + 2 * i + 1], hi[i], lo[i]);
                tc::tmem_st_32x32b_x8(tmem + lane_base + TR_A_HI + (uint32_t)(cg * 8), hi);
                tc::tmem_st_32x32b_x8(tmem + lane_base + TR_A_LO + (uint32_t)(cg * 8), lo);
#pragma unroll
                for (int c2 = 0; c2 < 2; ++c2) {
                    const float* s = e + cg * 16 + c2 * 8;
                    uint4 q;
                    q.x = pack_bf16x2(s[0], s[1]); q.y = pack_bf16x2(s[2], s[3]); q.z = pack_bf16x2(s[4], s[5]); q.w = pack_bf16x2(s[6], s[7]);
                    stg16(p.E16 + (size_t)tile * T16N_TILE_BYTES + t16_off<8>(row, cg * 2 + c2), q);
                }
            }
            publish();
            float head = 0.0f;
            for (int l = 0; l < 8; ++l) {
                uint32_t hh[16], hl[16];
                if (threadIdx.x == 64) dbg_mark(p.dbg, 2, (uint32_t)(t << 8 | l));
                // ---- first half (columns cg*32 ..), under the second half's MMAs ----------------------------------------
                wait_acc(0);
                act_half(l, cg * 32, hh, hl, head);
                // ---- second half: every MMA of the layer has read A, it may be overwritten ---------------------------------
                wait_acc(1);
                store_a(cg * 32, hh, hl);
                const int col0 = 128 + cg * 32;
                if (l == 3 && col0 >= 192) {
                    // skip input, columns 192..255 = [h3[192], e_0 .. e_62]
                    const float* e = s_enc + row * TR_ENC_LD;
                    float v[32], em0 = 0.0f;
                    if (col0 == 192) {
                        float a[32];
                        acc_load32(tmem, row, 192, a);
                        v[0] = softplus100_em(a[0] + __ldg(p.bias[3] + 192), em0);
#pragma unroll
                        for (int j = 1; j < 32; ++j) v[j] = e[j - 1];
                    } else {
#pragma unroll
                        for (int j = 0; j < 32; ++j) v[j] = e[31 + j];
                    }
#pragma unroll
                    for (int j = 0; j < 32; j += 8) {
                        float em[8] = {j == 0 ? em0 : 0.0f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
                        emit8(l, col0 + j, v + j, em, hh + (j >> 1), hl + (j >> 1));
                    }
                } else {
                    act_half(l, col0, hh, hl, head);
                }
                store_a(col0, hh, hl);
                publish();
            }
            // sdf = (h7 . W_out[0] + b_out[0]) / scale
            s_head[cg * TILE_M + row] = head;
            tc::named_bar_sync(1, EPI_THREADS);
            if (cg == 0 && gp < p.n) {
                float acc = 0.0f;
#pragma unroll
                for (int g = 0; g < EPI_CGROUPS; ++g) acc += s_head[g * TILE_M + row];
                p.sdf[gp] = (acc + __ldg(p.bias[8])) * p.inv_scale;
            }
            // ---- feature head: rows 1..256 of the output layer, no activation ----------------------------------------------
            for (int hf = 0; hf < 2; ++hf) {
                wait_acc(hf);
                const int col0 = 128 * hf + cg * 32;
                float v[32];
                acc_load32(tmem, row, col0, v);
                if (gp < p.n) {
                    const float* __restrict__ bias = p.bias[8] + 1;
                    float* __restrict__ fr = p.feat + gp * p.ld_feat + col0;
#pragma unroll
                    for (int j = 0; j < 32; j += 4)
                        st4(fr + j, make_float4(v[j] + __ldg(bias + col0 + j), v[j + 1] + __ldg(bias + col0 + j + 1),
                                                v[j + 2] + __ldg(bias + col0 + j + 2), v[j + 3] + __ldg(bias + col0 + j + 3)));
                }
            }
            tc::tc_fence_before_sync();       // the accumulator reads above precede the next tile's MMAs (ordered by its a_ready)
        }
        if (threadIdx.x == 64) dbg_mark(p.dbg, 2, 0xffffffffu);
    }
    tc::tc_fence_before_sync();
    __syncthreads();
    if (warp == 0) tc::tmem_dealloc(tmem, 512);
}

// ------------------------------------------------------------------------------------------------------------------
// normal sweep
// ------------------------------------------------------------------------------------------------------------------
struct Nsweep16Params {
    uint32_t* dbg;
    int64_t n;
    float inv_scale;
    const float* E;
    uint8_t* D16[8];
    float* EB;
    float* normal;
    const uint8_t* chain;
    const float* w_out0;
    int n_tiles;
};

__global__ void __launch_bounds__(SW_THREADS, 1)
nsweep16_kernel(const __grid_constant__ Nsweep16Params p, const __grid_constant__ SwProgram prog,
                const __grid_constant__ SwInputs inputs) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    __shared__ SwBarriers bar;
    uint8_t* smem = sw_setup(smem_raw, &bar);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int n_my_tiles = p.n_tiles > (int)blockIdx.x ? (p.n_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x : 0;
    if (warp == 0) {
        if (lane == 0) sw_producer(prog, p.chain, smem, &bar, n_my_tiles, p.dbg);
    } else if (warp == 1) {
        if (lane == 0) sw_mma(prog, smem, &bar, n_my_tiles, p.dbg);
    } else if (warp == 2 + EPI_WARPS) {
        if (lane == 0) sw_in_producer(inputs, smem, &bar, n_my_tiles, p.dbg);
    } else {
        const int row = (warp & 3) * 32 + lane, cg = (warp - 2) >> 2;
        const uint32_t lane_base = (uint32_t)(row & ~31) << 16;
        const uint32_t tmem = bar.tmem_base;
        uint32_t acc_par = 0, in_par[2] = {0u, 0u};
        for (int t = 0; t < n_my_tiles; ++t) {
            const int64_t tile = (int64_t)blockIdx.x + (int64_t)t * gridDim.x;
            const int64_t gp = tile * TILE_M + row;
            const bool live = gp < p.n;
            const size_t tb = (size_t)tile * T16_TILE_BYTES;
            // 16 columns [c, c + 16) of D_lyr -> fp16 hi / lo halves of the next A operand (tensor memory), bf16 chunks of D16
            auto emit16 = [&](int lyr, int c, const float* d) {
                uint32_t hi8[8], lo8[8];
#pragma unroll
                for (int i = 0; i < 8; ++i) split2_lo16(d[2 * i], d[2 * i + 1], hi8[i], lo8[i]);
                sw_st16(tmem, lane_base, SW_A0, c, hi8);
                sw_st16(tmem, lane_base, SW_A1, c, lo8);
                uint4 q0, q1;
                q0.x = pack_bf16x2(d[0], d[1]); q0.y = pack_bf16x2(d[2], d[3]); q0.z = pack_bf16x2(d[4], d[5]); q0.w = pack_bf16x2(d[6], d[7]);
                q1.x = pack_bf16x2(d[8], d[9]); q1.y = pack_bf16x2(d[10], d[11]); q1.z = pack_bf16x2(d[12], d[13]); q1.w = pack_bf16x2(d[14], d[15]);
                stg16(p.D16[lyr] + tb + t16_off(row, c >> 3), q0);
                stg16(p.D16[lyr] + tb + t16_off(row, (c >> 3) + 1), q1);
            };
            auto load_sp16 = [&](int hf, int c, float* sp) {       // s' = 1 - (em_hi + em_lo) of 16 columns, from the input slot
                const uint4 a = sw_in_ld(smem, hf, 0, row, c), b = sw_in_ld(smem, hf, 0, row, c + 8);
                const uint4 al = sw_in_ld(smem, hf, 1, row, c), bl = sw_in_ld(smem, hf, 1, row, c + 8);
                const uint32_t w[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
                const uint32_t wl[8] = {al.x, al.y, al.z, al.w, bl.x, bl.y, bl.z, bl.w};
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const float2 f = f16x2_unpack(w[i]), fl = f16x2_unpack(wl[i]);
                    sp[2 * i] = 1.0f - (f.x + fl.x);
                    sp[2 * i + 1] = 1.0f - (f.y + fl.y);
                }
            };
            // ---- seed: D_7 = s'(h_7) * W_out[0] / scale (the previous tile's MMAs are complete: A may be written) -------------
#pragma unroll 1
            for (int hf = 0; hf < 2; ++hf) {
                sw_in_wait(&bar, hf, in_par);
#pragma unroll
                for (int sub = 0; sub < 2; ++sub) {
                    const int c = 128 * hf + cg * 32 + 16 * sub;
                    float sp[16], d[16];
                    load_sp16(hf, c, sp);
#pragma unroll
                    for (int j = 0; j < 16; j += 4) {
                        const float4 w = __ldg(reinterpret_cast<const float4*>(p.w_out0 + c + j));
                        d[j] = sp[j] * w.x * p.inv_scale; d[j + 1] = sp[j + 1] * w.y * p.inv_scale;
                        d[j + 2] = sp[j + 2] * w.z * p.inv_scale; d[j + 3] = sp[j + 3] * w.w * p.inv_scale;
                    }
                    if (!live) {
#pragma unroll
                        for (int j = 0; j < 16; ++j) d[j] = 0.0f;
                    }
                    emit16(7, c, d);
                }
                sw_in_release(&bar, hf);
            }
            sw_publish(&bar);
            // ---- D_{l-1} = s'(h_{l-1}) * (D_l W_l), l = 7..1 ------------------------------------------------------------------
            for (int l = 7; l >= 1; --l) {
                if (threadIdx.x == 64) dbg_mark(p.dbg, 2, (uint32_t)(t << 8 | l));
                sw_wait_acc(&bar, acc_par);
                // four 16-column sub-blocks; the accumulator load of the next one is in flight while this one is processed
                float accv[2][16];
                sw_ld16_nowait(tmem, lane_base, cg * 32, accv[0]);
#pragma unroll
                for (int sb = 0; sb < 4; ++sb) {
                    const int hf = sb >> 1;
                    const int c = 128 * hf + cg * 32 + 16 * (sb & 1);
                    if ((sb & 1) == 0) sw_in_wait(&bar, hf, in_par);
                    float sp[16];
                    load_sp16(hf, c, sp);
                    tc::tmem_ld_wait();
                    if (sb < 3) sw_ld16_nowait(tmem, lane_base, 128 * ((sb + 1) >> 1) + cg * 32 + 16 * ((sb + 1) & 1), accv[(sb + 1) & 1]);
                    float* g = accv[sb & 1];
                    if (l == 4 && c + 15 > 192) {
                        // columns 193..255 of the skip layer's input are the encoding: their cotangent is the raw product,
                        // parked in EB until the encoding layer adds its own part
#pragma unroll
                        for (int j = 0; j < 16; ++j)
                            if (c + j > 192) {
                                p.EB[eoff(gp, c + j - 193)] = g[j];
                                g[j] = 0.0f;
                            }
                    }
#pragma unroll
                    for (int j = 0; j < 16; ++j) g[j] = live ? sp[j] * g[j] : 0.0f;
                    emit16(l - 1, c, g);
                    if (sb & 1) sw_in_release(&bar, hf);
                }
                sw_publish(&bar);
            }
            // ---- encoding layer: eb = D_0 W_0 + (skip part); normal = J_e^T eb ---------------------------------------------------
            sw_wait_acc(&bar, acc_par);
            tc::named_bar_sync(1, EPI_THREADS);          // the skip part was written by other threads of this CTA
            {
                float g[16];
                sw_ld16(tmem, lane_base, cg * 16, g);
                float* __restrict__ eb = p.EB + eoff(gp, cg * 16);
#pragma unroll
                for (int j = 0; j < 16; ++j) eb[j * TILE_M] = (cg * 16 + j == 63) ? 0.0f : g[j] + eb[j * TILE_M];
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
        if (threadIdx.x == 64) dbg_mark(p.dbg, 2, 0xffffffffu);
    }
    sw_teardown(&bar);
}

// ------------------------------------------------------------------------------------------------------------------
// second-order backward: tangent sweep + reverse sweep
// ------------------------------------------------------------------------------------------------------------------
struct Bwd16Params {
    uint32_t* dbg;
    int64_t n;
    float inv_scale;
    const float* E;
    const float* EB;
    const float* d_sdf;      // may be NULL
    const float* d_feat;     // may be NULL
    int64_t ld_dfeat;
    const float* d_normal;
    float* d_pts;            // may be NULL
    float* DE;               // fp32 column-major [64][128] tiles: cotangent of the encoding
    uint8_t* U16[8];
    uint8_t* X16[8];
    uint8_t* DZ16[8];
    uint8_t* UE16;
    uint8_t* DF16;
    int store_dw;            // 0: nobody needs weight gradients (pose fitting): skip the U16 / DZ16 / UE16 / DF16 stores
    const uint8_t* chain;
    const float* w_out0;
    int n_tiles;
};

__global__ void __launch_bounds__(SW_THREADS, 1)
bwd16_kernel(const __grid_constant__ Bwd16Params p, const __grid_constant__ SwProgram prog,
             const __grid_constant__ SwInputs inputs) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    __shared__ SwBarriers bar;
    uint8_t* smem = sw_setup(smem_raw, &bar);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int n_my_tiles = p.n_tiles > (int)blockIdx.x ? (p.n_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x : 0;
    if (warp == 0) {
        if (lane == 0) sw_producer(prog, p.chain, smem, &bar, n_my_tiles, p.dbg);
    } else if (warp == 1) {
        if (lane == 0) sw_mma(prog, smem, &bar, n_my_tiles, p.dbg);
    } else if (warp == 2 + EPI_WARPS) {
        if (lane == 0) sw_in_producer(inputs, smem, &bar, n_my_tiles, p.dbg);
    } else {
        const int row = (warp & 3) * 32 + lane, cg = (warp - 2) >> 2;
        const uint32_t lane_base = (uint32_t)(row & ~31) << 16;
        const uint32_t tmem = bar.tmem_base;
        uint32_t acc_par = 0, in_par[2] = {0u, 0u};
        for (int t = 0; t < n_my_tiles; ++t) {
            const int64_t tile = (int64_t)blockIdx.x + (int64_t)t * gridDim.x;
            const int64_t gp = tile * TILE_M + row;
            const bool live = gp < p.n;
            const size_t tb = (size_t)tile * T16_TILE_BYTES;
            float dn[3] = {0.f, 0.f, 0.f};
            if (live) { dn[0] = p.d_normal[gp * 3]; dn[1] = p.d_normal[gp * 3 + 1]; dn[2] = p.d_normal[gp * 3 + 2]; }
            const float gs = (live && p.d_sdf) ? p.d_sdf[gp] * p.inv_scale : 0.0f;
            const float* __restrict__ e_pt = p.E + eoff(gp);
            auto pack16 = [&](const float* v, uint32_t* w) {
#pragma unroll
                for (int i = 0; i < 8; ++i) w[i] = pack_bf16x2(v[2 * i], v[2 * i + 1]);
            };
            auto store16 = [&](uint8_t* arr, int c, const uint32_t* w) {
                stg16(arr + tb + t16_off(row, c >> 3), make_uint4(w[0], w[1], w[2], w[3]));
                stg16(arr + tb + t16_off(row, (c >> 3) + 1), make_uint4(w[4], w[5], w[6], w[7]));
            };
            // 16 values -> bf16 hi / lo halves of the next A operand (tensor memory); hi8 is also what the 16-bit stash holds
            auto emit_a16 = [&](int c, const float* v, uint32_t* hi8) {
                uint32_t lo8[8];
#pragma unroll
                for (int i = 0; i < 8; ++i) split2(v[2 * i], v[2 * i + 1], hi8[i], lo8[i]);
                sw_st16(tmem, lane_base, SW_A0, c, hi8);
                sw_st16(tmem, lane_base, SW_A1, c, lo8);
            };
            auto unpack_bf16 = [&](const uint4& a, const uint4& b, float* v) {
                const uint32_t w[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
#pragma unroll
                for (int i = 0; i < 8; ++i) { v[2 * i] = bf16_lo(w[i]); v[2 * i + 1] = bf16_hi(w[i]); }
            };
            auto unpack_f16 = [&](const uint4& a, const uint4& b, float* v) {
                const uint32_t w[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const float2 f = f16x2_unpack(w[i]);
                    v[2 * i] = f.x;
                    v[2 * i + 1] = f.y;
                }
            };
            // ---- ue = J_e(x) dn: tangent of the encoding, 16 columns per thread ---------------------------------------------
            {
                float ue[16];
#pragma unroll
                for (int j = 0; j < 16; ++j) ue[j] = enc_tangent_col(e_pt, dn[0], dn[1], dn[2], cg * 16 + j);
                uint32_t w[8];
                emit_a16(cg * 16, ue, w);
                if (p.store_dw) {
                    stg16(p.UE16 + (size_t)tile * T16N_TILE_BYTES + t16_off<8>(row, cg * 2), make_uint4(w[0], w[1], w[2], w[3]));
                    stg16(p.UE16 + (size_t)tile * T16N_TILE_BYTES + t16_off<8>(row, cg * 2 + 1), make_uint4(w[4], w[5], w[6], w[7]));
                }
            }
            sw_publish(&bar);
            // ---- tangent sweep: q_l = W_l u_{l-1}; u_l = s'(h_l) q_l; X_l = 100 (1 - s') D_l q_l ---------------------------------
            for (int l = 0; l < 8; ++l) {
                if (threadIdx.x == 64) dbg_mark(p.dbg, 2, (uint32_t)(t << 8 | l));
                sw_wait_acc(&bar, acc_par);
#pragma unroll 1
                for (int hf = 0; hf < 2; ++hf) {
                    sw_in_wait(&bar, hf, in_par);
#pragma unroll
                    for (int sub = 0; sub < 2; ++sub) {
                        const int c = 128 * hf + cg * 32 + 16 * sub;
                        float q[16], em[16], d[16];
                        unpack_f16(sw_in_ld(smem, hf, 0, row, c), sw_in_ld(smem, hf, 0, row, c + 8), em);
                        unpack_bf16(sw_in_ld(smem, hf, 1, row, c), sw_in_ld(smem, hf, 1, row, c + 8), d);
                        sw_ld16(tmem, lane_base, c, q);
#pragma unroll
                        for (int j = 0; j < 16; ++j) {
                            const float u = (1.0f - em[j]) * q[j];
                            d[j] = 100.0f * em[j] * d[j] * q[j];
                            q[j] = u;
                        }
                        uint32_t w[8];
                        if (l < 7) emit_a16(c, q, w); else pack16(q, w);
                        if (p.store_dw) store16(p.U16[l], c, w);
                        pack16(d, w);
                        store16(p.X16[l], c, w);
                    }
                    sw_in_release(&bar, hf);
                }
                if (l == 2) {
                    // The skip layer's input is [u_3 (193) | ue (63)]: pre-seed the accumulator of layer 3 with [0 | ue] (this
                    // thread's own columns, all read above) and let its MMAs accumulate: the weight rows 193..255 are zero
                    // padding and EM_3 is 0 there (s' = 1, X = 0), so the ordinary epilogue passes ue through untouched.
#pragma unroll 1
                    for (int sb = 0; sb < 4; ++sb) {
                        const int c = 128 * (sb >> 1) + cg * 32 + 16 * (sb & 1);
                        uint32_t v[16];
#pragma unroll
                        for (int j = 0; j < 16; ++j)
                            v[j] = __float_as_uint(c + j > 192 ? enc_tangent_col(e_pt, dn[0], dn[1], dn[2], c + j - 193) : 0.0f);
                        tc::tmem_st_32x32b_x16(tmem + lane_base + (uint32_t)c, v);
                    }
                }
                if (l == 7) {
                    // A operand of the output layer's reverse step: the point's row of d_feat
#pragma unroll 1
                    for (int sb = 0; sb < 4; ++sb) {
                        const int c = 128 * (sb >> 1) + cg * 32 + 16 * (sb & 1);
                        float v[16];
#pragma unroll
                        for (int j = 0; j < 16; j += 4) {
                            float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
                            if (live && p.d_feat) a = ld4(p.d_feat + gp * p.ld_dfeat + c + j);
                            v[j] = a.x; v[j + 1] = a.y; v[j + 2] = a.z; v[j + 3] = a.w;
                        }
                        uint32_t w[8];
                        emit_a16(c, v, w);
                        if (p.store_dw) store16(p.DF16, c, w);
                    }
                }
                sw_publish(&bar);
            }
            // ---- reverse sweep: dz_{l-1} = s'(h_{l-1}) (dz_l W_l) + X_{l-1}, l = 8..1 -------------------------------------------
            for (int l = 8; l >= 1; --l) {
                if (threadIdx.x == 64) dbg_mark(p.dbg, 2, (uint32_t)(t << 8 | (8 + (8 - l))));
                // X_{l-1} was written by THIS thread during the tangent sweep: plain loads, issued before the wait for the MMAs
                const uint8_t* xp = p.X16[l - 1] + tb;
                uint4 xq[8];
#pragma unroll
                for (int i = 0; i < 8; ++i) xq[i] = ldg16(xp + t16_off(row, ((128 * (i >> 2) + cg * 32) >> 3) + (i & 3)));
                sw_wait_acc(&bar, acc_par);
#pragma unroll
                for (int hf = 0; hf < 2; ++hf) {
                    sw_in_wait(&bar, hf, in_par);
#pragma unroll
                    for (int sub = 0; sub < 2; ++sub) {
                        const int c = 128 * hf + cg * 32 + 16 * sub;
                        float da[16], em[16], x[16];
                        unpack_f16(sw_in_ld(smem, hf, 0, row, c), sw_in_ld(smem, hf, 0, row, c + 8), em);
                        unpack_bf16(xq[4 * hf + 2 * sub], xq[4 * hf + 2 * sub + 1], x);
                        sw_ld16(tmem, lane_base, c, da);
                        if (l == 8) {
#pragma unroll
                            for (int j = 0; j < 16; j += 4) {
                                const float4 w = __ldg(reinterpret_cast<const float4*>(p.w_out0 + c + j));
                                da[j] += gs * w.x; da[j + 1] += gs * w.y; da[j + 2] += gs * w.z; da[j + 3] += gs * w.w;
                            }
                        }
                        if (l == 4 && hf == 1 && c + 15 > 192) {       // cotangent of the encoding part of the skip input
#pragma unroll
                            for (int j = 0; j < 16; ++j)
                                if (c + j > 192) {
                                    p.DE[eoff(gp, c + j - 193)] = da[j];
                                    da[j] = 0.0f;
                                    x[j] = 0.0f;
                                }
                        }
#pragma unroll
                        for (int j = 0; j < 16; ++j) da[j] = fmaf(1.0f - em[j], da[j], x[j]);
                        uint32_t w[8];
                        emit_a16(c, da, w);
                        if (p.store_dw) store16(p.DZ16[l - 1], c, w);
                    }
                    sw_in_release(&bar, hf);
                }
                sw_publish(&bar);
            }
            // ---- encoding layer: de = dz_0 W_0 + (skip part); d_x = J_e^T de + Hessian term -----------------------------------------
            sw_wait_acc(&bar, acc_par);
            tc::named_bar_sync(1, EPI_THREADS);          // the skip part was written by other threads of this CTA
            if (p.d_pts) {
                float g[16];
                sw_ld16(tmem, lane_base, cg * 16, g);
                float* __restrict__ de = p.DE + eoff(gp, cg * 16);
#pragma unroll
                for (int j = 0; j < 16; ++j) de[j * TILE_M] = (cg * 16 + j == 63) ? 0.0f : g[j] + de[j * TILE_M];
            }
            tc::tc_fence_before_sync();
            tc::named_bar_sync(1, EPI_THREADS);
            if (p.d_pts && cg < 3 && live) {
                const float* __restrict__ e = p.E + eoff(gp, 3 + cg * 20);
                const float* __restrict__ b = p.EB + eoff(gp, 3 + cg * 20);
                const float* __restrict__ g = p.DE + eoff(gp, 3 + cg * 20);
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
        if (threadIdx.x == 64) dbg_mark(p.dbg, 2, 0xffffffffu);
    }
    sw_teardown(&bar);
}

// ------------------------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------------------------
uint32_t* g_m16_dbg = nullptr;
static int g_dbg_count[4] = {0, 0, 0, 0};
uint32_t* dbg_slot(int kernel_id) {
    if (!g_m16_dbg) return nullptr;
    const int inst = g_dbg_count[kernel_id]++ & 3;
    return g_m16_dbg + ((size_t)kernel_id * 4 + inst) * 148 * 8;
}

struct M16Stash {
    float* E;
    float* EB;
    uint8_t* E16;
    uint8_t* EM[8];
    uint8_t* EML[8];
    uint8_t* A16[8];
    uint8_t* D16[8];
};
static M16Stash m16_stash(float* stash, int64_t np) {
    M16Stash s;
    s.E = stash;
    s.EB = stash + np * 64;
    uint8_t* b = reinterpret_cast<uint8_t*>(stash + np * 128);
    s.E16 = b; b += np * 128;
    for (int l = 0; l < 8; ++l) { s.EM[l] = b; b += np * 512; }
    for (int l = 0; l < 8; ++l) { s.EML[l] = b; b += np * 512; }
    for (int l = 0; l < 8; ++l) { s.A16[l] = b; b += np * 512; }
    for (int l = 0; l < 8; ++l) { s.D16[l] = b; b += np * 512; }
    return s;
}
int64_t m16_stash_floats(int64_t n) { return round_up(n, TILE_M) * (128 + 32 + 32 * 128); }
// backward workspace: U16[8] | X16[8] | DZ16[8] | DF16 | UE16 | DE (fp32) | partial sums of the dW kernel
int64_t m16_bwd_ws_floats(int64_t n) { return round_up(n, TILE_M) * (25 * 128 + 32 + 64) + dw_part_floats(9); }

static int check_m16(const hn_mlp_t* m) {
    HN_REQUIRE(m && m->n_layers == 9, "object SDF mlp must have 9 layers");
    HN_REQUIRE(m->chain && m->chain_bytes >= (int64_t)obj_layout().total && aligned16(m->chain),
               "HN_TC_MIXED16 needs the packed chain operands (hn_sdf_obj_chain_pack)");
    return HN_OK;
}

static void sw_layer(SwProgram& prog, int& k, uint32_t off, int n_mma, int kblocks, int f16, bool acc_in = false) {
    SwStep& st = prog.step[k++];
    st.b_off = off;
    st.n_mma = (uint16_t)n_mma;
    st.kblocks = (uint8_t)kblocks;
    st.f16 = (uint8_t)f16;
    st.acc_in = acc_in ? 1 : 0;
}

int launch_m16_fwd(const hn_mlp_t* m, const float* pts, int64_t n, float inv_scale, float* sdf, float* feat, int64_t ld_feat,
                   float* normal, float* stash, cudaStream_t s) {
    HN_PROPAGATE(check_m16(m));
    const ObjLayout L = obj_layout();
    const int64_t np = round_up(n, TILE_M);
    const M16Stash S = m16_stash(stash, np);
    const int n_tiles = (int)(np / TILE_M);
    const int grid = std::min(n_tiles, sm_count());
    static bool configured = false;
    if (!configured) {
        HN_CHECK_CUDA(cudaFuncSetAttribute(trunk16_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, TR_SMEM_BYTES));
        HN_CHECK_CUDA(cudaFuncSetAttribute(nsweep16_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SW_SMEM_BYTES));
        configured = true;
    }
    {
        Trunk16Params p;
        p.pts = pts; p.n = n; p.inv_scale = inv_scale; p.sdf = sdf; p.feat = feat; p.ld_feat = ld_feat;
        p.E = S.E; p.E16 = S.E16;
        for (int l = 0; l < 8; ++l) { p.EM[l] = S.EM[l]; p.EML[l] = S.EML[l]; p.A16[l] = S.A16[l]; }
        p.chain = reinterpret_cast<const uint8_t*>(m->chain);
        for (int l = 0; l < 9; ++l) p.bias[l] = m->b[l];
        p.w_out0 = m->W[8];
        p.n_tiles = n_tiles;
        p.dbg = dbg_slot(0);
        Program prog = {};
        prog.n_steps = 18;
        for (int l = 0; l < 9; ++l)
            for (int h = 0; h < 2; ++h) {
                Step& st = prog.step[2 * l + h];
                st.b_off = l < 8 ? L.nth_off[l][h] : L.nth8_off[h];
                st.n_mma = 128;
                st.kblocks = L.nt_kb[l];
                st.f16 = 1;
                st.no_wait = (uint8_t)h;
                st.acc_col = (uint16_t)(128 * h);
            }
        TimingScope ts(s, TT_SDF_FWD);
        trunk16_kernel<<<grid, THREADS, TR_SMEM_BYTES, s>>>(p, prog);
    }
    count_launch();
    HN_CHECK_LAUNCH();
    {
        Nsweep16Params p;
        p.n = n; p.inv_scale = inv_scale; p.E = S.E; p.EB = S.EB; p.normal = normal;
        for (int l = 0; l < 8; ++l) p.D16[l] = S.D16[l];
        SwInputs in = {};
        int ne = 0;
        for (int hf = 0; hf < 2; ++hf) in.ev[ne++] = SwInEvent{S.EM[7], S.EML[7]};                  // seed
        for (int l = 7; l >= 1; --l)
            for (int hf = 0; hf < 2; ++hf) in.ev[ne++] = SwInEvent{S.EM[l - 1], S.EML[l - 1]};
        in.n_events = ne;
        p.chain = reinterpret_cast<const uint8_t*>(m->chain);
        p.w_out0 = m->W[8];
        p.n_tiles = n_tiles;
        p.dbg = dbg_slot(1);
        SwProgram prog = {};
        int k = 0;
        for (int l = 7; l >= 0; --l) sw_layer(prog, k, L.nn16_off[l], L.nn_n[l], L.nn_kb[l], 1);         // d @ W_l
        prog.n_steps = k;
        TimingScope ts(s, TT_SDF_FWD);
        nsweep16_kernel<<<grid, SW_THREADS, SW_SMEM_BYTES, s>>>(p, prog, in);
    }
    count_launch();
    HN_CHECK_LAUNCH();
    return HN_OK;
}

int launch_m16_bwd(const hn_mlp_t* m, int64_t n, float inv_scale, const float* stash, const float* d_sdf, const float* d_feat,
                   int64_t ld_dfeat, const float* d_normal, float* d_pts, const hn_mlp_grad_t* grad, float* ws, cudaStream_t s) {
    HN_PROPAGATE(check_m16(m));
    HN_REQUIRE(!d_feat || (ld_dfeat % 4 == 0 && aligned16(d_feat)), "d_feat must be 16-byte aligned with ld %% 4 == 0");
    const ObjLayout L = obj_layout();
    const int64_t np = round_up(n, TILE_M);
    const M16Stash S = m16_stash(const_cast<float*>(stash), np);
    const int n_tiles = (int)(np / TILE_M);
    uint8_t* b = reinterpret_cast<uint8_t*>(ws);
    uint8_t *U16[8], *X16[8], *DZ16[8];
    for (int l = 0; l < 8; ++l) { U16[l] = b; b += np * 512; }
    for (int l = 0; l < 8; ++l) { X16[l] = b; b += np * 512; }
    for (int l = 0; l < 8; ++l) { DZ16[l] = b; b += np * 512; }
    uint8_t* DF16 = b; b += np * 512;
    uint8_t* UE16 = b; b += np * 128;
    float* DE = reinterpret_cast<float*>(b); b += np * 256;
    float* part = reinterpret_cast<float*>(b);
    static bool configured = false;
    if (!configured) {
        HN_CHECK_CUDA(cudaFuncSetAttribute(bwd16_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SW_SMEM_BYTES));
        configured = true;
    }
    {
        Bwd16Params p;
        p.n = n; p.inv_scale = inv_scale; p.E = S.E; p.EB = S.EB;
        for (int l = 0; l < 8; ++l) { p.U16[l] = U16[l]; p.X16[l] = X16[l]; p.DZ16[l] = DZ16[l]; }
        SwInputs in = {};
        int ne = 0;
        for (int l = 0; l < 8; ++l)
            for (int hf = 0; hf < 2; ++hf) in.ev[ne++] = SwInEvent{S.EM[l], S.D16[l]};              // tangent sweep
        for (int l = 8; l >= 1; --l)
            for (int hf = 0; hf < 2; ++hf) in.ev[ne++] = SwInEvent{S.EM[l - 1], nullptr};           // reverse sweep
        in.n_events = ne;
        p.d_sdf = d_sdf; p.d_feat = d_feat; p.ld_dfeat = ld_dfeat; p.d_normal = d_normal; p.d_pts = d_pts;
        p.UE16 = UE16; p.DF16 = DF16; p.DE = DE;
        p.store_dw = grad ? 1 : 0;
        p.chain = reinterpret_cast<const uint8_t*>(m->chain);
        p.w_out0 = m->W[8];
        p.n_tiles = n_tiles;
        p.dbg = dbg_slot(2);
        SwProgram prog = {};
        int k = 0;
        for (int l = 0; l < 8; ++l) sw_layer(prog, k, L.nt_off[l], L.nt_n[l], L.nt_kb[l], 0, l == 3);     // tangent: u @ W_l^T
        for (int l = 8; l >= 0; --l) sw_layer(prog, k, L.nn_off[l], L.nn_n[l], L.nn_kb[l], 0);            // reverse: dz @ W_l
        prog.n_steps = k;
        TimingScope ts(s, TT_SDF_BWD);
        bwd16_kernel<<<std::min(n_tiles, sm_count()), SW_THREADS, SW_SMEM_BYTES, s>>>(p, prog, in);
    }
    count_launch();
    HN_CHECK_LAUNCH();
    if (!grad) return HN_OK;
    // ---- all weight / bias gradients from the tiles the two sweep kernels left in HBM ------------------------------------
    Dw16Params dp = {};
    DwReduceParams rp;
    dp.n_tiles = n_tiles; dp.part = part; rp.part = part;
    int k = 0;
    for (int l = 0; l < 8; ++l, ++k) {
        Dw16Job& j = dp.job[k];
        const int out = m->out_dim[l], in = m->in_dim[l];
        j.P[0] = DZ16[l]; j.P[1] = S.D16[l];
        j.Q[0] = l == 0 ? S.E16 : S.A16[l - 1];          // A16[3] holds the skip input [h3 | e]
        j.Q[1] = l == 0 ? UE16 : U16[l - 1];             // U16[3] holds [u3 | ue]
        j.n_pairs = 2;
        j.q_chunks = l == 0 ? 8 : 32;
        j.n_mma = l == 0 ? 64 : 256;
        j.db = grad->db[l]; j.p_cols = out;
        rp.job[k] = reduce_job(grad->dW[l], m->ld[l], 0, out, in);
    }
    if (d_feat) {
        Dw16Job& j = dp.job[k];
        j.P[0] = DF16; j.Q[0] = S.A16[7]; j.P[1] = DF16; j.Q[1] = S.A16[7];
        j.n_pairs = 1; j.q_chunks = 32; j.n_mma = 256;
        j.db = grad->db[8] ? grad->db[8] + 1 : nullptr; j.p_cols = 256;
        rp.job[k] = reduce_job(grad->dW[8], m->ld[8], 1, 256, 256);
        ++k;
    }
    dp.n_jobs = k;
    HN_PROPAGATE(launch_dw16(dp, rp, s));
    if (grad->dW[8] || grad->db[8])
        HN_PROPAGATE(launch_out_row0_grad16(S.A16[7], U16[7], d_sdf, n, n_tiles, inv_scale, grad->dW[8], grad->db[8], s));
    return HN_OK;
}

}  // namespace chain
}  // namespace hn

extern "C" int hn_chain16_set_debug(void* host_mapped_words) {
    hn::chain::g_m16_dbg = reinterpret_cast<uint32_t*>(host_mapped_words);
    return HN_OK;
}
