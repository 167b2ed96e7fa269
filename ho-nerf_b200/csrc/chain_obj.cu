// Object SDF field as fused tile-chain kernels on tcgen05 (HN_TC_BF16X3), see chain_common.cuh.
// Replaces, per 128-point tile and without leaving the SM between layers,
//   SDFNetwork_OBJ.sdf       (utils/fields.py:316-331)          -> sdf_only_kernel
// The positional encoding (utils/fields.py:13-20) is computed by the epilogue warps straight into the
// first layer's A operand; the skip connection's [h3 | e] / sqrt2 (utils/fields.py:323-324) is formed in
// place (1/sqrt2 is folded into W_4 at pack time, SURVEY A-8).
#include <algorithm>

#include "chain_common.cuh"
#include "fields_common.cuh"

namespace hn {
namespace chain {

// ------------------------------------------------------------------------------------------------
// packing
// ------------------------------------------------------------------------------------------------
__global__ void pack_b_kernel(const float* __restrict__ src, int64_t ld, int row0, int col0, int rows, int cols,
                              int n_pad, int kblocks, uint8_t* __restrict__ dst) {
    int idx = blockIdx.x * blockDim.x + threadIdx.x;          // one 16-byte chunk (8 columns) of one row
    int total = n_pad * kblocks * 8;
    if (idx >= total) return;
    int n = idx / (kblocks * 8);
    int c = idx - n * (kblocks * 8);
    int kb = c >> 3, c16 = c & 7;
    float v[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        int k = kb * 64 + c16 * 8 + j;
        v[j] = (n < rows && k < cols) ? src[(int64_t)(row0 + n) * ld + col0 + k] : 0.0f;
    }
    uint4 hi, lo;
    split2(v[0], v[1], hi.x, lo.x);
    split2(v[2], v[3], hi.y, lo.y);
    split2(v[4], v[5], hi.z, lo.z);
    split2(v[6], v[7], hi.w, lo.w);
    size_t base = (size_t)kb * 2 * n_pad * 128 + tc::sw128_offset((uint32_t)n, (uint32_t)c16);
    *reinterpret_cast<uint4*>(dst + base) = hi;
    *reinterpret_cast<uint4*>(dst + base + (size_t)n_pad * 128) = lo;
}

int launch_pack_b(const float* src, int64_t ld, int row0, int col0, int rows, int cols, int n_pad, int kblocks,
                  uint8_t* dst, cudaStream_t stream) {
    int total = n_pad * kblocks * 8;
    pack_b_kernel<<<(unsigned)ceil_div(total, 256), 256, 0, stream>>>(src, ld, row0, col0, rows, cols, n_pad, kblocks, dst);
    count_launch();
    HN_CHECK_LAUNCH();
    return HN_OK;
}

// Packed operands of the object SDF net.  NT[l]: B(n = output feature, k = input feature) for
// a @ W_l^T (value trunk, tangent sweep); NN[l]: B(n = input feature, k = output feature) for
// d @ W_l (normal sweep, reverse sweep).  The output layer is packed without its sdf row (row 0),
// which is applied as a rank-one term by the epilogues.
struct ObjLayout {
    uint32_t nt_off[9], nn_off[9];
    uint16_t nt_n[9], nn_n[9];
    uint8_t nt_kb[9], nn_kb[9];
    uint32_t total;
};
static ObjLayout obj_layout() {
    ObjLayout L;
    uint32_t off = 0;
    for (int l = 0; l < 9; ++l) {
        L.nt_n[l] = l == 3 ? 208 : 256;
        L.nt_kb[l] = l == 0 ? 1 : 4;
        L.nt_off[l] = off;
        off += b_operand_bytes(L.nt_n[l], L.nt_kb[l]);
    }
    for (int l = 0; l < 9; ++l) {
        L.nn_n[l] = l == 0 ? 64 : 256;
        L.nn_kb[l] = 4;
        L.nn_off[l] = off;
        off += b_operand_bytes(L.nn_n[l], L.nn_kb[l]);
    }
    L.total = off;
    return L;
}

// ------------------------------------------------------------------------------------------------
// epilogue helpers
// ------------------------------------------------------------------------------------------------
// [x(3), sin/cos(2^k x_c)] of one point written as columns shift + j of the A operand; the column
// groups of a row share the 30 (coordinate, frequency) pairs.  Column shift+63 (the K padding of the
// first layer) is zeroed when shift == 0.
__device__ __forceinline__ void write_encoding(uint8_t* smem, int row, int cg, const float x[3], int shift) {
    if (cg == 0) {
#pragma unroll
        for (int c = 0; c < 3; ++c) a_store1(smem, row, shift + c, x[c]);
        if (shift == 0) a_store1(smem, row, 63, 0.0f);
    }
    for (int idx = cg; idx < 30; idx += EPI_CGROUPS) {
        const int c = idx / 10, k = idx - c * 10;
        float s, co;
        sincosf(x[c] * (float)(1 << k), &s, &co);
        a_store1(smem, row, shift + 3 + c * 20 + k, s);
        a_store1(smem, row, shift + 3 + c * 20 + 10 + k, co);
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
            write_encoding(smem, row, cg, x, 0);
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
                            for (int j = 0; j < 32; j += 8) a_store8(smem, row, col0 + j, v + j);
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
                        a_store1(smem, row, 192, softplus100_fast(v[0] + __ldg(bias + 192)));
                    }
                    write_encoding(smem, row, cg, x, 193);
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

static int check_chain_mlp(const hn_mlp_t* m) {
    HN_REQUIRE(m && m->n_layers == 9, "object SDF mlp must have 9 layers");
    HN_REQUIRE(m->chain && m->chain_bytes >= (int64_t)obj_layout().total && aligned16(m->chain),
               "HN_TC_BF16X3 needs the packed chain operands (hn_sdf_obj_chain_pack)");
    return HN_OK;
}

static long long* g_prof = nullptr;   // set by hn_chain_set_prof (diagnostics)

int launch_sdf_only(const hn_mlp_t* m, const float* pts, int64_t n, float inv_scale, float* sdf, cudaStream_t s) {
    HN_PROPAGATE(check_chain_mlp(m));
    const ObjLayout L = obj_layout();
    SdfOnlyParams p;
    p.pts = pts; p.n = n; p.inv_scale = inv_scale; p.sdf = sdf;
    p.chain = reinterpret_cast<const uint8_t*>(m->chain);
    for (int l = 0; l < 9; ++l) p.bias[l] = m->b[l];
    p.w_out0 = m->W[8];
    p.n_tiles = (int)ceil_div(n, TILE_M);
    p.prof = g_prof;
    Program prog;
    prog.n_steps = 8;
    for (int l = 0; l < 8; ++l) {
        prog.step[l].b_off = L.nt_off[l];
        prog.step[l].n_mma = L.nt_n[l];
        prog.step[l].kblocks = L.nt_kb[l];
        prog.step[l].a_kb0 = 0;
    }
    static bool configured = false;
    if (!configured) {
        HN_CHECK_CUDA(cudaFuncSetAttribute(sdf_only_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
        configured = true;
    }
    const int grid = std::min(p.n_tiles, sm_count());
    {
        TimingScope ts(s);
        sdf_only_kernel<<<grid, THREADS, SMEM_BYTES, s>>>(p, prog);
    }
    count_launch();
    HN_CHECK_LAUNCH();
    return HN_OK;
}

}  // namespace chain
}  // namespace hn

using namespace hn;

extern "C" {

int hn_chain_set_prof(void* buf) {
    chain::g_prof = reinterpret_cast<long long*>(buf);
    return HN_OK;
}

int64_t hn_sdf_obj_chain_bytes(void) { return (int64_t)chain::obj_layout().total; }

int hn_sdf_obj_chain_pack(const hn_mlp_t* m, void* chain_buf, int64_t chain_bytes, hn_stream_t stream) {
    HN_REQUIRE(m && m->n_layers == 9, "hn_sdf_obj_chain_pack: object SDF mlp must have 9 layers");
    const chain::ObjLayout L = chain::obj_layout();
    HN_REQUIRE(chain_buf && chain_bytes >= (int64_t)L.total && aligned16(chain_buf),
               "hn_sdf_obj_chain_pack: buffer too small or misaligned (need %u bytes)", L.total);
    static const int in_d[9] = {63, 256, 256, 256, 256, 256, 256, 256, 256};
    static const int out_d[9] = {256, 256, 256, 193, 256, 256, 256, 256, 257};
    cudaStream_t s = (cudaStream_t)stream;
    uint8_t* dst = reinterpret_cast<uint8_t*>(chain_buf);
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
    }
    return HN_OK;
}

}  // extern "C"
