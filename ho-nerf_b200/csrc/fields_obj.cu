// Object SDF field (value + feature + analytic normal, second-order backward) and object colour
// field, HN_SIMT_FP32 path: layer-by-layer fp32 GEMMs with fused epilogues, activations in HBM.
// Replaces utils/fields.py:316-347 (SDFNetwork_OBJ.forward/.sdf/.gradient) and :387-405
// (RenderingNetwork_OBJ.forward) plus everything autograd derives from them.
#include <algorithm>

#include "common.cuh"
#include "gemm_dispatch.cuh"
#include "fields_common.cuh"

namespace hn {

namespace chain {   // chain_obj.cu
int launch_sdf_only(const hn_mlp_t* m, const float* pts, int64_t n, float inv_scale, float* sdf, cudaStream_t s);
int launch_sdf_fwd(const hn_mlp_t* m, const float* pts, int64_t n, float inv_scale, float* sdf, float* feat,
                   int64_t ld_feat, float* normal, float* stash, cudaStream_t s);
int64_t color_stash_floats(int64_t n);
int64_t color_bwd_ws_floats(int64_t n);
int launch_color_fwd(const hn_mlp_t* m, const float* pts, const float* dirs, const float* feat, int64_t ld_feat,
                     const float* normal, int64_t n, float* rgb, float* stash, cudaStream_t s, bool s16);
int launch_color_bwd(const hn_mlp_t* m, int64_t n, const float* stash, const float* rgb, const float* d_rgb, float* d_pts,
                     float* d_dirs, float* d_feat, int64_t ld_dfeat, float* d_normal, const hn_mlp_grad_t* grad, float* ws,
                     cudaStream_t s, bool s16);
int64_t bwd_ws_floats(int64_t n);
int64_t stash_floats(int64_t n);
int launch_sdf_bwd_weights(const hn_mlp_t* m, int64_t n, float inv_scale, const float* stash, const float* d_sdf,
                           const float* d_feat, int64_t ld_dfeat, float* ws, const hn_mlp_grad_t* grad,
                           cudaStream_t s);
int launch_sdf_bwd(const hn_mlp_t* m, int64_t n, float inv_scale, const float* stash, const float* d_sdf,
                   const float* d_feat, int64_t ld_dfeat, const float* d_normal, float* d_pts, float* ws,
                   cudaStream_t s);
// chain16_obj.cu (HN_TC_MIXED16)
int64_t m16_stash_floats(int64_t n);
int64_t m16_bwd_ws_floats(int64_t n);
int launch_m16_fwd(const hn_mlp_t* m, const float* pts, int64_t n, float inv_scale, float* sdf, float* feat, int64_t ld_feat,
                   float* normal, float* stash, cudaStream_t s);
int launch_m16_bwd(const hn_mlp_t* m, int64_t n, float inv_scale, const float* stash, const float* d_sdf, const float* d_feat,
                   int64_t ld_dfeat, const float* d_normal, float* d_pts, const hn_mlp_grad_t* grad, float* ws, cudaStream_t s);
}

// ------------------------------------------------------------------------------------------
// elementwise kernels of the object SDF net
// ------------------------------------------------------------------------------------------
// E[p, 0:64] = [x(3), enc10(x) (60), 0];  A4[p, 193:256] = E[p, 0:63]   (A4 may be NULL)
__global__ void enc_obj_kernel(const float* __restrict__ pts, int64_t n, float* __restrict__ E,
                               float* __restrict__ A4) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n * 64) return;
    int64_t p = i >> 6;
    int j = (int)(i & 63);
    float v = 0.0f;
    if (j < 63) {
        const float x[3] = {pts[p * 3 + 0], pts[p * 3 + 1], pts[p * 3 + 2]};
        v = enc3_col(x, 10, j);
        if (A4) A4[p * 256 + 193 + j] = v;
    }
    E[i] = v;
}

// D7[p, c] = s'(H7[p, c]) * w_out0[c] * inv_scale    (seed of the normal sweep)
__global__ void normal_seed_kernel(const float* __restrict__ H7, const float* __restrict__ w_out0,
                                   float inv_scale, int64_t n, float* __restrict__ D7) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n * 64) return;
    int c = (int)(i & 63) * 4;
    int64_t p = i >> 6;
    float4 h = ld4(H7 + p * 256 + c);
    float4 w = ld4(w_out0 + c);
    st4(D7 + p * 256 + c, make_float4(sprime_from_h(h.x) * w.x * inv_scale, sprime_from_h(h.y) * w.y * inv_scale,
                                      sprime_from_h(h.z) * w.z * inv_scale, sprime_from_h(h.w) * w.w * inv_scale));
}

// sdf[p] = (H7[p,:] . w_out0 + b0) * inv_scale     one warp per point
__global__ void sdf_head_kernel(const float* __restrict__ H7, const float* __restrict__ w_out0,
                                const float* __restrict__ b_out, float inv_scale, int64_t n,
                                float* __restrict__ sdf) {
    int64_t p = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    int lane = threadIdx.x & 31;
    if (p >= n) return;
    float acc = 0.0f;
#pragma unroll
    for (int it = 0; it < 2; ++it) {
        int c = (it * 32 + lane) * 4;
        float4 h = ld4(H7 + p * 256 + c);
        float4 w = ld4(w_out0 + c);
        acc += h.x * w.x + h.y * w.y + h.z * w.z + h.w * w.w;
    }
    acc = warp_sum(acc);
    if (lane == 0) sdf[p] = (acc + b_out[0]) * inv_scale;
}

// normal[p, c] = J_e(x)^T EB[p, :]     (E holds x, sin, cos)
__global__ void normal_from_eb_kernel(const float* __restrict__ E, const float* __restrict__ EB,
                                      int64_t n, float* __restrict__ normal) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n * 3) return;
    int64_t p = i / 3;
    int c = (int)(i - p * 3);
    const float* e = E + p * 64 + 3 + c * 20;
    const float* g = EB + p * 64 + 3 + c * 20;
    float acc = EB[p * 64 + c];
    float f = 1.0f;
#pragma unroll
    for (int k = 0; k < 10; ++k) {
        acc += f * (e[10 + k] * g[k] - e[k] * g[10 + k]);
        f *= 2.0f;
    }
    normal[i] = acc;
}

// UE[p, :] = J_e(x) dn[p, :];  AU4[p, 193:256] = UE[p, 0:63]
__global__ void enc_tangent_kernel(const float* __restrict__ E, const float* __restrict__ dn,
                                   int64_t n, float* __restrict__ UE, float* __restrict__ AU4) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n * 64) return;
    int64_t p = i >> 6;
    int j = (int)(i & 63);
    float v = 0.0f;
    if (j < 3) {
        v = dn[p * 3 + j];
    } else if (j < 63) {
        int jj = j - 3;
        int c = jj / 20, r = jj - c * 20;
        int s = r / 10, k = r - s * 10;
        float f = (float)(1 << k);
        float t = dn[p * 3 + c];
        // d sin = f cos t ; d cos = -f sin t
        v = s == 0 ? f * E[p * 64 + 3 + c * 20 + 10 + k] * t : -f * E[p * 64 + 3 + c * 20 + k] * t;
    }
    UE[i] = v;
    if (j < 63) AU4[p * 256 + 193 + j] = v;
}

// DZ8[p, 0] = d_sdf[p]*inv_scale ; DZ8[p, 1+j] = d_feat[p, j] ; ld 260
__global__ void assemble_dz8_kernel(const float* __restrict__ d_sdf, const float* __restrict__ d_feat,
                                    int64_t ld_dfeat, float inv_scale, int64_t n,
                                    float* __restrict__ DZ8) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n * 260) return;
    int64_t p = i / 260;
    int j = (int)(i - p * 260);
    float v = 0.0f;
    if (j == 0) v = d_sdf ? d_sdf[p] * inv_scale : 0.0f;
    else if (j < 257) v = d_feat ? d_feat[p * ld_dfeat + (j - 1)] : 0.0f;
    DZ8[i] = v;
}

// d_pts[p,c] = J_e^T DE + dn_c * sum_k 4^k (-sin*EB[sin] - cos*EB[cos])
__global__ void dx_obj_kernel(const float* __restrict__ E, const float* __restrict__ DE,
                              const float* __restrict__ EB, const float* __restrict__ dn, int64_t n,
                              float* __restrict__ d_pts) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n * 3) return;
    int64_t p = i / 3;
    int c = (int)(i - p * 3);
    const float* e = E + p * 64 + 3 + c * 20;
    const float* g = DE + p * 64 + 3 + c * 20;
    const float* b = EB + p * 64 + 3 + c * 20;
    float acc = DE[p * 64 + c];
    float hess = 0.0f;
    float f = 1.0f;
#pragma unroll
    for (int k = 0; k < 10; ++k) {
        acc += f * (e[10 + k] * g[k] - e[k] * g[10 + k]);
        hess -= f * f * (e[k] * b[k] + e[10 + k] * b[10 + k]);
        f *= 2.0f;
    }
    d_pts[i] = acc + dn[i] * hess;
}

// ------------------------------------------------------------------------------------------
// stash / workspace layouts (floats per point)
// ------------------------------------------------------------------------------------------
struct ObjSdfStash {
    float* E;      // [n,64]
    float* H[8];   // [n,256]; H[3] is A4 = [h3 (193) | e (63)]
    float* D[8];   // [n,256]; normal-sweep cotangents s'(z_l)*hb_l, overwritten by X_l in backward
    float* EB;     // [n,64]  cotangent of the encoding in the normal sweep
    static constexpr int64_t kFloatsPerPoint = 64 + 8 * 256 + 8 * 256 + 64;
    ObjSdfStash(float* base, int64_t n) {
        float* p = base;
        E = p; p += n * 64;
        for (int l = 0; l < 8; ++l) { H[l] = p; p += n * 256; }
        for (int l = 0; l < 8; ++l) { D[l] = p; p += n * 256; }
        EB = p;
    }
};

static int check_obj_sdf_mlp(const hn_mlp_t* m) {
    HN_REQUIRE(m && m->n_layers == 9, "object SDF mlp must have 9 layers");
    static const int in_d[9] = {63, 256, 256, 256, 256, 256, 256, 256, 256};
    static const int out_d[9] = {256, 256, 256, 193, 256, 256, 256, 256, 257};
    for (int l = 0; l < 9; ++l) {
        HN_REQUIRE(m->in_dim[l] == in_d[l] && m->out_dim[l] == out_d[l],
                   "object SDF mlp layer %d is %dx%d, expected %dx%d", l, m->out_dim[l],
                   m->in_dim[l], out_d[l], in_d[l]);
        HN_REQUIRE(m->ld[l] >= round_up(in_d[l], 4) && m->ld[l] % 4 == 0, "bad ld for layer %d", l);
        HN_REQUIRE(m->W[l] && m->b[l] && aligned16(m->W[l]), "layer %d: null or misaligned weights", l);
    }
    return HN_OK;
}

static inline unsigned blocks_for(int64_t work, int threads) { return (unsigned)ceil_div(work, threads); }

// forward trunk: E -> H7 (writes every H when `stash_all`, otherwise ping-pongs between 2 buffers)
static int obj_trunk_fwd(const hn_mlp_t* m, const float* pts, int64_t n, float* E, float* const H[8],
                         cudaStream_t s, int precision) {
    enc_obj_kernel<<<blocks_for(n * 64, 256), 256, 0, s>>>(pts, n, E, H[3]);
    count_launch();
    HN_CHECK_LAUNCH();
    for (int l = 0; l < 8; ++l) {
        GemmArgs g;
        g.A = l == 0 ? E : H[l - 1];
        g.lda = l == 0 ? 64 : 256;
        g.B = m->W[l]; g.ldb = m->ld[l];
        g.M = (int)n; g.N = m->out_dim[l]; g.K = m->in_dim[l];
        g.C = H[l]; g.ldc = 256;
        g.bias = m->b[l];
        HN_PROPAGATE((gemm_nt<EPI_BIAS_SOFTPLUS>(g, s, precision, ROLE_VALUE)));
    }
    return HN_OK;
}

}  // namespace hn

using namespace hn;

extern "C" {

// sized for the tiled layout of the HN_TC_BF16X3 path (n padded to whole 128-point tiles)
int64_t hn_sdf_obj_stash_floats(int64_t n) { return std::max<int64_t>(std::max<int64_t>(n * ObjSdfStash::kFloatsPerPoint, chain::stash_floats(n)), chain::m16_stash_floats(n)); }

int64_t hn_sdf_obj_ws_floats(int64_t n, int kind) {
    switch (kind) {
        case HN_WS_SDF_ONLY: return n * (64 + 3 * 256);       // E, two ping-pong H, A4
        case HN_WS_FWD: return 4;                              // nothing beyond the stash
        case HN_WS_BWD: return std::max<int64_t>(std::max<int64_t>(n * (64 + 2 * 256 + 256 + 260 + 2 * 256 + 64), chain::bwd_ws_floats(n)), chain::m16_bwd_ws_floats(n));
        default: return -1;
    }
}

int hn_sdf_obj_sdf(const hn_mlp_t* mlp, const float* pts, int64_t n, float inv_scale, float* sdf,
                   float* ws, int64_t ws_floats, int precision, hn_stream_t stream) {
    HN_PROPAGATE(check_obj_sdf_mlp(mlp));
    HN_REQUIRE(precision_supported(precision), "hn_sdf_obj_sdf: precision %d not supported", precision);
    HN_REQUIRE(n >= 0 && n < (1ll << 31), "n_pts out of range");
    if (n == 0) return HN_OK;
    cudaStream_t s = (cudaStream_t)stream;
    if (precision == HN_TC_BF16X3 || precision == HN_TC_MIXED16) {
        HN_REQUIRE(pts && sdf, "hn_sdf_obj_sdf: null pointer");
        return chain::launch_sdf_only(mlp, pts, n, inv_scale, sdf, s);
    }
    HN_REQUIRE(pts && sdf && ws && ws_floats >= hn_sdf_obj_ws_floats(n, HN_WS_SDF_ONLY) && aligned16(ws),
               "hn_sdf_obj_sdf: workspace too small or misaligned");
    float* E = ws;
    float* P0 = E + n * 64;
    float* P1 = P0 + n * 256;
    float* A4 = P1 + n * 256;
    float* H[8] = {P0, P1, P0, A4, P0, P1, P0, P1};
    HN_PROPAGATE(obj_trunk_fwd(mlp, pts, n, E, H, s, precision));
    sdf_head_kernel<<<blocks_for(n * 32, 256), 256, 0, s>>>(H[7], mlp->W[8], mlp->b[8], inv_scale, n, sdf);
    count_launch();
    HN_CHECK_LAUNCH();
    return HN_OK;
}

int hn_sdf_obj_fwd(const hn_mlp_t* mlp, const float* pts, int64_t n, float inv_scale, float* sdf,
                   float* feat, int64_t ld_feat, float* normal, float* stash, int64_t stash_floats,
                   float* ws, int64_t ws_floats, int precision, hn_stream_t stream) {
    (void)ws; (void)ws_floats;
    HN_PROPAGATE(check_obj_sdf_mlp(mlp));
    HN_REQUIRE(precision_supported(precision), "hn_sdf_obj_fwd: precision %d not supported", precision);
    HN_REQUIRE(n >= 0 && n < (1ll << 31), "n_pts out of range");
    if (n == 0) return HN_OK;
    HN_REQUIRE(pts && sdf && feat && normal && stash, "hn_sdf_obj_fwd: null pointer");
    HN_REQUIRE(stash_floats >= hn_sdf_obj_stash_floats(n) && aligned16(stash), "stash too small or misaligned");
    HN_REQUIRE(ld_feat >= 256 && ld_feat % 4 == 0 && aligned16(feat), "feat must be 16B aligned with ld%%4==0");
    cudaStream_t s = (cudaStream_t)stream;
    if (precision == HN_TC_BF16X3) return chain::launch_sdf_fwd(mlp, pts, n, inv_scale, sdf, feat, ld_feat, normal, stash, s);
    if (precision == HN_TC_MIXED16) return chain::launch_m16_fwd(mlp, pts, n, inv_scale, sdf, feat, ld_feat, normal, stash, s);
    ObjSdfStash st(stash, n);
    HN_PROPAGATE(obj_trunk_fwd(mlp, pts, n, st.E, st.H, s, precision));
    // output layer: column 0 -> sdf, columns 1..256 -> feature
    sdf_head_kernel<<<blocks_for(n * 32, 256), 256, 0, s>>>(st.H[7], mlp->W[8], mlp->b[8], inv_scale, n, sdf);
    count_launch();
    HN_CHECK_LAUNCH();
    {
        GemmArgs g;
        g.A = st.H[7]; g.lda = 256;
        g.B = mlp->W[8] + mlp->ld[8]; g.ldb = mlp->ld[8];
        g.M = (int)n; g.N = 256; g.K = 256;
        g.C = feat; g.ldc = ld_feat; g.bias = mlp->b[8] + 1;
        HN_PROPAGATE((gemm_nt<EPI_STORE>(g, s, precision, ROLE_VALUE)));
    }
    // normal sweep
    normal_seed_kernel<<<blocks_for(n * 64, 256), 256, 0, s>>>(st.H[7], mlp->W[8], inv_scale, n, st.D[7]);
    count_launch();
    HN_CHECK_LAUNCH();
    for (int l = 7; l >= 1; --l) {
        GemmArgs g;
        g.A = st.D[l]; g.lda = 256;
        g.B = mlp->W[l]; g.ldb = mlp->ld[l]; g.BT = mlp->WT[l]; g.ldbt = mlp->ldT[l];
        g.M = (int)n; g.N = mlp->in_dim[l]; g.K = mlp->out_dim[l];
        g.C = st.D[l - 1]; g.ldc = 256;
        g.aux1 = st.H[l - 1]; g.ldaux1 = 256;
        if (l == 4) { g.nsplit = 193; g.C2 = st.EB; g.ldc2 = 64; }
        HN_PROPAGATE((gemm_nn<EPI_MUL_SPRIME>(g, s, precision, ROLE_VALUE)));
    }
    {
        GemmArgs g;
        g.A = st.D[0]; g.lda = 256;
        g.B = mlp->W[0]; g.ldb = mlp->ld[0]; g.BT = mlp->WT[0]; g.ldbt = mlp->ldT[0];
        g.M = (int)n; g.N = 63; g.K = 256;
        g.C = st.EB; g.ldc = 64; g.aux1 = st.EB; g.ldaux1 = 64;
        HN_PROPAGATE((gemm_nn<EPI_ADD_AUX>(g, s, precision, ROLE_VALUE)));
    }
    normal_from_eb_kernel<<<blocks_for(n * 3, 256), 256, 0, s>>>(st.E, st.EB, n, normal);
    count_launch();
    HN_CHECK_LAUNCH();
    return HN_OK;
}

int hn_sdf_obj_bwd(const hn_mlp_t* mlp, int64_t n, float inv_scale, float* stash, const float* d_sdf,
                   const float* d_feat, int64_t ld_dfeat, const float* d_normal, float* d_pts,
                   const hn_mlp_grad_t* grad, float* ws, int64_t ws_floats, int precision,
                   hn_stream_t stream) {
    HN_PROPAGATE(check_obj_sdf_mlp(mlp));
    HN_REQUIRE(precision_supported(precision), "hn_sdf_obj_bwd: precision %d not supported", precision);
    HN_REQUIRE(n >= 0 && n < (1ll << 31), "n_pts out of range");
    if (n == 0) return HN_OK;
    HN_REQUIRE(stash && d_normal && ws, "hn_sdf_obj_bwd: null pointer");
    HN_REQUIRE(ws_floats >= hn_sdf_obj_ws_floats(n, HN_WS_BWD) && aligned16(ws), "workspace too small or misaligned");
    HN_REQUIRE(!d_feat || (ld_dfeat >= 256), "bad ld_dfeat");
    cudaStream_t s = (cudaStream_t)stream;
    ObjSdfStash st(stash, n);
    if (precision == HN_TC_MIXED16)
        return chain::launch_m16_bwd(mlp, n, inv_scale, stash, d_sdf, d_feat, ld_dfeat, d_normal, d_pts, grad, ws, s);
    if (precision == HN_TC_BF16X3) {
        // fused tangent + reverse sweeps, then every weight-gradient contraction in one launch over the operands
        // the sweeps left in ws
        HN_PROPAGATE(chain::launch_sdf_bwd(mlp, n, inv_scale, stash, d_sdf, d_feat, ld_dfeat, d_normal, d_pts, ws, s));
        if (!grad) return HN_OK;
        return chain::launch_sdf_bwd_weights(mlp, n, inv_scale, stash, d_sdf, d_feat, ld_dfeat, ws, grad, s);
    }
    float* UE = ws;
    float* U[2] = {UE + n * 64, UE + n * 64 + n * 256};
    float* AU4 = U[1] + n * 256;
    float* DZ8 = AU4 + n * 256;
    float* DZ[2] = {DZ8 + n * 260, DZ8 + n * 260 + n * 256};
    float* DE = DZ[1] + n * 256;
    const int splits_target = 2 * sm_count();

    auto dw_gemm = [&](const float* P, int64_t ldp, int out, const float* Q, int64_t ldq, int in, int l) -> int {
        if (!grad || !grad->dW[l]) return HN_OK;
        GemmArgs g;
        g.A = P; g.lda = ldp; g.B = Q; g.ldb = ldq;
        g.M = out; g.N = in; g.K = (int)n;
        g.C = grad->dW[l]; g.ldc = mlp->ld[l];
        int tiles = (int)(ceil_div(out, GBM) * ceil_div(in, GBN));
        int splits = (int)max((int64_t)1, min((int64_t)ceil_div(splits_target, tiles), ceil_div(n, 256)));
        return gemm_tn(g, s, precision, splits);
    };
    auto db_sum = [&](const float* X, int64_t ldx, int cols, int l) -> int {
        if (!grad || !grad->db[l]) return HN_OK;
        return launch_colsum(X, ldx, n, cols, 1.0f, grad->db[l], s);
    };

    // ---- tangent sweep along d_normal (forward-mode through the trunk) ------------------------
    enc_tangent_kernel<<<blocks_for(n * 64, 256), 256, 0, s>>>(st.E, d_normal, n, UE, AU4);
    count_launch();
    HN_CHECK_LAUNCH();
    const float* au_prev = UE;
    int64_t ld_prev = 64;
    for (int l = 0; l < 8; ++l) {
        if (l == 4) { au_prev = AU4; ld_prev = 256; }
        HN_PROPAGATE(dw_gemm(st.D[l], 256, mlp->out_dim[l], au_prev, ld_prev, mlp->in_dim[l], l));
        GemmArgs g;
        g.A = au_prev; g.lda = ld_prev;
        g.B = mlp->W[l]; g.ldb = mlp->ld[l]; g.BT = mlp->WT[l]; g.ldbt = mlp->ldT[l];
        g.M = (int)n; g.N = mlp->out_dim[l]; g.K = mlp->in_dim[l];
        float* out = l == 3 ? AU4 : U[l & 1];
        g.C = out; g.ldc = 256;
        g.aux1 = st.H[l]; g.ldaux1 = 256;
        g.C2 = st.D[l]; g.ldc2 = 256;      // D_l -> X_l in place
        HN_PROPAGATE((gemm_nt<EPI_TANGENT>(g, s, precision)));
        au_prev = out; ld_prev = 256;
    }
    // the normal sweep is seeded with row 0 of the output layer: dW8[0,:] += colsum(U7)*inv_scale
    if (grad && grad->dW[8]) HN_PROPAGATE(launch_colsum(au_prev, 256, n, 256, inv_scale, grad->dW[8], s));

    // ---- reverse sweep ------------------------------------------------------------------------
    assemble_dz8_kernel<<<blocks_for(n * 260, 256), 256, 0, s>>>(d_sdf, d_feat, ld_dfeat, inv_scale, n, DZ8);
    count_launch();
    HN_CHECK_LAUNCH();
    const float* dz = DZ8;
    int64_t ld_dz = 260;
    for (int l = 8; l >= 1; --l) {
        HN_PROPAGATE(dw_gemm(dz, ld_dz, mlp->out_dim[l], st.H[l - 1], 256, mlp->in_dim[l], l));
        HN_PROPAGATE(db_sum(dz, ld_dz, mlp->out_dim[l], l));
        GemmArgs g;
        g.A = dz; g.lda = ld_dz;
        g.B = mlp->W[l]; g.ldb = mlp->ld[l]; g.BT = mlp->WT[l]; g.ldbt = mlp->ldT[l];
        g.M = (int)n; g.N = mlp->in_dim[l]; g.K = mlp->out_dim[l];
        float* out = DZ[l & 1];
        g.C = out; g.ldc = 256;
        g.aux1 = st.H[l - 1]; g.ldaux1 = 256;
        g.aux2 = st.D[l - 1]; g.ldaux2 = 256;   // X_{l-1}
        if (l == 4) { g.nsplit = 193; g.C2 = DE; g.ldc2 = 64; }
        HN_PROPAGATE((gemm_nn<EPI_REVERSE>(g, s, precision)));
        dz = out; ld_dz = 256;
    }
    HN_PROPAGATE(dw_gemm(dz, 256, 256, st.E, 64, 63, 0));
    HN_PROPAGATE(db_sum(dz, 256, 256, 0));
    if (d_pts) {
        GemmArgs g;
        g.A = dz; g.lda = 256;
        g.B = mlp->W[0]; g.ldb = mlp->ld[0]; g.BT = mlp->WT[0]; g.ldbt = mlp->ldT[0];
        g.M = (int)n; g.N = 63; g.K = 256;
        g.C = DE; g.ldc = 64; g.aux1 = DE; g.ldaux1 = 64;
        HN_PROPAGATE((gemm_nn<EPI_ADD_AUX>(g, s, precision)));
        dx_obj_kernel<<<blocks_for(n * 3, 256), 256, 0, s>>>(st.E, DE, st.EB, d_normal, n, d_pts);
        count_launch();
        HN_CHECK_LAUNCH();
    }
    return HN_OK;
}

}  // extern "C"

// ==========================================================================================
// Object colour field
// ==========================================================================================
namespace hn {

constexpr int CIN_LD = 384;   // 63 + 27 + 256 + 27 = 373, padded
constexpr int CIN_OFF_DIR = 63, CIN_OFF_FEAT = 90, CIN_OFF_NRM = 346, CIN_DIM = 373;

__global__ void color_obj_input_kernel(const float* __restrict__ pts, const float* __restrict__ dirs,
                                       const float* __restrict__ feat, int64_t ld_feat,
                                       const float* __restrict__ normal, int64_t n,
                                       float* __restrict__ CIN) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n * CIN_LD) return;
    int64_t p = i / CIN_LD;
    int j = (int)(i - p * CIN_LD);
    float v = 0.0f;
    if (j < CIN_OFF_DIR) {
        const float x[3] = {pts[p * 3], pts[p * 3 + 1], pts[p * 3 + 2]};
        v = enc3_col(x, 10, j);
    } else if (j < CIN_OFF_FEAT) {
        const float x[3] = {dirs[p * 3], dirs[p * 3 + 1], dirs[p * 3 + 2]};
        v = enc3_col(x, 4, j - CIN_OFF_DIR);
    } else if (j < CIN_OFF_NRM) {
        v = feat[p * ld_feat + (j - CIN_OFF_FEAT)];
    } else if (j < CIN_DIM) {
        const float x[3] = {normal[p * 3], normal[p * 3 + 1], normal[p * 3 + 2]};
        v = enc3_col(x, 4, j - CIN_OFF_NRM);
    }
    CIN[i] = v;
}

// DZ4[p, j] = d_rgb[p,j] * rgb (1-rgb), ld 4
__global__ void sigmoid_bwd_kernel(const float* __restrict__ rgb, const float* __restrict__ d_rgb,
                                   int64_t n, float* __restrict__ DZ) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n * 4) return;
    int64_t p = i >> 2;
    int j = (int)(i & 3);
    float v = 0.0f;
    if (j < 3) {
        float y = rgb[p * 3 + j];
        v = d_rgb[p * 3 + j] * y * (1.0f - y);
    }
    DZ[i] = v;
}

// scatter the input cotangent DCIN [n,384] back to pts / dirs / feat / normal
__global__ void color_obj_input_bwd_kernel(const float* __restrict__ CIN, const float* __restrict__ DCIN,
                                           int64_t n, float* __restrict__ d_pts,
                                           float* __restrict__ d_dirs, float* __restrict__ d_feat,
                                           int64_t ld_dfeat, float* __restrict__ d_normal) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n * 265) return;     // 256 feature columns + 9 encoded coordinates per point
    int64_t p = i / 265;
    int j = (int)(i - p * 265);
    const float* e = CIN + p * CIN_LD;
    const float* g = DCIN + p * CIN_LD;
    if (j < 256) {
        if (d_feat) d_feat[p * ld_dfeat + j] = g[CIN_OFF_FEAT + j];
        return;
    }
    j -= 256;
    int which = j / 3, c = j - which * 3;
    if (which == 0) {
        if (d_pts) d_pts[p * 3 + c] = enc3_jt_from_enc(e, g, 10, c);
    } else if (which == 1) {
        if (d_dirs) d_dirs[p * 3 + c] = enc3_jt_from_enc(e + CIN_OFF_DIR, g + CIN_OFF_DIR, 4, c);
    } else {
        if (d_normal) d_normal[p * 3 + c] = enc3_jt_from_enc(e + CIN_OFF_NRM, g + CIN_OFF_NRM, 4, c);
    }
}

static int check_color_obj_mlp(const hn_mlp_t* m) {
    HN_REQUIRE(m && m->n_layers == 5, "object colour mlp must have 5 layers");
    static const int in_d[5] = {373, 256, 256, 256, 256};
    static const int out_d[5] = {256, 256, 256, 256, 3};
    for (int l = 0; l < 5; ++l) {
        HN_REQUIRE(m->in_dim[l] == in_d[l] && m->out_dim[l] == out_d[l], "object colour mlp layer %d has wrong shape", l);
        HN_REQUIRE(m->ld[l] >= round_up(in_d[l], 4) && m->ld[l] % 4 == 0, "bad ld for layer %d", l);
        HN_REQUIRE(m->W[l] && m->b[l] && aligned16(m->W[l]), "layer %d: null or misaligned weights", l);
    }
    return HN_OK;
}

}  // namespace hn

extern "C" {

int64_t hn_color_obj_stash_floats(int64_t n) { return std::max<int64_t>(n * (CIN_LD + 4 * 256), chain::color_stash_floats(n)); }

int64_t hn_color_obj_ws_floats(int64_t n, int kind) {
    if (kind == HN_WS_BWD) return std::max<int64_t>(n * (2 * 256 + 4 + CIN_LD), chain::color_bwd_ws_floats(n));
    return 4;
}

int hn_color_obj_fwd(const hn_mlp_t* mlp, const float* pts, const float* dirs, const float* feat,
                     int64_t ld_feat, const float* normal, int64_t n, float* rgb, float* stash,
                     int64_t stash_floats, int precision, hn_stream_t stream) {
    HN_PROPAGATE(check_color_obj_mlp(mlp));
    const bool s16 = precision == HN_TC_MIXED16;      // same chain arithmetic, 16-bit dW-ready stash (chain_color.cu)
    precision = base_precision(precision);
    HN_REQUIRE(precision_supported(precision), "hn_color_obj_fwd: precision %d not supported", precision);
    HN_REQUIRE(n >= 0 && n < (1ll << 31), "n_pts out of range");
    if (n == 0) return HN_OK;
    HN_REQUIRE(pts && dirs && feat && normal && rgb && stash, "hn_color_obj_fwd: null pointer");
    HN_REQUIRE(stash_floats >= hn_color_obj_stash_floats(n) && aligned16(stash), "stash too small or misaligned");
    cudaStream_t s = (cudaStream_t)stream;
    if (precision == HN_TC_BF16X3) return chain::launch_color_fwd(mlp, pts, dirs, feat, ld_feat, normal, n, rgb, stash, s, s16);
    float* CIN = stash;
    float* R[4];
    for (int l = 0; l < 4; ++l) R[l] = stash + n * CIN_LD + (int64_t)l * n * 256;
    color_obj_input_kernel<<<blocks_for(n * CIN_LD, 256), 256, 0, s>>>(pts, dirs, feat, ld_feat, normal, n, CIN);
    count_launch();
    HN_CHECK_LAUNCH();
    for (int l = 0; l < 4; ++l) {
        GemmArgs g;
        g.A = l == 0 ? CIN : R[l - 1]; g.lda = l == 0 ? CIN_LD : 256;
        g.B = mlp->W[l]; g.ldb = mlp->ld[l]; g.BT = mlp->WT[l]; g.ldbt = mlp->ldT[l];
        g.M = (int)n; g.N = 256; g.K = mlp->in_dim[l];
        g.C = R[l]; g.ldc = 256; g.bias = mlp->b[l];
        HN_PROPAGATE((gemm_nt<EPI_BIAS_RELU>(g, s, precision)));
    }
    GemmArgs g;
    g.A = R[3]; g.lda = 256; g.B = mlp->W[4]; g.ldb = mlp->ld[4]; g.BT = mlp->WT[4]; g.ldbt = mlp->ldT[4];
    g.M = (int)n; g.N = 3; g.K = 256; g.C = rgb; g.ldc = 3; g.bias = mlp->b[4];
    HN_PROPAGATE((gemm_nt<EPI_BIAS_SIGMOID>(g, s, precision)));
    return HN_OK;
}

int hn_color_obj_bwd(const hn_mlp_t* mlp, int64_t n, float* stash, const float* rgb, const float* d_rgb,
                     float* d_pts, float* d_dirs, float* d_feat, int64_t ld_dfeat, float* d_normal,
                     const hn_mlp_grad_t* grad, float* ws, int64_t ws_floats, int precision,
                     hn_stream_t stream) {
    HN_PROPAGATE(check_color_obj_mlp(mlp));
    const bool s16 = precision == HN_TC_MIXED16;
    precision = base_precision(precision);
    HN_REQUIRE(precision_supported(precision), "hn_color_obj_bwd: precision %d not supported", precision);
    HN_REQUIRE(n >= 0 && n < (1ll << 31), "n_pts out of range");
    if (n == 0) return HN_OK;
    HN_REQUIRE(stash && rgb && d_rgb && ws, "hn_color_obj_bwd: null pointer");
    HN_REQUIRE(ws_floats >= hn_color_obj_ws_floats(n, HN_WS_BWD) && aligned16(ws), "workspace too small or misaligned");
    cudaStream_t s = (cudaStream_t)stream;
    if (precision == HN_TC_BF16X3)
        return chain::launch_color_bwd(mlp, n, stash, rgb, d_rgb, d_pts, d_dirs, d_feat, ld_dfeat, d_normal, grad, ws, s, s16);
    float* CIN = stash;
    float* R[4];
    for (int l = 0; l < 4; ++l) R[l] = stash + n * CIN_LD + (int64_t)l * n * 256;
    float* DZ[2] = {ws, ws + n * 256};
    float* DZ4 = ws + 2 * n * 256;
    float* DCIN = DZ4 + n * 4;
    const int splits_target = 2 * sm_count();
    auto dw_gemm = [&](const float* P, int64_t ldp, int out, const float* Q, int64_t ldq, int in, int l) -> int {
        if (!grad || !grad->dW[l]) return HN_OK;
        GemmArgs g;
        g.A = P; g.lda = ldp; g.B = Q; g.ldb = ldq;
        g.M = out; g.N = in; g.K = (int)n;
        g.C = grad->dW[l]; g.ldc = mlp->ld[l];
        int tiles = (int)(ceil_div(out, GBM) * ceil_div(in, GBN));
        int splits = (int)max((int64_t)1, min((int64_t)ceil_div(splits_target, tiles), ceil_div(n, 256)));
        return gemm_tn(g, s, precision, splits);
    };
    sigmoid_bwd_kernel<<<blocks_for(n * 4, 256), 256, 0, s>>>(rgb, d_rgb, n, DZ4);
    count_launch();
    HN_CHECK_LAUNCH();
    const float* dz = DZ4;
    int64_t ld_dz = 4;
    for (int l = 4; l >= 1; --l) {
        HN_PROPAGATE(dw_gemm(dz, ld_dz, mlp->out_dim[l], R[l - 1], 256, 256, l));
        if (grad && grad->db[l]) HN_PROPAGATE(launch_colsum(dz, ld_dz, n, mlp->out_dim[l], 1.0f, grad->db[l], s));
        GemmArgs g;
        g.A = dz; g.lda = ld_dz; g.B = mlp->W[l]; g.ldb = mlp->ld[l]; g.BT = mlp->WT[l]; g.ldbt = mlp->ldT[l];
        g.M = (int)n; g.N = 256; g.K = mlp->out_dim[l];
        g.C = DZ[l & 1]; g.ldc = 256; g.aux1 = R[l - 1]; g.ldaux1 = 256;
        HN_PROPAGATE((gemm_nn<EPI_RELU_BWD>(g, s, precision)));
        dz = DZ[l & 1]; ld_dz = 256;
    }
    HN_PROPAGATE(dw_gemm(dz, 256, 256, CIN, CIN_LD, CIN_DIM, 0));
    if (grad && grad->db[0]) HN_PROPAGATE(launch_colsum(dz, 256, n, 256, 1.0f, grad->db[0], s));
    if (d_pts || d_dirs || d_feat || d_normal) {
        GemmArgs g;
        g.A = dz; g.lda = 256; g.B = mlp->W[0]; g.ldb = mlp->ld[0]; g.BT = mlp->WT[0]; g.ldbt = mlp->ldT[0];
        g.M = (int)n; g.N = CIN_DIM; g.K = 256; g.C = DCIN; g.ldc = CIN_LD;
        HN_PROPAGATE((gemm_nn<EPI_STORE>(g, s, precision)));
        color_obj_input_bwd_kernel<<<blocks_for(n * 265, 256), 256, 0, s>>>(CIN, DCIN, n, d_pts, d_dirs, d_feat,
                                                                           ld_dfeat, d_normal);
        count_launch();
        HN_CHECK_LAUNCH();
    }
    return HN_OK;
}

}  // extern "C"
