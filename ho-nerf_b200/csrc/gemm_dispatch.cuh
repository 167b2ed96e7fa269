// Precision-aware dispatch of the dense contractions:
//   HN_SIMT_FP32  -> gemm_simt.cuh (fp32 FFMA, verification path)
//   HN_TC_TF32    -> tcgen05, single-pass TF32 operands everywhere (fast, ~1e-3 relative per layer)
//   HN_TC_BF16X3  -> the fused chain kernels where they exist (object nets); elsewhere gemm_bx3.cuh (pre-packed bf16
//                    hi/lo weights, three bf16 MMAs per product) when the net carries packed operands, else TF32X3
//   HN_TC_TF32X3  -> tcgen05, split (hi+lo) TF32 operands everywhere: three MMAs per product, ~fp32
//                    accuracy; this is the mode that meets the north-star tolerances with margin
//                    (weight-norm backward amplifies TF32-level errors of dW by ~10x, and the colour
//                    net sees sin/cos(8 n) of the normals, so single-pass sweeps are not enough).
#pragma once
#include "gemm_simt.cuh"
#include "gemm_tc.cuh"
#include "gemm_bx3.cuh"

namespace hn {

enum GemmRole { ROLE_VALUE = 0, ROLE_OTHER = 1 };

inline bool precision_supported(int p) {
    return p == HN_SIMT_FP32 || p == HN_TC_TF32 || p == HN_TC_TF32X3 || p == HN_TC_BF16X3 || p == HN_TC_MIXED16;
}
// HN_TC_MIXED16 only exists for the object SDF field (chain16_obj.cu); every other entry point runs HN_TC_BF16X3
inline int base_precision(int p) { return p == HN_TC_MIXED16 ? HN_TC_BF16X3 : p; }

// C = epi(A @ W^T)
template <int EPI>
int gemm_nt(const GemmArgs& g, cudaStream_t s, int precision, int role = ROLE_OTHER) {
    (void)role;
    if (precision == HN_SIMT_FP32) return launch_gemm<true, true, EPI>(g, s);
    if (precision == HN_TC_BF16X3 && g.Bp) return launch_gemm_bx3<EPI>(g, s);
    if (precision == HN_TC_TF32X3 || precision == HN_TC_BF16X3) return launch_gemm_tc<false, 3, EPI>(g, s);
    return launch_gemm_tc<false, 1, EPI>(g, s);
}
// C = epi(A @ W)
template <int EPI>
int gemm_nn(const GemmArgs& g, cudaStream_t s, int precision, int role = ROLE_OTHER) {
    (void)role;
    if (precision == HN_SIMT_FP32) return launch_gemm<true, false, EPI>(g, s);
    if (precision == HN_TC_BF16X3 && g.BTp) {
        GemmArgs t = g;
        t.Bp = g.BTp; t.bp_tile_bytes = g.btp_tile_bytes; t.bp_kb0 = 0;
        return launch_gemm_bx3<EPI>(t, s);
    }
    if (g.BT) {
        // a pre-transposed copy of the weights exists: run as x @ (W^T)^T, the cheap staging path
        GemmArgs t = g;
        t.B = g.BT; t.ldb = g.ldbt; t.BT = nullptr;
        if (precision == HN_TC_TF32X3 || precision == HN_TC_BF16X3) return launch_gemm_tc<false, 3, EPI>(t, s);
        return launch_gemm_tc<false, 1, EPI>(t, s);
    }
    if (precision == HN_TC_TF32X3 || precision == HN_TC_BF16X3) return launch_gemm_tc<true, 3, EPI>(g, s);
    return launch_gemm_tc<true, 1, EPI>(g, s);
}
// C += A^T @ B over points (weight gradients)
inline int gemm_tn(const GemmArgs& g, cudaStream_t s, int precision, int splits) {
    if (precision == HN_SIMT_FP32) return launch_gemm<false, false, EPI_ATOMIC>(g, s, splits);
    if (precision == HN_TC_TF32X3 || precision == HN_TC_BF16X3) return launch_gemm_tc_tn<3>(g, s, splits);
    return launch_gemm_tc_tn<1>(g, s, splits);
}

}  // namespace hn
