// Layout of the packed tcgen05 operands of the object SDF net inside the chain buffer (hn_sdf_obj_chain_pack).
#pragma once
#include "chain_common.cuh"

namespace hn {
namespace chain {

// Packed operands of the object SDF net.  NT[l]: B(n = output feature, k = input feature) for
// a @ W_l^T (value trunk, tangent sweep); NN[l]: B(n = input feature, k = output feature) for
// d @ W_l (normal sweep, reverse sweep).  The output layer is packed without its sdf row (row 0),
// which is applied as a rank-one term by the epilogues.
struct ObjLayout {
    uint32_t nt_off[9], nn_off[9];
    uint32_t nt16_off[9];         // a @ W_l^T operands again as fp16 (hi + lo) pairs: the forward value trunk
    uint32_t nth_off[8][2];       // layers 0..7 once more as two 128-row halves (fp16 pairs): chain_ts.cu
    // HN_TC_MIXED16 (chain16_obj.cu)
    uint32_t nth8_off[2];         // feature head (rows 1..256 of the output layer) as two 128-row halves, fp16 pairs: a @ W_8^T
    uint32_t nn16_off[8];         // normal sweep d @ W_l as fp16 pairs (same shapes as nn_off)
    uint16_t nt_n[9], nn_n[9];
    uint8_t nt_kb[9], nn_kb[9];
    uint32_t total;
};
inline ObjLayout obj_layout() {
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
    for (int l = 0; l < 9; ++l) {
        L.nt16_off[l] = off;
        off += b_operand_bytes(L.nt_n[l], L.nt_kb[l]);
    }
    for (int l = 0; l < 8; ++l)
        for (int h = 0; h < 2; ++h) {
            L.nth_off[l][h] = off;
            off += b_operand_bytes(128, L.nt_kb[l]);
        }
    for (int h = 0; h < 2; ++h) { L.nth8_off[h] = off; off += b_operand_bytes(128, 4); }
    for (int l = 0; l < 8; ++l) { L.nn16_off[l] = off; off += b_operand_bytes(L.nn_n[l], L.nn_kb[l]); }
    L.total = off;
    return L;
}


}  // namespace chain
}  // namespace hn
