// Layout of the packed tcgen05 operands of the hand SDF net's 256 x 256 layers (hn_sdf_hand_chain_pack), appended to the
// per-layer HN_TC_BF16X3 operands (gemm_bx3.cuh) in the same buffer: chain16_hand.cu.
#pragma once
#include "chain_common.cuh"

namespace hn {
namespace chain {

// Layers 1..7 are 256 -> 256 (layer 4 restricted to the h3 part of its input: columns 0..255), layer 8 is 256 -> 1 + 256 and
// is packed without its sdf row, which the epilogues apply as a rank-one term.  NT: B(n = output, k = input) for a @ W^T,
// NN: B(n = input, k = output) for d @ W.
struct HandLayout {
    uint32_t nth_off[9][2];     // a @ W_l^T, l = 1..8, fp16 pairs, two 128-row halves (value trunk + feature head)
    uint32_t nn16_off[8];       // d @ W_l, l = 1..7, fp16 pairs (normal sweep)
    uint32_t nt_off[8];         // u @ W_l^T, l = 1..7, bf16 pairs (tangent sweep)
    uint32_t nn_off[9];         // dz @ W_l, l = 1..8, bf16 pairs (reverse sweep)
    uint32_t total;
};
inline HandLayout hand_layout() {
    HandLayout L = {};
    uint32_t off = 0;
    for (int l = 1; l <= 8; ++l)
        for (int h = 0; h < 2; ++h) { L.nth_off[l][h] = off; off += b_operand_bytes(128, 4); }
    for (int l = 1; l <= 7; ++l) { L.nn16_off[l] = off; off += b_operand_bytes(256, 4); }
    for (int l = 1; l <= 7; ++l) { L.nt_off[l] = off; off += b_operand_bytes(256, 4); }
    for (int l = 1; l <= 8; ++l) { L.nn_off[l] = off; off += b_operand_bytes(256, 4); }
    L.total = off;
    return L;
}

// fields_hand.cu <-> chain16_hand.cu
int64_t hand16_stash_floats(int64_t n);
int64_t hand16_bwd_ws_floats(int64_t n);

}  // namespace chain
}  // namespace hn
