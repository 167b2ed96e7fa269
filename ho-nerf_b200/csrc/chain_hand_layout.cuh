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
    // feature-side contractions out of the chain, in 256-column chunks of the 1386 HALO features (the last one 106 -> 112):
    // [0]: d @ W_4[:, 256:], [1]: d @ W_0;  B(n = feature, k = output)
    uint32_t nnf16_off[2][6];   // fp16 pairs (normal sweep -> FB)
    uint32_t nnf_off[2][6];     // bf16 pairs (reverse sweep -> DF)
    // feature-side contractions INTO the chain (value trunk): F @ W_0^T ([0]) and F @ W_4[:, 256:]^T ([1]) as fp16 pairs,
    // B(n = output 256, k = feature, 22 k-blocks of 64 = 1408 >= 1386); A = the fp16 pair tiles halo_feature16 writes
    uint32_t ntf16_off[2];
    uint32_t total;
};
constexpr int HAND_F_KBLOCKS = 22;
constexpr int HAND_F16_KB_BYTES = 2 * 128 * 128;                           // one k-block of a 128-point feature tile: hi + lo, 32 KB
constexpr int HAND_F16_TILE_BYTES = HAND_F_KBLOCKS * HAND_F16_KB_BYTES;    // 704 KB per 128 points
constexpr int HAND_F_CHUNKS = 6;
__host__ __device__ inline int hand_f_chunk_n(int ch) { return ch < 5 ? 256 : 112; }       // UMMA N of chunk ch
__host__ __device__ inline int hand_f_chunk_valid(int ch) { return ch < 5 ? 256 : 106; }   // features in chunk ch
inline HandLayout hand_layout() {
    HandLayout L = {};
    uint32_t off = 0;
    for (int l = 1; l <= 8; ++l)
        for (int h = 0; h < 2; ++h) { L.nth_off[l][h] = off; off += b_operand_bytes(128, 4); }
    for (int l = 1; l <= 7; ++l) { L.nn16_off[l] = off; off += b_operand_bytes(256, 4); }
    for (int l = 1; l <= 7; ++l) { L.nt_off[l] = off; off += b_operand_bytes(256, 4); }
    for (int l = 1; l <= 8; ++l) { L.nn_off[l] = off; off += b_operand_bytes(256, 4); }
    for (int w = 0; w < 2; ++w)
        for (int ch = 0; ch < 6; ++ch) { L.nnf16_off[w][ch] = off; off += b_operand_bytes(ch < 5 ? 256 : 112, 4); }
    for (int w = 0; w < 2; ++w)
        for (int ch = 0; ch < 6; ++ch) { L.nnf_off[w][ch] = off; off += b_operand_bytes(ch < 5 ? 256 : 112, 4); }
    for (int w = 0; w < 2; ++w) { L.ntf16_off[w] = off; off += b_operand_bytes(256, 22); }
    L.total = off;
    return L;
}

// fields_hand.cu <-> chain16_hand.cu
int64_t hand16_stash_floats(int64_t n);
int64_t hand16_bwd_ws_floats(int64_t n);

}  // namespace chain
}  // namespace hn
