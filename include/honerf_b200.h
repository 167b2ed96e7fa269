/* honerf_b200 -- C ABI of the B200-native NeuS volume-rendering hot path of HO-NeRF.
 *
 * The reference (iscas3dv/HO-NeRF) has no FFI layer: its hot path is plain PyTorch
 * (utils/fields.py, utils/renderer.py, utils/renderer_batch.py).  This header is the boundary a
 * maintainer would bind instead: every entry point names the reference code it replaces
 * (file:line).  INTEGRATION.md shows the ctypes stub.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer unless stated otherwise; all tensors are dense row-major
 *     fp32 (indices int64), 16-byte aligned, leading dimensions multiples of 4 floats;
 *   - functions never allocate, never synchronise and enqueue all work on `stream`
 *     (a cudaStream_t); outputs, stashes and workspaces are caller-allocated -- the *_floats()
 *     queries give their sizes.  The one exception is the multi-GPU set-up (hn_peer_alloc / _open /
 *     _close / _free): peer-mapped memory cannot be caller-allocated torch memory, these four run
 *     once, outside the step;
 *   - return value: HN_OK (0) or a negative hn_status; hn_last_error() gives the message of the
 *     calling thread's last failure;
 *   - there is NO CPU fallback: with no CUDA device every compute entry point fails.
 */
#ifndef HONERF_B200_H
#define HONERF_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define HN_API __attribute__((visibility("default")))

typedef void* hn_stream_t; /* cudaStream_t */

enum hn_status { HN_OK = 0, HN_ERR_ARG = -1, HN_ERR_CUDA = -2, HN_ERR_UNSUPPORTED = -3 };

/* Arithmetic of the dense contractions.  HN_SIMT_FP32 is the verification path (fp32 FFMA);
 * the HN_TC_* values run on tcgen05 tensor cores with fp32 accumulation in TMEM:
 *   HN_TC_TF32   every contraction with single-pass TF32 operands (fast; ~1e-3 relative per layer);
 *   HN_TC_TF32X3 every contraction with split (hi+lo) TF32 operands -- three MMAs per product,
 *                ~fp32 accuracy; the mode that meets the parity tolerances (1e-3 abs on colour/SDF,
 *                1e-2 relative on gradients) with margin.
 *   HN_TC_BF16X3 fused tile-chain kernels: a 128-point tile's activations stay in shared memory
 *                between layers as split (hi+lo) bf16 operands, three bf16 MMAs per product, weights
 *                streamed from the pre-packed chain buffer (hn_*_chain_pack).  ~16 mantissa bits:
 *                8e-6 abs on the SDF, 5e-5 relative on weight gradients (oracle/analytic.py).
 *                Entry points without a chain kernel run their HN_TC_TF32X3 path.
 *   HN_TC_MIXED16 the object SDF field (utils/fields.py:316-347 and its double backward) with a 16-bit activation
 *                stash and tensor-memory-resident operands (csrc/chain16*.cu): the value trunk keeps three fp16 MMAs
 *                per product (sdf / feature ~2e-6), the normal / tangent / reverse sweeps round their A operand to one
 *                16-bit value (two MMAs per product), the weight gradients use the stored bf16 operands (one MMA):
 *                normals ~4e-4, weight gradients <= 6e-3 relative (north star: 1e-2).  Every other entry point runs
 *                its HN_TC_BF16X3 path. */
enum hn_precision { HN_SIMT_FP32 = 0, HN_TC_TF32 = 1, HN_TC_TF32X3 = 2, HN_TC_BF16X3 = 3, HN_TC_MIXED16 = 4 };

HN_API const char* hn_last_error(void);
HN_API int hn_version(void);
/* Number of kernels this library has launched on behalf of the calling process (all streams). */
HN_API int64_t hn_launch_count(void);

/* Profiling aid for bench.py's roofline: when enabled, CUDA events are recorded on the launching
 * stream around every dense-contraction (GEMM) launch; collect() synchronises them and returns the
 * summed device time and the number of launches since the last reset. */
HN_API int hn_timing_enable(int on);
HN_API int hn_timing_reset(void);
HN_API int hn_timing_collect(double* total_ms, int64_t* n_launches); /* HOST pointers */
/* Same, split by kernel family: 0 per-layer contractions, 1 chain SDF-only, 2 chain SDF forward (value + feature +
 * normal sweep), 3 chain SDF backward (tangent + reverse sweeps), 4 chain weight gradients, 5 chain colour forward,
 * 6 chain colour backward.  HOST arrays of n_tags entries. */
HN_API int hn_timing_collect_tags(double* ms_per_tag, int64_t* launches_per_tag, int n_tags);

/* ---------------------------------------------------------------------------------------------
 * Weight-normalised MLP parameters (nn.utils.weight_norm, dim=0:
 * utils/fields.py:120-121, 216-217, 307-308, 382-383).
 * ------------------------------------------------------------------------------------------- */
#define HN_MAX_LAYERS 12

typedef struct hn_mlp {
    int32_t n_layers;
    int32_t in_dim[HN_MAX_LAYERS];
    int32_t out_dim[HN_MAX_LAYERS];
    int32_t ld[HN_MAX_LAYERS];     /* leading dimension of W[l], >= round_up(in_dim, 4) */
    const float* W[HN_MAX_LAYERS]; /* effective weights g*v/||v|| (times post_scale), [out, ld] */
    const float* b[HN_MAX_LAYERS]; /* bias [out] */
    /* transposed copy [in, ldT] (ldT >= round_up(out, 4), zero padded): the tensor-core path reads
     * both x @ W^T and d @ W as K-major operands.  May be NULL for HN_SIMT_FP32. */
    const float* WT[HN_MAX_LAYERS];
    int32_t ldT[HN_MAX_LAYERS];
    /* HN_TC_BF16X3: bf16 hi/lo operands pre-swizzled for tcgen05 (hn_*_chain_pack); else NULL */
    const void* chain;
    int64_t chain_bytes;
} hn_mlp_t;

typedef struct hn_mlp_grad {
    float* dW[HN_MAX_LAYERS]; /* same layout as W; ACCUMULATED into (caller zeroes) */
    float* db[HN_MAX_LAYERS]; /* [out]; ACCUMULATED into */
} hn_mlp_grad_t;

/* W[o, :in] = post_scale * g[o] * v[o, :] / ||v[o, :]||, padding columns [in, ld) zeroed; when WT
 * is non-NULL also WT[k, o] = W[o, k] ([in, ldT], padding columns [out, ldT) zeroed).
 * Replaces the per-call `_weight_norm` recomputation of every Linear. */
HN_API int hn_wn_pack(const float* v, const float* g, int out_dim, int in_dim, int ld,
                      float post_scale, float* W, float* WT, int ldT, hn_stream_t stream);
/* Same with `gap` zero columns inserted after input column `gap_at` of the packed layouts (packed
 * in_dim = in_dim + gap; ld >= in_dim + gap).  Used by the hand colour net, whose input row is
 * [xyz_feature 1386 | pad 2 | feature 256 | enc(normal) 27]. */
HN_API int hn_wn_pack_gap(const float* v, const float* g, int out_dim, int in_dim, int ld,
                          float post_scale, int gap_at, int gap, float* W, float* WT, int ldT,
                          hn_stream_t stream);
HN_API int hn_wn_bwd_gap(const float* v, const float* g, const float* dW, int out_dim, int in_dim,
                         int ld, float post_scale, int gap_at, int gap, float* dv, float* dg,
                         hn_stream_t stream);
/* (dv, dg) from dW; dW is the gradient w.r.t. the PACKED weight (so it is
 * multiplied by post_scale first).  dv [out,in] and dg [out] are overwritten. */
HN_API int hn_wn_bwd(const float* v, const float* g, const float* dW, int out_dim, int in_dim,
                     int ld, float post_scale, float* dv, float* dg, hn_stream_t stream);

/* All layers of a net in ONE launch each (the per-layer entry points above cost a launch per layer and step).
 * jobs is a HOST array of n <= HN_MAX_LAYERS entries; pack uses {v, g, out_dim, in_dim, ld, post_scale, gap_at,
 * gap, W, WT, ldT}, bwd uses {v, g, dW, out_dim, in_dim, ld, post_scale, gap_at, gap, dv, dg}. */
typedef struct hn_wn_job {
    const float* v;
    const float* g;
    const float* dW;
    float* W;
    float* WT;
    float* dv;
    float* dg;
    int32_t out_dim, in_dim, ld, ldT, gap_at, gap;
    float post_scale;
    int32_t pad_;
} hn_wn_job_t;
HN_API int hn_wn_pack_batch(const hn_wn_job_t* jobs, int n, hn_stream_t stream);
HN_API int hn_wn_bwd_batch(const hn_wn_job_t* jobs, int n, hn_stream_t stream);

/* HN_TC_BF16X3 for nets without a fused chain kernel (the hand field, utils/fields.py:56-240): every layer's weights
 * pre-packed as bf16 hi/lo tcgen05 tiles, both as the operand of x @ W^T and of d @ W (needs W and WT of every layer).
 * Point hn_mlp_t::chain at the packed buffer; the per-layer contractions then run three bf16 MMAs per product instead
 * of the split-TF32 kernel.  hn_mlp_bx3_bytes: size of that buffer. */
HN_API int64_t hn_mlp_bx3_bytes(const hn_mlp_t* m);
HN_API int hn_mlp_bx3_pack(const hn_mlp_t* m, void* buf, int64_t bytes, hn_stream_t stream);
/* Hand SDF net (SDFNetwork, utils/fields.py:56-177): the per-layer operands of hn_mlp_bx3_pack followed by the tile-chain
 * operands of its 256 x 256 layers.  With this buffer in mlp->chain, HN_TC_MIXED16 runs those layers of hn_sdf_hand_sdf /
 * _fwd / _bwd as persistent chain kernels with the activations in tensor memory (csrc/chain16_hand.cu); that path keeps a
 * 16-bit stash and computes NO weight gradients (hn_sdf_hand_bwd rejects grad != NULL): pose fitting and rendering.  A
 * forward whose backward needs weight gradients is made with HN_TC_BF16X3. */
HN_API int64_t hn_sdf_hand_chain_bytes(const hn_mlp_t* m);
HN_API int hn_sdf_hand_chain_pack(const hn_mlp_t* m, void* buf, int64_t bytes, hn_stream_t stream);

/* Adam over one flat fp32 parameter buffer (torch.optim.Adam semantics; the reference builds one Adam over all
 * networks' parameters, exp_runner.py:83, and rewrites param_groups[i]['lr'] every iteration, exp_runner.py:258-268).
 * p, m, v [n] are updated in place from g [n] * grad_scale; `step` is a DEVICE float holding the 1-based step count (the
 * caller increments it before the call) and `lr_dev` (may be NULL: use `lr`) a DEVICE float holding the learning rate, so
 * the launch is the same every step and can be replayed from a CUDA graph while the schedule still applies.  `skipped`
 * (may be NULL) holds per 32-element block of the buffer how many steps that block's parameter received no gradient:
 * its bias correction uses step - skipped, as torch's per-parameter step does.  p .. v and the blocks of `skipped`
 * refer to the SAME element range (the caller offsets all of them for a sub-range that starts on a block boundary).
 * The hyper-parameters are doubles: 1 - beta is formed in double and rounded once, as torch.optim.Adam does. */
HN_API int hn_adam_flat(float* p, const float* g, float* m, float* v, int64_t n, const float* step, const float* lr_dev,
                        const float* skipped, double lr, double beta1, double beta2, double eps, double weight_decay,
                        double grad_scale, hn_stream_t stream);

/* Multi-GPU training step (one process per GPU, rays sharded): the step's ONE exchange -- the all-reduce of
 * the flat gradient buffer -- fused with hn_adam_flat into one kernel over NVLink / NVSwitch peer memory (csrc/peer.cu)
 * instead of ncclAllReduce followed by the Adam launch.  The reference trains on one GPU (exp_runner.py:83-90, :206-230:
 * loss.backward(); optimizer.step()), so there is no upstream interface to mirror: these are the calls a data-parallel
 * exp_runner would bind.
 *   hn_peer_block_bytes  size of one rank's peer block for n parameters: [g : n floats | reduced g | barrier flags]
 *   hn_peer_alloc        cudaMalloc + zero + cudaIpcGetMemHandle: *ptr = the block (its first n floats are the rank's
 *                        flat gradient buffer), handle64 = the 64-byte handle the other ranks open
 *   hn_peer_open/_close  map / unmap another rank's block (cudaIpcOpenMemHandle, peer access enabled lazily)
 *   hn_peer_free         release an hn_peer_alloc block
 *   hn_peer_adam_flat    blocks[world] (HOST array of the ranks' block addresses in this process, own block at [rank]):
 *                        p, m, v [n] <- Adam(sum over ranks of g * grad_scale); two-shot (each rank sums 1/world of the
 *                        buffer in rank order, every rank reads the sums), so all ranks hold bit-identical parameters.
 *                        mode 1: p [n] = grad_scale * sum (plain all-reduce into a local buffer, m / v / step unused).
 *                        `epoch` [148] device words (zero at start, owned by the kernel) carry the barrier generation so
 *                        the launch replays from a CUDA graph; *err (device word, zero at start) becomes non-zero if a
 *                        peer did not arrive within 60 s -- the kernel never hangs the GPU, the caller must check it.
 *                        Every rank must make the same sequence of calls (same n, world, mode).  n % 4 == 0. */
HN_API int64_t hn_peer_block_bytes(int64_t n, int world);
HN_API int hn_peer_alloc(int64_t bytes, void** ptr, uint8_t* handle64);
HN_API int hn_peer_open(const uint8_t* handle64, void** ptr);
HN_API int hn_peer_close(void* ptr);
HN_API int hn_peer_free(void* ptr);
HN_API int hn_peer_adam_flat(float* p, float* m, float* v, int64_t n, const void* const* blocks, int rank, int world,
                             uint32_t* epoch, uint32_t* err, int mode, const float* step, const float* lr_dev, double lr,
                             double beta1, double beta2, double eps, double weight_decay, double grad_scale,
                             hn_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * Object SDF field: SDFNetwork_OBJ.forward / .sdf / .gradient (utils/fields.py:316-347) as ONE
 * operator (value + feature + analytic normal) with a hand-written second-order backward.
 * mlp: 9 layers 63->256->256->256->193->[cat 63]->256->256->256->256->257; W[4] must be packed
 * with post_scale = 1/sqrt(2) (the skip concat's division is folded into it).
 * ------------------------------------------------------------------------------------------- */
enum hn_ws_kind { HN_WS_SDF_ONLY = 0, HN_WS_FWD = 1, HN_WS_BWD = 2 };
HN_API int64_t hn_sdf_obj_stash_floats(int64_t n_pts);
HN_API int64_t hn_sdf_obj_ws_floats(int64_t n_pts, int ws_kind);

/* Pack the object SDF net's weights (W and WT of `mlp`, already weight-normalised by hn_wn_pack)
 * into the HN_TC_BF16X3 chain buffer: bf16 hi/lo tiles in the tcgen05 shared-memory layout. */
HN_API int64_t hn_sdf_obj_chain_bytes(void);
HN_API int hn_sdf_obj_chain_pack(const hn_mlp_t* mlp, void* chain, int64_t chain_bytes,
                                 hn_stream_t stream);

/* Diagnostics: device buffer of int64 [n_ctas][4] that the chain kernels fill with cycle counters
 * {MMA warp waiting for activations, waiting for weights, total, epilogue waiting for the
 * accumulator}; NULL disables. */
HN_API int hn_chain_set_prof(void* buf);
/* Tuning: CTA i of the forward / backward chain kernels starts (i % 4) * cycles late, which de-phases the
 * HBM-heavy epilogues of different SMs (0 = all CTAs in lock step). */
HN_API int hn_chain_set_stagger(int fwd_cycles, int bwd_cycles);

/* sdf[n] = SDFNetwork_OBJ.sdf(pts) (utils/fields.py:330-331); no stash, no normal. */
/* extract_geometry's lattice (utils/renderer.py:262-278) in ONE launch: u[ix, iy, iz] = sdf(xs[ix], ys[iy], zs[iz]) / scale with
 * the lattice points generated inside the SDF kernel (ij-meshgrid order; the axes are the caller's linspace values): no
 * [nx ny nz, 3] point tensor is written or read.  Tensor-core chain kernel (needs hn_sdf_obj_chain_pack). */
HN_API int hn_sdf_obj_grid(const hn_mlp_t* mlp, const float* xs, int nx, const float* ys, int ny, const float* zs, int nz,
                           float inv_scale, float* u, hn_stream_t stream);
HN_API int hn_sdf_obj_sdf(const hn_mlp_t* mlp, const float* pts, int64_t n_pts, float inv_scale,
                          float* sdf, float* ws, int64_t ws_floats, int precision,
                          hn_stream_t stream);
/* sdf [n], feat [n, ld_feat>=256], normal [n,3] = d sdf / d pts (replaces the second forward +
 * autograd.grad(create_graph=True) of utils/fields.py:336-347).  `stash` keeps what the backward
 * needs. */
HN_API int hn_sdf_obj_fwd(const hn_mlp_t* mlp, const float* pts, int64_t n_pts, float inv_scale,
                          float* sdf, float* feat, int64_t ld_feat, float* normal, float* stash,
                          int64_t stash_floats, float* ws, int64_t ws_floats, int precision,
                          hn_stream_t stream);
/* Backward of hn_sdf_obj_fwd given cotangents of (sdf, feat, normal): accumulates dW/db, and
 * writes d_pts [n,3] when non-NULL.  This contains the Hessian-vector products the reference
 * obtains by differentiating through autograd.grad (SoftplusBackwardBackward).  The stash is
 * consumed (overwritten). d_feat may be NULL (treated as zero). */
HN_API int hn_sdf_obj_bwd(const hn_mlp_t* mlp, int64_t n_pts, float inv_scale, float* stash,
                          const float* d_sdf, const float* d_feat, int64_t ld_dfeat,
                          const float* d_normal, float* d_pts, const hn_mlp_grad_t* grad,
                          float* ws, int64_t ws_floats, int precision, hn_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * Object colour field: RenderingNetwork_OBJ.forward (utils/fields.py:387-405).
 * mlp: 5 layers 373->256->256->256->256->3, ReLU, sigmoid.
 * ------------------------------------------------------------------------------------------- */
HN_API int64_t hn_color_obj_stash_floats(int64_t n_pts);
HN_API int64_t hn_color_obj_ws_floats(int64_t n_pts, int ws_kind);
/* HN_TC_BF16X3 chain operands of the colour net (see hn_sdf_obj_chain_pack). */
HN_API int64_t hn_color_obj_chain_bytes(void);
HN_API int hn_color_obj_chain_pack(const hn_mlp_t* mlp, void* chain, int64_t chain_bytes,
                                   hn_stream_t stream);
HN_API int hn_color_obj_fwd(const hn_mlp_t* mlp, const float* pts, const float* dirs,
                            const float* feat, int64_t ld_feat, const float* normal,
                            int64_t n_pts, float* rgb, float* stash, int64_t stash_floats,
                            int precision, hn_stream_t stream);
/* d_pts, d_dirs, d_normal [n,3] and d_feat [n, ld_dfeat] are overwritten (any may be NULL);
 * grad may be NULL (weights frozen: pose fitting, SURVEY D-11). */
HN_API int hn_color_obj_bwd(const hn_mlp_t* mlp, int64_t n_pts, float* stash, const float* rgb,
                            const float* d_rgb, float* d_pts, float* d_dirs, float* d_feat,
                            int64_t ld_dfeat, float* d_normal, const hn_mlp_grad_t* grad,
                            float* ws, int64_t ws_floats, int precision, hn_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * Hand SDF field (HALO pose-conditioned): anerf_emb_point(_batch) + SDFNetwork.forward/.sdf/.gradient
 * (utils/fields.py:22-52, 132-177) as one operator, second-order backward included.
 * pts [n,3] are grouped by frame: point p belongs to frame p / pts_per_frame; bt_inv
 * [frames,21,4,4] (row-major), T_pose [frames,21,3].  mlp: 9 layers 1386->256 x3 ->256->[cat 1386]
 * ->256 x4 ->257 (W[4] packed with post_scale 1/sqrt(2)); no division by `scale` (SURVEY A-9).
 * Outputs: sdf [n], feat [n, ld_feat], normal [n,3], xyz_feature [n, ld_xyz >= 1386] (may be NULL).
 * Backward: cotangents d_sdf, d_feat, d_normal, d_xyz_feature (any of d_sdf/d_feat/d_xyz may be
 * NULL); writes d_pts [n,3] (may be NULL), ACCUMULATES d_bt_inv [frames,21,4,4] and d_T_pose
 * [frames,21,3] (may be NULL; caller zeroes) and dW/db.
 * ------------------------------------------------------------------------------------------- */
HN_API int64_t hn_sdf_hand_stash_floats(int64_t n_pts);
HN_API int64_t hn_sdf_hand_ws_floats(int64_t n_pts, int ws_kind);
HN_API int hn_sdf_hand_sdf(const hn_mlp_t* mlp, const float* pts, const float* bt_inv,
                           const float* T_pose, int64_t n_pts, int64_t pts_per_frame, float* sdf,
                           float* ws, int64_t ws_floats, int precision, hn_stream_t stream);
HN_API int hn_sdf_hand_fwd(const hn_mlp_t* mlp, const float* pts, const float* bt_inv,
                           const float* T_pose, int64_t n_pts, int64_t pts_per_frame, float* sdf,
                           float* feat, int64_t ld_feat, float* normal, float* xyz_feature,
                           int64_t ld_xyz, float* stash, int64_t stash_floats, int precision,
                           hn_stream_t stream);
/* Same call for a forward that will never be differentiated (rendering under no_grad): the stash is still the scratch of the
 * call (same size), but what only hn_sdf_hand_bwd would read is not written (HN_TC_MIXED16: the D16 tiles of the normal sweep). */
HN_API int hn_sdf_hand_fwd_render(const hn_mlp_t* mlp, const float* pts, const float* bt_inv,
                           const float* T_pose, int64_t n_pts, int64_t pts_per_frame, float* sdf,
                           float* feat, int64_t ld_feat, float* normal, float* xyz_feature,
                           int64_t ld_xyz, float* stash, int64_t stash_floats, int precision,
                           hn_stream_t stream);
HN_API int hn_sdf_hand_bwd(const hn_mlp_t* mlp, const float* pts, const float* bt_inv,
                           const float* T_pose, int64_t n_pts, int64_t pts_per_frame, float* stash,
                           const float* d_sdf, const float* d_feat, int64_t ld_dfeat,
                           const float* d_normal, const float* d_xyz_feature, int64_t ld_dxyz,
                           float* d_pts, float* d_bt_inv, float* d_T_pose, const hn_mlp_grad_t* grad,
                           float* ws, int64_t ws_floats, int precision, hn_stream_t stream);

/* Hand colour field: RenderingNetwork.forward (utils/fields.py:222-240), input
 * cat[xyz_feature (1386), feature (256), normal + enc4 (27)] = 1669 -> 256 x4 -> 3, sigmoid.
 * mlp layer 0 must be packed with hn_wn_pack_gap(gap_at = 1386, gap = 2) (in_dim 1671). */
HN_API int64_t hn_color_hand_stash_floats(int64_t n_pts);
HN_API int64_t hn_color_hand_ws_floats(int64_t n_pts, int ws_kind);
HN_API int hn_color_hand_fwd(const hn_mlp_t* mlp, const float* xyz_feature, int64_t ld_xyz,
                             const float* feat, int64_t ld_feat, const float* normal, int64_t n_pts,
                             float* rgb, float* stash, int64_t stash_floats, int precision,
                             hn_stream_t stream);
/* Forward that will never be differentiated (rendering): with the operands of hn_color_hand_chain_pack in mlp->chain and
 * HN_TC_MIXED16, layers 1..3 and the output layer run on the colour chain kernel (csrc/chain_color.cu) and nothing is stashed for
 * a backward; the stash is still the scratch of the call (same size). */
HN_API int hn_color_hand_fwd_render(const hn_mlp_t* mlp, const float* xyz_feature, int64_t ld_xyz,
                             const float* feat, int64_t ld_feat, const float* normal, int64_t n_pts,
                             float* rgb, float* stash, int64_t stash_floats, int precision,
                             hn_stream_t stream);
/* per-layer operands of hn_mlp_bx3_pack followed by the chain operands of layers 1..4 */
HN_API int64_t hn_color_hand_chain_bytes(const hn_mlp_t* m);
HN_API int hn_color_hand_chain_pack(const hn_mlp_t* m, void* buf, int64_t bytes, hn_stream_t stream);
HN_API int hn_color_hand_bwd(const hn_mlp_t* mlp, int64_t n_pts, float* stash, const float* rgb,
                             const float* d_rgb, float* d_xyz_feature, int64_t ld_dxyz, float* d_feat,
                             int64_t ld_dfeat, float* d_normal, const hn_mlp_grad_t* grad, float* ws,
                             int64_t ws_floats, int precision, hn_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * Ray helpers and hierarchical sampling (utils/renderer.py:10-37, 60-105, 119-127, 204-234).
 * ------------------------------------------------------------------------------------------- */
/* pts[b,i,:] = o[b,:] + d[b,:] * z[b,i]  with a separately rounded multiply and add, bit-exact
 * with eager PyTorch (utils/renderer.py:91,216). */
HN_API int hn_ray_points(const float* rays_o, const float* rays_d, const float* z, int64_t n_rays,
                         int n, float* pts, hn_stream_t stream);
/* dists / mid-point samples of render_core (utils/renderer.py:119-127):
 * dists[b,i] = z[b,i+1]-z[b,i] (last = sample_dist), mid = z + dists*0.5, pts = o + d*mid; when `dirs` is not NULL
 * also dirs[b,i,:] = d[b,:] (the expanded view directions the colour field reads). */
HN_API int hn_mid_points(const float* rays_o, const float* rays_d, const float* z, int64_t n_rays, int n,
                         float sample_dist, float* pts, float* dists, float* dirs, hn_stream_t stream);
/* Its backward: d_rays_o[b] = sum_i d_pts[b,i], d_rays_d[b] = sum_i (d_pts[b,i] * mid[b,i] + d_dirs[b,i])
 * (d_dirs may be NULL); one warp per ray. */
HN_API int hn_mid_points_bwd(const float* d_pts, const float* d_dirs, const float* z, const float* dists,
                             int64_t n_rays, int n, float* d_rays_o, float* d_rays_d, hn_stream_t stream);
/* Rays into the object frame (convert_obj_to_local, utils/renderer.py:180-188): o' = Ro (o - To), d' = Ro d with
 * Ro [3,3] row-major and To [3] in DEVICE memory (trained pose parameters). */
HN_API int hn_rays_to_local(const float* rays_o, const float* rays_d, const float* Ro, const float* To,
                            int64_t n_rays, float* local_o, float* local_d, hn_stream_t stream);
/* Backward in one launch (one CTA, deterministic): d_Ro [9], d_To [3] overwritten; d_local_o / d_local_d may be NULL
 * (= zero); d_rays_o / d_rays_d [n_rays,3] optional (NULL when the rays are inputs). */
HN_API int hn_rays_to_local_bwd(const float* d_local_o, const float* d_local_d, const float* rays_o,
                                const float* rays_d, const float* Ro, const float* To, int64_t n_rays,
                                float* d_Ro, float* d_To, float* d_rays_o, float* d_rays_d, hn_stream_t stream);
/* NeuSRenderer.up_sample (utils/renderer.py:60-86): one warp per ray; section weights, fp64
 * running cdf (matching torch CPU cumsum/cumprod), inverse-CDF search. */
HN_API int hn_up_sample(const float* z, const float* sdf, const float* u, int64_t n_rays, int m, int n_importance,
                        float inv_s, float* new_z, hn_stream_t stream);
/* The inverse-CDF step of sample_pdf (utils/renderer.py:18-35) given an explicit cdf [B,m]:
 * u = linspace(.5/n, 1-.5/n, n) supplied by the caller (host-generated with torch to stay
 * bit-identical); searchsorted(right=True); below/above (int64, may be NULL). */
HN_API int hn_inverse_cdf(const float* bins, const float* cdf, const float* u, int64_t n_rays,
                          int m, int n_samples, float* samples, int64_t* below, int64_t* above,
                          hn_stream_t stream);
/* cat_z_vals (utils/renderer.py:88-105): stable merge of two sorted rows z_a [B,m], z_b [B,k]
 * -> z_out [B,m+k], index (position in cat[z_a,z_b]; int64, may be NULL); when sdf_a/sdf_b are
 * given, sdf_out[b,i] = cat[sdf_a,sdf_b][src_row(b), index[b,i]] with src_row(b) = b, or
 * b % sdf_row_mod when sdf_row_mod > 0 (the frame-0 gather quirk of
 * utils/renderer_batch.py:108-111, SURVEY D-7). */
HN_API int hn_merge_sorted(const float* z_a, int m, const float* z_b, int k, int64_t n_rays,
                           float* z_out, int64_t* index, const float* sdf_a, const float* sdf_b,
                           int64_t sdf_row_mod, float* sdf_out, hn_stream_t stream);
/* Stable ascending sort of every row of x [B,n] (n <= 1024): the final torch.sort of
 * NeuSRenderer_fitting.render (utils/renderer.py:498). index may be NULL. */
HN_API int hn_sort_rows(const float* x, int64_t n_rays, int n, float* out, int64_t* index,
                        hn_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * s-density alpha + transmittance + compositing (utils/renderer.py:144-169), one warp per ray.
 * variance: device pointer to SingleVarianceNetwork.variance; inv_s = clip(exp(10 v), 1e-6, 1e6).
 * seed_with_c0 != 0 reproduces render_core's cumprod seeded with c[:, :1] (SURVEY D-1).
 * Outputs: weights [B,n], cdf c [B,n], alpha [B,n] (may be NULL), color [B,3], weight_sum [B],
 * weight_max [B], eik [B] = per-ray sum of (||normal||-1)^2.
 * ------------------------------------------------------------------------------------------- */
HN_API int hn_neus_composite_fwd(const float* sdf, const float* normal, const float* rgb,
                                 const float* dists, const float* rays_d, const float* variance,
                                 int64_t n_rays, int n, int seed_with_c0, float* weights,
                                 float* cdf, float* alpha, float* color, float* weight_sum,
                                 float* weight_max, float* eik, hn_stream_t stream);
/* Backward (SURVEY E-5).  Cotangents: d_color [B,3], d_weight_sum [B] (may be NULL),
 * d_weights [B,n] (may be NULL), d_eik [B] (may be NULL).  Outputs (overwritten): d_sdf [B*n],
 * d_normal [B*n,3], d_rgb [B*n,3], d_rays_d [B,3] (may be NULL); d_variance (1 float) is
 * ACCUMULATED with atomics (caller zeroes). */
HN_API int hn_neus_composite_bwd(const float* sdf, const float* normal, const float* rgb,
                                 const float* dists, const float* rays_d, const float* variance,
                                 const float* weights, int64_t n_rays, int n, int seed_with_c0,
                                 const float* d_color, const float* d_weight_sum,
                                 const float* d_weights, const float* d_eik, float* d_sdf,
                                 float* d_normal, float* d_rgb, float* d_rays_d,
                                 float* d_variance, hn_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * Two-field (hand + object) fitting compositor, one warp per ray, n <= 256.
 * hn_neus_alpha_*: NeuSRenderer_fitting.get_alpha_sample_color's alpha and eikonal sums
 *   (utils/renderer.py:396-420; utils/renderer_batch.py:150-172): alpha [B,n], eik [B] = per-ray
 *   sums of (||normal||-1)^2.  Backward overwrites d_sdf [B*n], d_normal [B*n,3], d_rays_d [B,3]
 *   (may be NULL) and ACCUMULATES d_variance (1 float, caller zeroes); d_alpha / d_eik may be NULL.
 * hn_fit_composite_*: T_i = prod_{j<i} (1-a_h+1e-7)(1-a_o+1e-7) with a leading ONE (SURVEY D-1),
 *   color = sum a_h T rgb_h + sum a_o T rgb_o, weight_sum = sum a_h T + sum a_o T
 *   (utils/renderer.py:512-524; utils/renderer_batch.py:258-270).  trans [B,n] is the stash the
 *   backward reads.  d_color [B,3] / d_weight_sum [B] may be NULL.
 * ------------------------------------------------------------------------------------------- */
HN_API int hn_neus_alpha_fwd(const float* sdf, const float* normal, const float* dists,
                             const float* rays_d, const float* variance, int64_t n_rays, int n,
                             float* alpha, float* eik, hn_stream_t stream);
HN_API int hn_neus_alpha_bwd(const float* sdf, const float* normal, const float* dists,
                             const float* rays_d, const float* variance, int64_t n_rays, int n,
                             const float* d_alpha, const float* d_eik, float* d_sdf,
                             float* d_normal, float* d_rays_d, float* d_variance,
                             hn_stream_t stream);
HN_API int hn_fit_composite_fwd(const float* alpha_h, const float* rgb_h, const float* alpha_o,
                                const float* rgb_o, int64_t n_rays, int n, float* trans,
                                float* color, float* weight_sum, hn_stream_t stream);
HN_API int hn_fit_composite_bwd(const float* alpha_h, const float* rgb_h, const float* alpha_o,
                                const float* rgb_o, const float* trans, int64_t n_rays, int n,
                                const float* d_color, const float* d_weight_sum, float* d_alpha_h,
                                float* d_rgb_h, float* d_alpha_o, float* d_rgb_o,
                                hn_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * Ray generation: utils/utils.py:31-115 (_xy_to_ray_bundle) on top of pytorch3d's
 * PerspectiveCameras.unproject_points(from_ndc=True).  A camera is 16 device floats:
 * R[9] (row-major, pytorch3d row-vector convention X_view = X_world R + T) | T[3] | fx fy | px py (NDC).
 * origins = P(depth 1) - dir,  dir = normalize(P(depth 2) - P(depth 1)).
 * ------------------------------------------------------------------------------------------- */
/* xy [n_cams, n_per_cam, 2] NDC coordinates, cams [n_cams, 16] -> rays_o, rays_d [n_cams * n_per_cam, 3]. */
HN_API int hn_rays_from_ndc(const float* xy, const float* cams, int64_t n_cams, int64_t n_per_cam,
                            float* rays_o, float* rays_d, hn_stream_t stream);
/* Full-image form (exp_runner.py:338-350): pixel p = row * W + col has NDC (xs[col], ys[row]); xs [W], ys [H] are
 * the reference's linspace values (made by torch on the host so they are its values bit for bit).  Writes the
 * `count` rays of pixels [first, first + count): a chunk of rays_o.split(batch_size) without the full-image list. */
HN_API int hn_rays_ndc_grid(const float* xs, const float* ys, int W, int H, const float* cam, int64_t first,
                            int64_t count, float* rays_o, float* rays_d, hn_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * Loss epilogues: one forward launch (deterministic two-level sums) + one elementwise
 * backward launch writing the cotangents of the render outputs.  `ws`: hn_loss_ws_floats() floats, ZEROED once
 * by the caller before its first use (the kernels leave it clean); `out`: 8 floats.
 * ------------------------------------------------------------------------------------------- */
HN_API int64_t hn_loss_ws_floats(void);
/* Render loss of exp_runner.py:206-227 (training: color_div <= 0 -> divide by mask_sum + 1e-5),
 * fitting_single.py:253-256 (color_div = n_rays) and fitting_video.py:287-291 (color_div = F * P):
 *   color_loss = sum |(color - true_rgb) * mask| / div,  mask_loss = mean BCE(clip(weight_sum, 1e-3, 1 - 1e-3), mask),
 *   total = color_weight * color_loss + mask_weight * mask_loss + igr_weight * (*gradient_error, may be NULL).
 * color, true_rgb [n,3]; weight_sum, true_mask [n] (mask already thresholded to 0/1).
 * color_div_dev (may be NULL): a DEVICE scalar that overrides color_div -- a ray shard of a larger batch passes the
 * whole batch's mask_sum + 1e-5 (and color/mask/igr weights scaled by its share of the rays), so that the shards'
 * totals add up to the batch loss without a host round trip.
 * out: [0] total, [1] color_loss, [2] mask_loss, [3] psnr (exp_runner.py:222), [4] divisor, [5] mask_sum + 1e-5,
 * [6] eikonal term. */
HN_API int hn_render_loss_fwd(const float* color, const float* weight_sum, const float* true_rgb,
                              const float* true_mask, const float* gradient_error, int64_t n_rays,
                              float color_div, const float* color_div_dev, float color_weight,
                              float mask_weight, float igr_weight, float* ws, float* out, hn_stream_t stream);
/* g_loss: device scalar (NULL = 1).  fwd_out: the forward's `out`.  d_gradient_error may be NULL. */
HN_API int hn_render_loss_bwd(const float* g_loss, const float* color, const float* weight_sum,
                              const float* true_rgb, const float* true_mask, const float* fwd_out,
                              int64_t n_rays, float color_weight, float mask_weight, float igr_weight,
                              float* d_color, float* d_weight_sum, float* d_gradient_error,
                              hn_stream_t stream);
/* Contact / penetration terms of fitting_single.py:268-283 and fitting_video.py:295-309 on column 0 of the
 * per-sample SDFs (element i at sdf[i * ld]):
 *   contact = mean over {|h| + |o| < contact_thr} of (|h| + |o|),  penet = mean over {h < 0, o < 0} of (|h| + |o|),
 *   counts carry the reference's + 1e-9.  out: [0] w_contact * contact + w_penet * penet, [1] contact, [2] penet,
 *   [3] contact count, [4] penetration count.  d_sdf_hand / d_sdf_obj: dense [n_pts]. */
HN_API int hn_interaction_loss_fwd(const float* sdf_hand, int64_t ld_hand, const float* sdf_obj, int64_t ld_obj,
                                   int64_t n_pts, float contact_thr, float w_contact, float w_penet, float* ws,
                                   float* out, hn_stream_t stream);
HN_API int hn_interaction_loss_bwd(const float* g_loss, const float* sdf_hand, int64_t ld_hand,
                                   const float* sdf_obj, int64_t ld_obj, const float* fwd_out, int64_t n_pts,
                                   float contact_thr, float w_contact, float w_penet, float* d_sdf_hand,
                                   float* d_sdf_obj, hn_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * Nearest-neighbour selection of get_stable_loss_cross (utils/renderer_batch.py:346-357, where
 * the reference builds a scipy cKDTree per frame on the host).  pts [n_pts,3] (shared by all frames); in_mask,
 * out_mask, flag: [n_frames, n_pts] bytes.  For each frame t and each p with in_mask[t,p]: q* = argmin over
 * {q : out_mask[t,q]} of |pts[q] - pts[p]|^2 (fp64 distances like cKDTree, exact ties -> lowest q); sets
 * flag[t,q*] = 1 (flag must be zeroed by the caller; = np.unique(near_out_id) as a flag array) and, if `nearest` is
 * not NULL, nearest[t,p] = q* (-1 for points outside in_mask or when the frame has no candidate).
 * ------------------------------------------------------------------------------------------- */
HN_API int hn_nn_select(const float* pts, const uint8_t* in_mask, const uint8_t* out_mask, int n_frames,
                        int n_pts, uint8_t* flag, int64_t* nearest, hn_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * Marching cubes on the device (replaces mcubes.marching_cubes(u, threshold) of utils/renderer.py:279,561 and
 * utils/renderer_batch.py:309; PyMCubes is an un-vendored dependency: parity unpinned).  u [nx,ny,nz] fp32 on the device.
 *   hn_mc_set_tables  case table (ho-nerf_b200/mcubes_tables.py: n_tris[256], tris[256][15] edge ids, owner[12][4]) into the
 *                     CURRENT device's constant memory (HOST pointers; once per device)
 *   hn_mc_classify    flags [3][nx][ny][nz] int32 (edge from the lattice point along +x/+y/+z crosses iso),
 *                     cell_tris [(nx-1)(ny-1)(nz-1)] int32 (triangles of the cell)
 *   hn_mc_emit        given the INCLUSIVE prefix sums of both arrays (caller: torch.cumsum), writes vertices [V,3] (index
 *                     coordinates, shared between cells) and triangles [T,3] int32 (normals towards lower values; the
 *                     reference reverses them afterwards)
 * ------------------------------------------------------------------------------------------- */
HN_API int hn_mc_set_tables(const int8_t* n_tris, const int8_t* tris, const int8_t* owner);
HN_API int hn_mc_classify(const float* u, int nx, int ny, int nz, float iso, int32_t* flags, int32_t* cell_tris,
                          hn_stream_t stream);
HN_API int hn_mc_emit(const float* u, int nx, int ny, int nz, float iso, const int32_t* flags, const int32_t* vscan,
                      const int32_t* cell_tris, const int32_t* tscan, float* vertices, int32_t* triangles,
                      hn_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * Diagnostics / test hooks of the product kernels (the tcgen05 bring-up self-test GEMMs live in a separate library:
 * include/honerf_b200_selftest.h, libhonerf_b200_selftest.so)
 * ------------------------------------------------------------------------------------------- */
/* The weight-gradient kernel of the HN_TC_BF16X3 path, for tests: C [out, ldc] += P^T Q (+ P2^T Q2) over n
 * points, P [n, out] and Q [n, in] fp32: row-major (x_tiled = 0, leading dimension ld), tiled (x_tiled = 1:
 * [tile][col/4][128][4], n padded to 128) or column-major tiles (x_tiled = 2: [tile][ld columns][128 rows]);
 * db [out] += column sums of P (may be NULL); part: workspace of >= 16 * 65536 floats. */
/* HN_TC_MIXED16 weight-gradient kernel on fp32 operands rounded to bf16 dW-ready tiles (diagnostics / tests):
 * C[out, in] = P^T Q (+ P2^T Q2), db = column sums of P.  `tiles`: 4 * round_up(n,128) * 512 bytes of workspace. */
HN_API int hn_dw16_test(const float* P, int out, const float* Q, int in, const float* P2, const float* Q2, int64_t n,
                        float* C, int64_t ldc, float* db, void* tiles, int64_t tiles_bytes, float* part,
                        int64_t part_floats, hn_stream_t stream);
HN_API int hn_dw16_set_debug(int swap_lbo_sbo);
/* diagnostics: progress words of the HN_TC_MIXED16 kernels in a host-mapped buffer of 4*4*148*8 uint32 (NULL: off) */
HN_API int hn_chain16_set_debug(void* host_mapped_words);
HN_API int hn_dw_test(const float* P, int64_t ldp, int p_tiled, int out, const float* Q, int64_t ldq,
                      int q_tiled, int in, const float* P2, const float* Q2, int64_t n, float* C,
                      int64_t ldc, float* db, float* part, int64_t part_floats, hn_stream_t stream);
/* ---- render_core_outside (background branch of a NeuS renderer) ---------------------------------------------------------
 * Named by the north star; the HO-NeRF reference only stores n_outside (utils/renderer.py:47,56; every config sets 0) and has
 * no such method: these follow the semantics of the NeuS renderer it was derived from (PARITY UNPINNED; oracle:
 * oracle/honerf_oracle.render_core_outside).
 * hn_outside_points: z_vals [n_rays, n] -> inverted-sphere query points pts4 [n_rays*n, 4] = (p / r, 1 / r), r = clip(|p|, 1,
 *   1e10), p = o + d (z + dists / 2); dirs [n_rays*n, 3]; dists [n_rays, n] (= diff(z), sample_dist appended).
 * hn_outside_composite_fwd: density [n_rays, n], raw_rgb [n_rays, n, 3] (the NeRF's outputs) -> sampled_color = sigmoid(raw),
 *   alpha = 1 - exp(-softplus(density) dists), weights = alpha * exclusive cumprod(1 - alpha + 1e-7), color [n_rays, 3] =
 *   sum w c (+ background[3] (1 - sum w) when background != NULL).
 * hn_outside_composite_bwd: cotangents of any of the four outputs (NULL = none) -> d_density, d_raw_rgb.  n <= 512. */
HN_API int hn_outside_points(const float* rays_o, const float* rays_d, const float* z_vals, float sample_dist, int64_t n_rays,
                             int n, float* pts4, float* dirs, float* dists, hn_stream_t stream);
HN_API int hn_outside_composite_fwd(const float* density, const float* raw_rgb, const float* dists, const float* background,
                                    int64_t n_rays, int n, float* sampled_color, float* alpha, float* weights, float* color,
                                    hn_stream_t stream);
HN_API int hn_outside_composite_bwd(const float* density, const float* dists, const float* background,
                                    const float* sampled_color, const float* alpha, const float* weights, int64_t n_rays, int n,
                                    const float* g_color, const float* g_sampled_color, const float* g_alpha,
                                    const float* g_weights, float* d_density, float* d_raw_rgb, hn_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* HONERF_B200_H */
