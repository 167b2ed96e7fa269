"""The hand field and the two-field fitting renderers under the PRODUCT DEFAULT precision (tensor-core contractions),
against the same golden vectors / fp64 oracles and with the same bounds as under the fp32 SIMT verification path
(tests/test_gpu_hand.py, tests/test_gpu_fit.py pin `simt_fp32` through conftest.py; this module does not, so the
very same test bodies run on what fitting_single.py / fitting_video.py would run): values, normals and the gradients
with respect to bt_inv, T_pose_21, Ro, To and every weight."""
import pytest

import test_gpu_fit as TF
import test_gpu_hand as TH

pytestmark = pytest.mark.gpu


def _assert_default():
    import honerf_b200 as H
    assert H.ops.default_precision() != H.ops._PRECISIONS["simt_fp32"]


def test_hand_fields_vs_golden_default():
    _assert_default()
    TH.test_hand_fields_vs_golden()


def test_hand_second_order_backward_vs_fp64_default():
    """d bt_inv, d T_pose_21, d pts and all weight gradients of the hand SDF operator: 1e-2 relative (L2)."""
    _assert_default()
    TH.test_hand_embedding_jacobian_paths_vs_fp64_oracle()


def test_hand_color_backward_default():
    _assert_default()
    TH.test_hand_color_backward_vs_oracle()


def test_hand_render_vs_golden_default():
    _assert_default()
    TH.test_hand_render_vs_golden()


def test_hand_render_core_given_same_z_default():
    """Same-z hand render_core under the tensor-core default: colour 3e-3 / weights 1e-4 as under SIMT; every gradient
    (weights, variance, bt_inv, T_pose_21) within 2.5e-2 -- measured on the B200: 2.0-2.2e-2 where the reference's own fp32
    arithmetic is 0.5-1.4e-2 off fp64 (ill-conditioned synthetic case, see _hand_same_z: the per-layer bf16 hi+lo
    contractions carry 16 mantissa bits, the normal of magnitude ~70 enters the colour net as sin / cos(8 n), and a flipped
    ReLU gate moves the summed weight gradients); the fp32 SIMT path holds 1e-2 on the same case and the object field's
    same-z test holds the plain 1e-2 under the default."""
    _assert_default()
    TH._hand_same_z(2.0, floor=2.5e-2)


def test_fit_render_vs_golden_default():
    _assert_default()
    TF.test_fit_render_vs_golden()


def test_fit_render_given_same_z_default():
    """NeuSRenderer_fitting (utils/renderer.py:434-535) on the oracle's z_vals: colour, weight sums, both SDFs, and the
    pose gradients d bt_inv / d Ro / d To vs fp64."""
    _assert_default()
    TF._same_z(False)


def test_fit_render_batch_given_same_z_default():
    """renderer_batch.NeuSRenderer_fitting (utils/renderer_batch.py:184-281), frame-batched: pose gradients within
    max(1.5e-2, 2 x the reference's own fp32-vs-fp64 error) under the tensor-core default (measured 0.5-1.0e-2 where the
    reference's fp32 arithmetic is 0.4e-2 off; the SIMT path holds max(1e-2, 2 x own))."""
    _assert_default()
    TF._same_z(True, floor=1.5e-2)


def test_fit_render_batch_vs_golden_default():
    _assert_default()
    TF.test_fit_render_batch_vs_golden_including_frame0_gather_quirk()


def test_fit_render_given_same_z_frozen_nets_take_the_hand_chain():
    """The same comparison with the nets FROZEN (how fitting_single.py runs them): under the default precision the hand SDF net
    then goes through its tile-chain kernels (csrc/chain16_hand.cu) -- values and pose gradients vs fp64 with the same bounds."""
    _assert_default()
    TF._same_z(False, freeze=True)


def test_fit_render_batch_given_same_z_frozen_nets_take_the_hand_chain():
    """Frame-batched renderer, frozen nets: 2 frames x 960 points (a frame boundary inside a 128-point tile: the per-thread
    frame index and the non-uniform pose-gradient accumulation of halo_bwd_tiled_kernel)."""
    _assert_default()
    TF._same_z(True, floor=1.5e-2, freeze=True)
