"""The product as a user gets it (default precision = the fused tcgen05 chain kernels): end-to-end render against
the reference's golden vectors, the training-step gradients, the SDF lattice of extract_geometry, and the smoke entry."""
import pytest
import torch

import cases
import honerf_oracle as O
from golden_util import check_grads_against_golden, load_golden, max_abs, rel_err
from gpu_util import DEV, obj_modules

pytestmark = pytest.mark.gpu


def test_default_precision_is_the_tensor_core_chain_path():
    import honerf_b200 as H
    assert H.ops.default_precision() == H.ops._PRECISIONS["tc_bf16x3"]


def test_render_vs_golden_with_default_precision():
    """NeuSRenderer.render (utils/renderer.py:190-258) through the chain kernels vs the reference: colour / weight sums
    1e-3 abs, loss 1e-3 relative, gradients 5e-2 relative (see below)."""
    import honerf_b200 as H
    import ref_conf
    from test_gpu_render import _fixed_rand
    g = load_golden("obj_render")
    c = cases.obj_render_case()
    R = c["R"]
    sdf, col, dev, _, _ = obj_modules()
    r = H.NeuSRenderer(sdf, dev, col, "obj", **ref_conf.RENDERER_CONF)
    Ro, To = R["Ro"].to(DEV).requires_grad_(True), R["To"].to(DEV).requires_grad_(True)
    with _fixed_rand(R["t_rand"]):
        out = r.render(R["rays_o"].to(DEV), R["rays_d"].to(DEV), R["near"], R["far"], None, None, None, Ro, To, 0)
    for k in ("color_fine", "s_val", "weight_sum", "weight_max"):
        assert max_abs(out[k], g[k]) < 1e-3, k
    assert rel_err(out["gradient_error"], g["gradient_error"]) < 1e-3
    loss = O.training_loss(out, c["true_rgb"].to(DEV), c["true_mask"].to(DEV))
    assert rel_err(loss, g["loss"]) < 1e-3
    loss.backward()
    # end to end the importance samples themselves move when a coarse SDF value differs by 1e-5 (the sampler's
    # sigmoid has inv_s up to 512), which shifts some gradients by ~1 %: same 5e-2 bound as the SIMT end-to-end test;
    # the strict 1e-2 check on identical z_vals is test_gpu_obj_tc.py::test_render_core_given_same_z_tc[tc_bf16x3]
    grads = {"sdf." + k: p.grad.cpu() for k, p in sdf.named_parameters() if p.grad is not None}
    grads.update({"color." + k: p.grad.cpu() for k, p in col.named_parameters() if p.grad is not None})
    grads.update({"variance": dev.variance.grad.cpu(), "Ro": Ro.grad.cpu(), "To": To.grad.cpu()})
    check_grads_against_golden(g, grads, 5e-2)


def test_sdf_grid_vs_golden_with_default_precision():
    """The `u` lattice of extract_geometry (utils/renderer.py:262-278) through the SDF-only chain kernel."""
    import honerf_b200 as H
    import ref_conf
    g = load_golden("sdf_grid")
    c = cases.sdf_grid_case()
    sdf, col, dev, _, _ = obj_modules(requires_grad=False)
    r = H.NeuSRenderer(sdf, dev, col, "obj", **ref_conf.RENDERER_CONF)
    lo, hi = torch.full((3,), c["lo"]), torch.full((3,), c["hi"])
    u = r.sdf_grid(lo, hi, c["res"])
    assert u.shape == (c["res"],) * 3 and max_abs(u, g["u"]) < 1e-4


def test_smoke_entry_point():
    import __graft_entry__ as E
    E.smoke()


def test_hand_field_and_fitting_renderer_with_default_precision():
    """The hand nets have no chain kernels yet: under the default precision their contractions run on the per-layer
    tcgen05 kernels with split TF32 operands (HN_TC_BF16X3 -> HN_TC_TF32X3 in gemm_dispatch.cuh).  Hand SDF / colour vs
    the reference's golden vectors (relative bounds, see test_gpu_hand.py) and one two-field fitting render + backward."""
    import honerf_b200 as H
    import ref_conf
    from golden_util import rel_l2
    from gpu_util import hand_modules
    g = load_golden("hand_fields")
    c = cases.hand_fields_case()
    hs, hc, hd, _, _ = hand_modules()
    pts, bt, T = c["pts"].to(DEV), c["bt_inv"].to(DEV), c["T_pose_21"].to(DEV)
    out, xyz, _, _ = hs(pts, bt, T)
    n = hs.gradient(pts, bt, T).squeeze(1)
    assert max_abs(xyz, g["xyz_feature"]) < 2e-4 and max_abs(out, g["sdf_out"]) < 1e-3
    assert rel_l2(n, g["gradient"]) < 1e-2
    assert max_abs(hc(None, xyz, out[:, 1:], None, n, 0), g["rgb"]) < 2e-3
    # two-field renderer: hand (per-layer TF32x3) + object (chain kernels) in one render, gradients to the poses
    os_, oc, od, _, _ = obj_modules()
    r = H.renderer.NeuSRenderer_fitting(hs, hd, hc, os_, od, oc, **ref_conf.RENDERER_CONF)
    fc = cases.fit_render_case()
    R = fc["R"]
    btg = fc["bt_inv"].to(DEV).requires_grad_(True)
    Ro, To = fc["Ro"].to(DEV).requires_grad_(True), fc["To"].to(DEV).requires_grad_(True)
    o = r.render(R["rays_o"].to(DEV), R["rays_d"].to(DEV), R["near"], R["far"], btg, fc["T_pose_21"].to(DEV), None, Ro, To)
    cases.fit_loss(o, fc["true_rgb"].to(DEV)).backward()
    assert torch.isfinite(o["color_fine"]).all() and o["color_fine"].shape == (10, 3)
    assert torch.isfinite(btg.grad).all() and torch.isfinite(Ro.grad).all() and float(btg.grad.abs().max()) > 0
