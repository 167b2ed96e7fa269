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
    assert H.ops.default_precision() == H.ops._PRECISIONS["tc_mixed16"]


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
    # the strict 1e-2 check on identical z_vals is test_gpu_obj_tc.py::test_render_core_given_same_z_tc[tc_mixed16] and tests/test_gpu_bench_shape.py
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


def test_ray_streams_render_equals_single_stream():
    """NeuSRenderer.ray_streams = 2 (two ray shards on concurrent streams, one shared parameter edge per net) against
    the single-stream render of the same rays: rendered outputs identical to 1e-6 (per-ray arithmetic does not depend on
    the batch a ray is in; the eikonal mean is recombined from shard means: 1e-6 relative), loss equal to 1e-6 relative, every gradient to 1e-4 relative (summation order of the weight
    gradients differs); also under torch.no_grad, with an odd shard split, and repeated (stash reuse across streams)."""
    import honerf_b200 as H
    import ref_conf
    import synth
    from golden_util import max_abs, rel_err
    from gpu_util import DEV, obj_modules
    sdf, col, var, _, _ = obj_modules()
    r = H.NeuSRenderer(sdf, var, col, "obj", **dict(ref_conf.RENDERER_CONF, perturb=0.0))
    R = synth.object_rays(301, seed=77)
    gen = torch.Generator().manual_seed(3)
    true_rgb = torch.rand(301, 3, generator=gen).to(DEV)
    true_mask = (torch.rand(301, 1, generator=gen) > 0.5).float().to(DEV)
    Ro, To = R["Ro"].to(DEV).requires_grad_(True), R["To"].to(DEV).requires_grad_(True)
    params = [p for m in (sdf, col, var) for n, p in m.named_parameters() if n != "se3_refine"] + [Ro, To]

    def run(k):
        r.ray_streams = k
        out = r.render(R["rays_o"].to(DEV), R["rays_d"].to(DEV), 0.4, 1.5, None, None, None, Ro, To, 0)
        loss = H.losses.training_loss(out, true_rgb, true_mask)
        grads = torch.autograd.grad(loss, params, allow_unused=True)
        torch.cuda.synchronize()
        return out, loss.detach(), grads
    out1, l1, g1 = run(1)
    for _ in range(2):
        out2, l2, g2 = run(2)
        assert set(out1) == set(out2)
        for k in out1:
            assert out1[k].shape == out2[k].shape, k
            if k == "gradient_error":       # a mean over all samples: shard means recombined, 1e-6 relative
                assert rel_err(out2[k], out1[k]) < 1e-6
            else:
                assert max_abs(out1[k], out2[k]) < 1e-6, k
        assert rel_err(l2, l1) < 1e-6
        for a, b in zip(g2, g1):
            assert (a is None) == (b is None)
            if a is not None:
                assert rel_err(a, b) < 1e-4
    out3, l3, g3 = run(3)
    assert max_abs(out3["color_fine"], out1["color_fine"]) < 1e-6 and rel_err(g3[0], g1[0]) < 1e-4
    with torch.no_grad():
        r.ray_streams = 2
        o = r.render(R["rays_o"].to(DEV), R["rays_d"].to(DEV), 0.4, 1.5, None, None, None, Ro, To, 0)
        torch.cuda.synchronize()
    assert max_abs(o["color_fine"], out1["color_fine"]) < 1e-6 and not o["color_fine"].requires_grad
    assert sdf.packed().shared_token is None


def test_render_sharded_per_shard_loss_adds_up_to_the_batch_loss():
    """NeuSRenderer.render_sharded with the fused loss evaluated per ray shard (batch-wide mask_sum + 1e-5 passed as a
    device scalar, BCE / eikonal weights = the shard's share of the rays): the shard totals add up to the loss of the
    reference's lines on the merged render (1e-6 relative) and give the same gradients (1e-4 relative)."""
    import honerf_b200 as H
    import ref_conf
    import synth
    from golden_util import rel_err
    from gpu_util import DEV, obj_modules
    sdf, col, var, _, _ = obj_modules()
    r = H.NeuSRenderer(sdf, var, col, "obj", **dict(ref_conf.RENDERER_CONF, perturb=0.0))
    n = 301
    R = synth.object_rays(n, seed=78)
    gen = torch.Generator().manual_seed(4)
    true_rgb = torch.rand(n, 3, generator=gen).to(DEV)
    true_mask = (torch.rand(n, 1, generator=gen) > 0.5).float().to(DEV)
    Ro, To = R["Ro"].to(DEV).requires_grad_(True), R["To"].to(DEV).requires_grad_(True)
    params = [p for m in (sdf, col, var) for k, p in m.named_parameters() if k != "se3_refine"] + [Ro, To]
    args = (R["rays_o"].to(DEV), R["rays_d"].to(DEV), 0.4, 1.5, None, None, None, Ro, To, 0)
    r.ray_streams = 1
    ref_loss = O.training_loss(r.render(*args), true_rgb, true_mask, igr_weight=0.3, mask_weight=0.7)
    ref_g = torch.autograd.grad(ref_loss, params, allow_unused=True)
    div = true_mask.sum() + 1e-5

    def shard_loss(out, lo, hi):
        w = (hi - lo) / float(n)
        return H.ops.render_loss(out["color_fine"], out["weight_sum"], true_rgb[lo:hi], true_mask[lo:hi],
                                 out["gradient_error"], div, 1.0, 0.7 * w, 0.3 * w)[0]
    for k in (1, 3):
        r.ray_streams = k
        parts = r.render_sharded(*args, shard_loss)
        assert len(parts) == k
        loss = torch.stack(parts).sum()
        g = torch.autograd.grad(loss, params, allow_unused=True)
        torch.cuda.synchronize()
        assert rel_err(loss, ref_loss) < 1e-6
        for a, b in zip(g, ref_g):
            assert (a is None) == (b is None)
            if a is not None:
                assert rel_err(a, b) < 1e-4
