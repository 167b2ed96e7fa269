"""GPU parity of the tensor-core paths at the north-star tolerances (max-abs <= 1e-3 on colour / SDF, <= 1e-2
relative on normals and gradients): 'tc_mixed16' (the product default: fused tile-chain kernels with the 16-bit
activation stash), 'tc_bf16x3' (fused tile-chain kernels, fp32 stash) and 'tc_tf32x3' (per-layer tcgen05 GEMMs, split
TF32 operands); 'tc_tf32' (single pass) is the reduced-accuracy mode and is only held to 1e-2."""
import pytest
import torch

import cases
import honerf_oracle as O
import synth
from golden_util import load_golden, max_abs, rel_err, rel_l2
from gpu_util import DEV, obj_modules

pytestmark = pytest.mark.gpu


@pytest.fixture(params=["tc_mixed16", "tc_bf16x3", "tc_tf32x3", "tc_tf32"])
def tc_precision(request):
    import honerf_b200 as H
    H.set_default_precision(request.param)
    yield request.param
    H.set_default_precision("simt_fp32")


def test_fields_vs_golden_tc(tc_precision):
    g = load_golden("obj_fields")
    c = cases.obj_fields_case()
    sdf, col, dev, _, _ = obj_modules()
    pts, dirs = c["pts"].to(DEV), c["dirs"].to(DEV)
    s, f, n = sdf.fused(pts)
    # 'tc_tf32' (single pass everywhere) is the fast, reduced-accuracy mode: it is only required to
    # stay within 1e-2; 'tc_tf32x3' must meet the north-star bounds
    strict = tc_precision != "tc_tf32"
    tol = 1e-4 if strict else 1e-2
    print(tc_precision, "sdf err %.2e feat err %.2e normal rel %.2e" % (
        max_abs(s, g["sdf_out"][:, :1]), max_abs(f, g["sdf_out"][:, 1:]), rel_err(n, g["gradient"])))
    assert max_abs(s, g["sdf_out"][:, :1]) < tol and max_abs(f, g["sdf_out"][:, 1:]) < tol
    assert rel_err(n, g["gradient"]) < (1e-3 if strict else 1e-2)
    rgb = col(pts, dirs, f, n, 0)
    print(tc_precision, "colour err %.2e" % max_abs(rgb, g["rgb"]))
    assert max_abs(rgb, g["rgb"]) < (1e-3 if strict else 1e-2)
    assert max_abs(sdf.sdf(pts.detach()), g["sdf_out"][:, :1]) < tol


def test_second_order_backward_tc(tc_precision):
    sdf, col, dev, sp, cp = obj_modules()
    n = 1500
    gen = torch.Generator().manual_seed(11)
    pts = 0.45 * torch.randn(n, 3, generator=gen)
    d_sdf = torch.randn(n, 1, generator=gen)
    d_feat = 0.1 * torch.randn(n, 256, generator=gen)
    d_n = torch.randn(n, 3, generator=gen)
    spd = {k: v.double().requires_grad_(True) for k, v in sp.items() if k != "se3_refine"}
    x = pts.double().requires_grad_(True)
    out = O.sdf_obj_forward(spd, x)
    nrm = O.sdf_gradient(lambda q: O.sdf_obj_forward(spd, q)[:, :1], x)
    L = (out[:, :1] * d_sdf.double()).sum() + (out[:, 1:] * d_feat.double()).sum() + (nrm * d_n.double()).sum()
    names = list(spd)
    ref = dict(zip(["pts"] + names, torch.autograd.grad(L, [x] + [spd[k] for k in names])))
    xg = pts.to(DEV).requires_grad_(True)
    s, f, nn = sdf.fused(xg)
    ((s * d_sdf.to(DEV)).sum() + (f * d_feat.to(DEV)).sum() + (nn * d_n.to(DEV)).sum()).backward()
    worst = {"pts": rel_err(xg.grad, ref["pts"])}
    for k, p in sdf.named_parameters():
        if p.grad is not None:
            worst[k] = rel_err(p.grad, ref[k])
    print("worst relative gradient errors:", sorted(worst.items(), key=lambda kv: -kv[1])[:4])
    bad = {k: v for k, v in worst.items() if not v < 1e-2}
    assert not bad, bad


def test_render_core_given_same_z_tc(tc_precision):
    if tc_precision == "tc_tf32":
        pytest.skip("north-star tolerances are claimed for the split-operand modes only")
    import honerf_b200 as H
    import ref_conf
    c = cases.obj_render_case()
    R = c["R"]
    sdf, col, dev, sp, cp = obj_modules()
    r = H.NeuSRenderer(sdf, dev, col, "obj", **ref_conf.RENDERER_CONF)
    zref = O.render_obj(sp, cp, torch.tensor(0.3), R["rays_o"], R["rays_d"], R["near"], R["far"], R["Ro"],
                        R["To"], R["t_rand"])["z_vals"]
    from gpu_util import oracle_core_fp64
    rcore, ref_loss, ref_g, names = oracle_core_fp64(c, zref)
    ref = {"z_vals": zref, "color_fine": rcore["color"], "weights": rcore["weights"], "cdf_fine": rcore["cdf"]}
    Ro = R["Ro"].to(DEV).requires_grad_(True)
    To = R["To"].to(DEV).requires_grad_(True)
    lo, ld = r.convert_obj_to_local(R["rays_o"].to(DEV), R["rays_d"].to(DEV), Ro, To)
    r.index = 0
    core = r.render_core(lo, ld, None, None, None, ref["z_vals"].to(DEV), 1.1 / 64, sdf, dev, col)
    out = {"color_fine": core["color"], "weight_sum": core["weights"].sum(-1, keepdim=True),
           "gradient_error": core["gradient_error"]}
    assert max_abs(out["color_fine"], ref["color_fine"]) < 1e-3
    assert max_abs(core["weights"], ref["weights"]) < 1e-3
    loss = O.training_loss(out, c["true_rgb"].to(DEV), c["true_mask"].to(DEV))
    assert rel_err(loss, ref_loss) < 1e-3
    loss.backward()
    got = {"sdf." + k: p.grad for k, p in sdf.named_parameters() if p.grad is not None}
    got.update({"color." + k: p.grad for k, p in col.named_parameters() if p.grad is not None})
    got.update({"variance": dev.variance.grad, "Ro": Ro.grad, "To": To.grad})
    worst = {k: rel_l2(got[k], ref_g[k]) for k in names}
    worst_max = {k: rel_err(got[k], ref_g[k]) for k in names}
    print("worst max-norm relative gradient errors:", sorted(worst_max.items(), key=lambda kv: -kv[1])[:3])
    assert max(worst_max.values()) < 5e-2
    print("worst relative gradient errors:", sorted(worst.items(), key=lambda kv: -kv[1])[:4])
    bad = {k: v for k, v in worst.items() if not v < 1e-2}
    assert not bad, bad
