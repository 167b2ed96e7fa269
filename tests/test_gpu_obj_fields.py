"""GPU parity: object SDF / colour fields (HN_SIMT_FP32 path) against the CPU oracle and the
reference's golden vectors.  Tolerances are stated per test."""
import pytest
import torch

import cases
import honerf_oracle as O
import synth
from golden_util import load_golden, max_abs, rel_err
from gpu_util import DEV, obj_modules

pytestmark = pytest.mark.gpu


def test_wn_pack_matches_weight_norm():
    from honerf_b200 import ops
    sdf, col, dev, sp, cp = obj_modules()
    pk = sdf.packed().get()
    for l in range(9):
        v, g = sp["lin%d.weight_v" % l], sp["lin%d.weight_g" % l]
        w = torch._weight_norm(v, g, 0) * (ops.SQRT1_2 if l == 4 else 1.0)
        o, i = v.shape
        ld = pk.lds[l]
        W = pk.W[pk.offsets[l]: pk.offsets[l] + o * ld].reshape(o, ld).cpu()
        assert rel_err(W[:, :i], w) < 1e-6
        assert float(W[:, i:].abs().sum()) == 0.0


def test_fields_vs_golden():
    """sdf/feature <= 2e-5 abs, normal <= 1e-4 rel, colour <= 2e-5 abs vs the reference's outputs."""
    g = load_golden("obj_fields")
    c = cases.obj_fields_case()
    sdf, col, dev, _, _ = obj_modules()
    pts, dirs = c["pts"].to(DEV), c["dirs"].to(DEV)
    out = sdf(pts)
    assert max_abs(out, g["sdf_out"]) < 2e-5
    n = sdf.gradient(pts)
    assert n.shape == (96, 1, 3)
    assert rel_err(n.squeeze(1), g["gradient"]) < 1e-4
    rgb = col(pts, dirs, out[:, 1:], n.squeeze(1), 0)
    assert max_abs(rgb, g["rgb"]) < 2e-5
    s_only = sdf.sdf(pts.detach())
    assert max_abs(s_only, g["sdf_out"][:, :1]) < 2e-5


@pytest.mark.parametrize("n", [1, 127, 1000, 4099])
def test_fields_vs_oracle_ragged_sizes(n):
    sdf, col, dev, sp, cp = obj_modules()
    gen = torch.Generator().manual_seed(n)
    pts = 0.5 * torch.randn(n, 3, generator=gen)
    dirs = torch.nn.functional.normalize(torch.randn(n, 3, generator=gen), dim=-1)
    ref = O.sdf_obj_forward(sp, pts)
    rn = O.sdf_gradient(lambda q: O.sdf_obj_forward(sp, q)[:, :1], pts.clone()).detach()
    s, f, nrm = sdf.fused(pts.to(DEV))
    assert max_abs(s, ref[:, :1]) < 2e-5 and max_abs(f, ref[:, 1:]) < 2e-5
    assert rel_err(nrm, rn) < 1e-4
    rgb = col(pts.to(DEV), dirs.to(DEV), f, nrm, 0)
    rr = O.color_obj_forward(cp, pts, dirs, ref[:, 1:].detach(), rn)
    assert max_abs(rgb, rr) < 3e-5
    assert max_abs(sdf.sdf(pts.to(DEV)), ref[:, :1]) < 2e-5


def test_empty_input():
    sdf, col, dev, _, _ = obj_modules()
    pts = torch.zeros(0, 3, device=DEV)
    s, f, n = sdf.fused(pts)
    assert s.shape == (0, 1) and f.shape == (0, 256) and n.shape == (0, 3)
    assert sdf.sdf(pts).shape == (0, 1)


def _param_grads(mod):
    return {k: p.grad.detach().cpu() for k, p in mod.named_parameters() if p.grad is not None}


def test_sdf_second_order_backward_vs_oracle():
    """Random cotangents on (sdf, feature, normal): every parameter gradient and d_pts within 2e-3
    relative (max-norm per tensor) of fp64 autograd through the oracle (double backward)."""
    sdf, col, dev, sp, cp = obj_modules()
    n = 777
    gen = torch.Generator().manual_seed(11)
    pts = 0.45 * torch.randn(n, 3, generator=gen)
    d_sdf = torch.randn(n, 1, generator=gen)
    d_feat = 0.1 * torch.randn(n, 256, generator=gen)
    d_n = torch.randn(n, 3, generator=gen)
    # oracle in fp64
    spd = {k: v.double().requires_grad_(True) for k, v in sp.items() if k != "se3_refine"}
    x = pts.double().requires_grad_(True)
    out = O.sdf_obj_forward(spd, x)
    nrm = O.sdf_gradient(lambda q: O.sdf_obj_forward(spd, q)[:, :1], x)
    L = (out[:, :1] * d_sdf.double()).sum() + (out[:, 1:] * d_feat.double()).sum() + (nrm * d_n.double()).sum()
    names = list(spd)
    ref = dict(zip(["pts"] + names, torch.autograd.grad(L, [x] + [spd[k] for k in names])))
    # CUDA
    xg = pts.to(DEV).requires_grad_(True)
    s, f, nn = sdf.fused(xg)
    Lg = (s * d_sdf.to(DEV)).sum() + (f * d_feat.to(DEV)).sum() + (nn * d_n.to(DEV)).sum()
    Lg.backward()
    assert rel_err(xg.grad, ref["pts"]) < 2e-3
    got = _param_grads(sdf)
    worst = {k: rel_err(got[k], ref[k]) for k in names}
    bad = {k: v for k, v in worst.items() if not v < 2e-3}
    assert not bad, bad


def test_color_backward_vs_oracle():
    sdf, col, dev, sp, cp = obj_modules()
    n = 515
    gen = torch.Generator().manual_seed(12)
    pts = 0.45 * torch.randn(n, 3, generator=gen)
    dirs = torch.nn.functional.normalize(torch.randn(n, 3, generator=gen), dim=-1)
    feat = torch.randn(n, 256, generator=gen)
    nrm = torch.randn(n, 3, generator=gen)
    d_rgb = torch.randn(n, 3, generator=gen)
    cpd = {k: v.double().requires_grad_(True) for k, v in cp.items()}
    ins = [t.double().requires_grad_(True) for t in (pts, dirs, feat, nrm)]
    rgb = O.color_obj_forward(cpd, *ins)
    names = list(cpd)
    ref = torch.autograd.grad((rgb * d_rgb.double()).sum(), ins + [cpd[k] for k in names])
    gins = [t.to(DEV).requires_grad_(True) for t in (pts, dirs, feat, nrm)]
    out = col(*gins, 0)
    assert max_abs(out, rgb) < 2e-5
    (out * d_rgb.to(DEV)).sum().backward()
    for t, r, nm in zip(gins, ref[:4], ("pts", "dirs", "feat", "normal")):
        assert rel_err(t.grad, r) < 1e-3, nm
    got = _param_grads(col)
    worst = {k: rel_err(got[k], r) for k, r in zip(names, ref[4:])}
    bad = {k: v for k, v in worst.items() if not v < 1e-3}
    assert not bad, bad


def test_frozen_weights_still_give_point_gradients():
    sdf, col, dev, sp, cp = obj_modules(requires_grad=False)
    pts = (0.4 * torch.randn(64, 3)).to(DEV).requires_grad_(True)
    s, f, n = sdf.fused(pts)
    (s.sum() + n.sum()).backward()
    assert pts.grad is not None and torch.isfinite(pts.grad).all()
    assert all(p.grad is None for p in sdf.parameters())


def test_sdf_grid_vs_golden():
    import honerf_b200 as H
    g = load_golden("sdf_grid")
    c = cases.sdf_grid_case()
    sdf, col, dev, _, _ = obj_modules(requires_grad=False)
    r = H.NeuSRenderer(sdf, dev, col, "obj", 64, 64, 0, 4, 1.0)
    u = r.sdf_grid(torch.full((3,), c["lo"]), torch.full((3,), c["hi"]), c["res"], chunk_points=500)
    assert max_abs(u, g["u"]) < 2e-5


def test_in_place_weight_update_between_forward_and_backward_raises():
    """The backward reads the packed weights shared by every call of the net; an in-place parameter update (optimizer step)
    between a call's forward and its backward is an error like torch's own version-counter check, not a silent wrong
    gradient."""
    import honerf_b200 as H
    sdf, col, _, _, _ = obj_modules()
    x = (0.4 * torch.randn(300, 3)).to(DEV).requires_grad_(True)
    s, f, n = sdf.fused(x)
    with torch.no_grad():
        sdf.lin3.weight_v.mul_(1.0001)
    with pytest.raises(H.HonerfError, match="changed between forward and backward"):
        (s.sum() + n.sum()).backward()
    # the next forward re-packs and works again
    s, f, n = sdf.fused(x)
    (s.sum() + f.sum() + n.sum()).backward()
    assert torch.isfinite(x.grad).all()
