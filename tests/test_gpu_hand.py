"""GPU parity of the hand (HALO pose-conditioned) field: embedding, SDF value/feature/normal, colour,
second-order backward to weights, points and bone transforms -- against the reference's golden vectors
and the fp64 oracle.  The hand field is ill-conditioned in fp32 (q = R x + t - T cancels metres down
to centimetres before 2^9-frequency encodings; normals reach |n| ~ 1e3), so bounds are relative."""
import pytest
import torch

import cases
import honerf_oracle as O
import synth
from golden_util import load_golden, max_abs, rel_err, rel_l2
from gpu_util import DEV, hand_modules

pytestmark = pytest.mark.gpu


def test_hand_fields_vs_golden():
    g = load_golden("hand_fields")
    c = cases.hand_fields_case()
    sdf, col, dev, _, _ = hand_modules()
    pts, bt, T = c["pts"].to(DEV), c["bt_inv"].to(DEV), c["T_pose_21"].to(DEV)
    out, xyz, _, _ = sdf(pts, bt, T)
    n = sdf.gradient(pts, bt, T).squeeze(1)
    print("xyz %.2e out %.2e normal rel %.2e" % (max_abs(xyz, g["xyz_feature"]), max_abs(out, g["sdf_out"]),
                                                  rel_l2(n, g["gradient"])))
    assert max_abs(xyz, g["xyz_feature"]) < 2e-4
    assert max_abs(out, g["sdf_out"]) < 1e-3
    assert rel_l2(n, g["gradient"]) < 1e-2
    rgb = col(None, xyz, out[:, 1:], None, n, 0)
    print("rgb %.2e" % max_abs(rgb, g["rgb"]))
    assert max_abs(rgb, g["rgb"]) < 1e-3
    assert max_abs(sdf.sdf(pts.detach(), bt, T), g["sdf_out"][:, :1]) < 1e-3


def test_hand_embedding_jacobian_paths_vs_fp64_oracle():
    """Every gradient of a random scalar functional of (sdf, feature, normal, xyz_feature): weights, points,
    bone transforms and T-pose joints; relative L2 error <= 1e-2 vs fp64 autograd (observed ~1e-4)."""
    c = cases.hand_fields_case()
    sdf, col, dev, sp, cp = hand_modules()
    pts = c["pts"]
    n = pts.shape[0]
    gen = torch.Generator().manual_seed(5)
    d_sdf, d_feat = torch.randn(n, 1, generator=gen), 0.1 * torch.randn(n, 256, generator=gen)
    d_n, d_xyz = 1e-2 * torch.randn(n, 3, generator=gen), 0.1 * torch.randn(n, 1386, generator=gen)
    spd = {k: v.double().requires_grad_(True) for k, v in sp.items() if k != "se3_refine"}
    x = pts.double().requires_grad_(True)
    btd = c["bt_inv"].double().requires_grad_(True)
    Td = c["T_pose_21"].double().requires_grad_(True)
    out, feat, _, _ = O.sdf_hand_forward(spd, x, btd, Td)
    nrm = O.sdf_gradient(lambda q: O.sdf_hand_forward(spd, q, btd, Td)[0][:, :1], x)
    L = (out[:, :1] * d_sdf.double()).sum() + (out[:, 1:] * d_feat.double()).sum() + (nrm * d_n.double()).sum() + \
        (feat * d_xyz.double()).sum()
    names = list(spd)
    ref = dict(zip(["pts", "bt_inv", "T"] + names, torch.autograd.grad(L, [x, btd, Td] + [spd[k] for k in names])))
    xg = pts.to(DEV).requires_grad_(True)
    btg = c["bt_inv"].to(DEV).requires_grad_(True)
    Tg = c["T_pose_21"].to(DEV).requires_grad_(True)
    s, f, nn, xyz = sdf.fused(xg, btg, Tg)
    assert rel_l2(nn, nrm) < 1e-2
    ((s * d_sdf.to(DEV)).sum() + (f * d_feat.to(DEV)).sum() + (nn * d_n.to(DEV)).sum() + (xyz * d_xyz.to(DEV)).sum()).backward()
    got = {"pts": xg.grad, "bt_inv": btg.grad, "T": Tg.grad}
    got.update({k: p.grad for k, p in sdf.named_parameters() if p.grad is not None})
    worst = {k: rel_l2(got[k], ref[k]) for k in ref}
    print("worst:", sorted(worst.items(), key=lambda kv: -kv[1])[:5])
    # the last row of each 4x4 is not a function input of the embedding
    assert float(btg.grad[:, 3, :].abs().max()) == 0.0
    bad = {k: v for k, v in worst.items() if not v < 1e-2}
    assert not bad, bad


def test_hand_color_backward_vs_oracle():
    sdf, col, dev, sp, cp = hand_modules()
    n = 300
    gen = torch.Generator().manual_seed(6)
    xyz, feat = 0.3 * torch.randn(n, 1386, generator=gen), torch.randn(n, 256, generator=gen)
    nrm, d_rgb = torch.randn(n, 3, generator=gen), torch.randn(n, 3, generator=gen)
    cpd = {k: v.double().requires_grad_(True) for k, v in cp.items()}
    ins = [t.double().requires_grad_(True) for t in (xyz, feat, nrm)]
    rgb = O.color_hand_forward(cpd, *ins)
    names = list(cpd)
    ref = torch.autograd.grad((rgb * d_rgb.double()).sum(), ins + [cpd[k] for k in names])
    gins = [t.to(DEV).requires_grad_(True) for t in (xyz, feat, nrm)]
    out = col(None, gins[0], gins[1], None, gins[2], 0)
    assert max_abs(out, rgb) < 5e-5
    (out * d_rgb.to(DEV)).sum().backward()
    for t, r, nm in zip(gins, ref[:3], ("xyz", "feat", "normal")):
        assert rel_l2(t.grad, r) < 1e-3, nm
    got = {k: p.grad for k, p in col.named_parameters()}
    worst = {k: rel_l2(got[k], r) for k, r in zip(names, ref[3:])}
    bad = {k: v for k, v in worst.items() if not v < 1e-3}
    assert not bad, bad


def test_hand_batched_equals_per_frame():
    """use_batch=True with [F,P,3] points and per-frame transforms == un-batched calls frame by frame
    (utils/fields.py:134-140; SURVEY E-3)."""
    sdf, col, dev, _, _ = hand_modules(requires_grad=False)
    btF, TF, JF = synth.hand_pose(n_frames=3)
    gen = torch.Generator().manual_seed(7)
    pts = torch.stack([JF[f][torch.randint(0, 21, (40,), generator=gen)] + 0.02 * torch.randn(40, 3, generator=gen)
                       for f in range(3)])
    s_b, f_b, n_b, x_b = sdf.fused(pts.to(DEV), btF.to(DEV), TF.to(DEV))
    for f in range(3):
        s, ft, n, x = sdf.fused(pts[f].to(DEV), btF[f].to(DEV), TF[f].to(DEV))
        sl = slice(f * 40, (f + 1) * 40)
        assert torch.equal(s, s_b[sl]) and torch.equal(x, x_b[sl]) and torch.equal(n, n_b[sl])


def test_hand_render_vs_golden():
    """NeuSRenderer.render with model_type='hand' (utils/renderer.py:190-258) vs the reference, end to end.
    The hand field is steep (|normal| up to 1e3), so an importance sample that crosses a cdf knot because of
    a 1e-6 SDF difference can move one ray's colour by 1e-2: most rays must agree to 2e-3 and the median
    to 1e-4; the strict check is the same-z test below."""
    import honerf_b200 as H
    import ref_conf
    from test_gpu_render import _fixed_rand
    g = load_golden("hand_render")
    c = cases.hand_render_case()
    R = c["R"]
    sdf, col, dev, _, _ = hand_modules()
    r = H.NeuSRenderer(sdf, dev, col, "hand", **ref_conf.RENDERER_CONF)
    bt = c["bt_inv"].to(DEV).requires_grad_(True)
    T = c["T_pose_21"].to(DEV).requires_grad_(True)
    with _fixed_rand(R["t_rand"]):
        out = r.render(R["rays_o"].to(DEV), R["rays_d"].to(DEV), R["near"], R["far"], bt, T, None, None, None, 0)
    assert set(out) == {"color_fine", "s_val", "cdf_fine", "weight_sum", "weight_max", "gradient_error"}
    err = (out["color_fine"].cpu() - g["color_fine"]).abs().max(dim=-1)[0]
    print("per-ray colour error:", err.tolist())
    assert (err < 2e-3).float().mean() >= 0.75 and err.median() < 2e-4
    loss = cases.hand_render_loss(out, c["true_rgb"].to(DEV))
    loss.backward()
    assert torch.isfinite(bt.grad).all() and torch.isfinite(T.grad).all()


def test_hand_render_core_given_same_z():
    """render_core (hand branch) on the oracle's z_vals: colour / weights <= 1e-3 abs, gradients to bone
    transforms, T-pose joints, variance and all weights <= 1e-2 relative (L2) vs the fp64 oracle."""
    _hand_same_z(None)


def _hand_same_z(own_factor, floor=1e-2):
    """own_factor: None -> every gradient within `floor`; a number -> within max(floor, own_factor x the error of the
    reference's OWN fp32 arithmetic against fp64 on this case) -- the synthetic hand weights make normals of magnitude
    up to ~70 whose sin/cos(8 n) encodings amplify any rounding of n (the reference in fp32 is itself 0.5-1.4e-2 off,
    depending on the host's thread count)."""
    import honerf_b200 as H
    import ref_conf
    c = cases.hand_render_case()
    R = c["R"]
    sdf, col, dev, sp, cp = hand_modules()
    r = H.NeuSRenderer(sdf, dev, col, "hand", **ref_conf.RENDERER_CONF)
    zref = O.render_hand(sp, cp, torch.tensor(0.3), R["rays_o"], R["rays_d"], R["near"], R["far"], c["bt_inv"],
                         c["T_pose_21"], R["t_rand"])["z_vals"]
    spd = {k: v.double().requires_grad_(k != "se3_refine") for k, v in sp.items()}
    cpd = {k: v.double().requires_grad_(True) for k, v in cp.items()}
    var = torch.tensor(0.3, dtype=torch.float64, requires_grad=True)
    btd = c["bt_inv"].double().requires_grad_(True)
    Td = c["T_pose_21"].double().requires_grad_(True)
    core = O.render_core_hand(spd, cpd, var, R["rays_o"].double(), R["rays_d"].double(), zref.double(), 1.1 / 64, btd, Td)
    ref_out = {"color_fine": core["color"], "weight_sum": core["weights"].sum(-1, keepdim=True),
               "gradient_error": core["gradient_error"]}
    ref_loss = cases.hand_render_loss(ref_out, c["true_rgb"].double())
    names = ["sdf." + k for k in spd if k != "se3_refine"] + ["color." + k for k in cpd] + ["variance", "bt_inv", "T"]
    tens = [v for k, v in spd.items() if k != "se3_refine"] + list(cpd.values()) + [var, btd, Td]
    ref_g = dict(zip(names, torch.autograd.grad(ref_loss, tens)))
    bt = c["bt_inv"].to(DEV).requires_grad_(True)
    T = c["T_pose_21"].to(DEV).requires_grad_(True)
    r.index = 0
    got_core = r.render_core(R["rays_o"].to(DEV), R["rays_d"].to(DEV), bt, T, None, zref.to(DEV), 1.1 / 64, sdf, dev, col)
    out = {"color_fine": got_core["color"], "weight_sum": got_core["weights"].sum(-1, keepdim=True),
           "gradient_error": got_core["gradient_error"]}
    print("colour %.2e weights %.2e eik rel %.2e" % (max_abs(out["color_fine"], ref_out["color_fine"]),
                                                      max_abs(got_core["weights"], core["weights"]),
                                                      rel_err(out["gradient_error"], ref_out["gradient_error"])))
    # weights agree to 1e-5; the colour net sees sin/cos(8 n) of normals with |n| up to ~1e3 on these
    # synthetic weights, so fp32 rounding of n (1e-6 relative) alone moves a colour by ~1e-3 -- in the
    # reference's own fp32 path as well.  Bound: 3e-3 against the fp64 oracle.
    assert max_abs(out["color_fine"], ref_out["color_fine"]) < 3e-3
    assert max_abs(got_core["weights"], core["weights"]) < 1e-4
    loss = cases.hand_render_loss(out, c["true_rgb"].to(DEV))
    loss.backward()
    got = {"sdf." + k: p.grad for k, p in sdf.named_parameters() if p.grad is not None}
    got.update({"color." + k: p.grad for k, p in col.named_parameters() if p.grad is not None})
    got.update({"variance": dev.variance.grad, "bt_inv": bt.grad, "T": T.grad})
    worst = {k: rel_l2(got[k], ref_g[k]) for k in names}
    print("worst:", sorted(worst.items(), key=lambda kv: -kv[1])[:5])
    bound = {k: floor for k in names}
    if own_factor is not None:
        sp32 = {k: v.clone().requires_grad_(k != "se3_refine") for k, v in sp.items()}
        cp32 = {k: v.clone().requires_grad_(True) for k, v in cp.items()}
        v32 = torch.tensor(0.3, requires_grad=True)
        bt32, T32 = c["bt_inv"].clone().requires_grad_(True), c["T_pose_21"].clone().requires_grad_(True)
        core32 = O.render_core_hand(sp32, cp32, v32, R["rays_o"], R["rays_d"], zref, 1.1 / 64, bt32, T32)
        out32 = {"color_fine": core32["color"], "weight_sum": core32["weights"].sum(-1, keepdim=True),
                 "gradient_error": core32["gradient_error"]}
        t32 = [v for k, v in sp32.items() if k != "se3_refine"] + list(cp32.values()) + [v32, bt32, T32]
        own = dict(zip(names, torch.autograd.grad(cases.hand_render_loss(out32, c["true_rgb"]), t32)))
        own = {k: rel_l2(own[k], ref_g[k]) for k in names}
        print("reference fp32 vs fp64:", sorted(own.items(), key=lambda kv: -kv[1])[:5])
        bound = {k: max(floor, own_factor * own[k]) for k in names}
    bad = {k: (v, bound[k]) for k, v in worst.items() if not v < bound[k]}
    assert not bad, bad
