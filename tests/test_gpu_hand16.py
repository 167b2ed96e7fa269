"""HN_TC_MIXED16 on the hand SDF field (csrc/chain16_hand.cu): the 256 x 256 layers of SDFNetwork (utils/fields.py:56-177) as
tile-chain kernels with the activations in tensor memory, used when no weight gradient is asked for (pose fitting,
rendering).  Checked against the fp64 oracle (oracle/honerf_oracle.py, pinned to the reference's golden vectors in
tests/test_oracle_golden.py) and against the per-layer path (HN_TC_BF16X3) on the same inputs: sdf / feature 1e-3 abs (observed
~1e-5), normal 1e-2 relative, gradients to the points and the bone transforms 1e-2 relative (L2)."""
import pytest
import torch

import cases
import honerf_oracle as O
import synth
from golden_util import load_golden, max_abs, rel_l2
from gpu_util import DEV, hand_modules

pytestmark = pytest.mark.gpu


def _case(n_rep=1, seed=5):
    c = cases.hand_fields_case()
    pts = c["pts"]
    if n_rep > 1:
        g = torch.Generator().manual_seed(seed)
        pts = torch.cat([pts + 0.004 * torch.randn(pts.shape, generator=g) for _ in range(n_rep)], 0)
    return pts, c["bt_inv"], c["T_pose_21"]


def test_frozen_weights_take_the_chain_path_and_match_golden():
    import honerf_b200 as H
    g = load_golden("hand_fields")
    c = cases.hand_fields_case()
    sdf, col, dev, _, _ = hand_modules(requires_grad=False)
    pts, bt, T = c["pts"].to(DEV), c["bt_inv"].to(DEV), c["T_pose_21"].to(DEV)
    k0 = H.runtime.launch_count() if hasattr(H, "runtime") else None
    s, f, n, xyz = sdf.fused(pts, bt, T)
    out = torch.cat([s, f], -1)
    print("m16: out %.2e normal rel %.2e xyz %.2e" % (max_abs(out, g["sdf_out"]), rel_l2(n, g["gradient"]), max_abs(xyz, g["xyz_feature"])))
    assert max_abs(xyz, g["xyz_feature"]) < 2e-4
    assert max_abs(out, g["sdf_out"]) < 1e-3
    assert rel_l2(n, g["gradient"]) < 1e-2
    assert max_abs(sdf.sdf(pts, bt, T), g["sdf_out"][:, :1]) < 1e-3
    # same call on the per-layer path
    p3 = H.ops._PRECISIONS["tc_bf16x3"]
    s3, f3, n3, _ = H.ops.sdf_hand(sdf.packed(), pts, bt, T, precision=p3)
    print("m16 vs bf16x3: sdf %.2e feat %.2e normal rel %.2e" % (max_abs(s, s3), max_abs(f, f3), rel_l2(n, n3)))
    assert max_abs(s, s3) < 5e-5 and max_abs(f, f3) < 2e-4 and rel_l2(n, n3) < 1e-3


@pytest.mark.parametrize("n_rep", [1, 9])
def test_pose_and_point_gradients_vs_fp64_oracle(n_rep):
    """Gradients of a random scalar functional of (sdf, feature, normal, xyz_feature) to the points, the bone transforms
    and the T-pose joints, weights frozen; n_rep = 9 makes several tiles per CTA-free ragged sizes (n % 128 != 0)."""
    pts, bt0, T0 = _case(n_rep)
    sdf, col, dev, sp, cp = hand_modules(requires_grad=False)
    n = pts.shape[0]
    gen = torch.Generator().manual_seed(5)
    d_sdf, d_feat = torch.randn(n, 1, generator=gen), 0.1 * torch.randn(n, 256, generator=gen)
    d_n, d_xyz = 1e-2 * torch.randn(n, 3, generator=gen), 0.1 * torch.randn(n, 1386, generator=gen)
    spd = {k: v.double() for k, v in sp.items() if k != "se3_refine"}
    x = pts.double().requires_grad_(True)
    btd = bt0.double().requires_grad_(True)
    Td = T0.double().requires_grad_(True)
    out, feat, _, _ = O.sdf_hand_forward(spd, x, btd, Td)
    nrm = O.sdf_gradient(lambda q: O.sdf_hand_forward(spd, q, btd, Td)[0][:, :1], x)
    L = (out[:, :1] * d_sdf.double()).sum() + (out[:, 1:] * d_feat.double()).sum() + (nrm * d_n.double()).sum() + \
        (feat * d_xyz.double()).sum()
    ref = dict(zip(["pts", "bt_inv", "T"], torch.autograd.grad(L, [x, btd, Td])))
    xg = pts.to(DEV).requires_grad_(True)
    btg = bt0.to(DEV).requires_grad_(True)
    Tg = T0.to(DEV).requires_grad_(True)
    s, f, nn, xyz = sdf.fused(xg, btg, Tg)
    print("n=%d: sdf %.2e feat %.2e normal rel %.2e" % (n, max_abs(s, out[:, :1]), max_abs(f, out[:, 1:]), rel_l2(nn, nrm)))
    assert max_abs(s, out[:, :1]) < 1e-3 and max_abs(f, out[:, 1:]) < 1e-3 and rel_l2(nn, nrm) < 1e-2
    ((s * d_sdf.to(DEV)).sum() + (f * d_feat.to(DEV)).sum() + (nn * d_n.to(DEV)).sum() + (xyz * d_xyz.to(DEV)).sum()).backward()
    errs = {"pts": rel_l2(xg.grad, ref["pts"]), "bt_inv": rel_l2(btg.grad[..., :3, :], ref["bt_inv"][..., :3, :]),
            "T": rel_l2(Tg.grad, ref["T"])}
    print("  gradient rel errors:", {k: "%.2e" % v for k, v in errs.items()})
    for k, v in errs.items():
        assert v < 1e-2, (k, v)
    # without the normal / xyz cotangents (None gradients reach the operator)
    xg.grad = btg.grad = Tg.grad = None
    s, f, nn, xyz = sdf.fused(xg, btg, Tg)
    (s * d_sdf.to(DEV)).sum().backward()
    L2 = (O.sdf_hand_forward(spd, x, btd, Td)[0][:, :1] * d_sdf.double()).sum()
    r2 = torch.autograd.grad(L2, [x, btd])
    assert rel_l2(xg.grad, r2[0]) < 1e-2 and rel_l2(btg.grad[..., :3, :], r2[1][..., :3, :]) < 1e-2


def test_trainable_weights_stay_on_the_per_layer_path():
    """With trainable weights (and grad mode on) the default precision falls back to the per-layer contractions (fp32 stash,
    weight gradients): hn_sdf_hand_bwd rejects grad != NULL under HN_TC_MIXED16 rather than returning zeros."""
    import honerf_b200 as H
    pts, bt0, T0 = _case()
    sdf_t, _, _, _, _ = hand_modules(requires_grad=True)
    sdf_f, _, _, _, _ = hand_modules(requires_grad=False)
    x = pts.to(DEV)
    s_t, f_t, n_t, _ = sdf_t.fused(x, bt0.to(DEV), T0.to(DEV))
    s3, f3, n3, _ = H.ops.sdf_hand(sdf_f.packed(), x, bt0.to(DEV), T0.to(DEV), precision=H.ops._PRECISIONS["tc_bf16x3"])
    assert torch.equal(s_t, s3) and torch.equal(f_t, f3) and torch.equal(n_t, n3)
    (s_t.sum() + n_t.sum()).backward()
    assert all(p.grad is not None and torch.isfinite(p.grad).all() for k, p in sdf_t.named_parameters() if k != "se3_refine")
    # under no_grad the same net renders through the chain kernels (bit-identical to the frozen net's default path)
    with torch.no_grad():
        s_n, f_n, n_n, _ = sdf_t.fused(x, bt0.to(DEV), T0.to(DEV))
    s_f, f_f, n_f, _ = sdf_f.fused(x, bt0.to(DEV), T0.to(DEV))
    assert torch.equal(s_n, s_f) and torch.equal(f_n, f_f) and torch.equal(n_n, n_f) and not torch.equal(s_n, s3)


def test_sdf_only_sizes():
    """sdf-only trunk (up-sampling passes) at ragged sizes, against the per-layer path."""
    import honerf_b200 as H
    sdf, _, _, _, _ = hand_modules(requires_grad=False)
    bt0, T0, J = synth.hand_pose()
    for n in (1, 127, 129, 148 * 128 + 5):
        g = torch.Generator().manual_seed(n)
        x = (J[torch.randint(0, 21, (n,), generator=g)] + 0.03 * torch.randn(n, 3, generator=g)).to(DEV)
        a = sdf.sdf(x, bt0.to(DEV), T0.to(DEV))
        b = H.ops.sdf_hand_sdf_only(sdf.packed(), x, bt0.to(DEV), T0.to(DEV), precision=H.ops._PRECISIONS["tc_bf16x3"])
        assert a.shape == (n, 1) and max_abs(a, b) < 5e-5, (n, max_abs(a, b))


def test_many_tiles_per_cta_match_the_per_layer_path():
    """More tiles than CTAs (every persistent CTA walks several tiles: barrier phases, input-slot parities): forward and the
    pose / point gradients against the per-layer HN_TC_BF16X3 path on the same 38 000-odd points."""
    import honerf_b200 as H
    sdf, _, _, _, _ = hand_modules(requires_grad=False)
    bt0, T0, J = synth.hand_pose()
    n = 148 * 128 * 2 + 77
    g = torch.Generator().manual_seed(11)
    x0 = (J[torch.randint(0, 21, (n,), generator=g)] + 0.03 * torch.randn(n, 3, generator=g)).to(DEV)
    d_sdf, d_feat = torch.randn(n, 1, generator=g).to(DEV), (0.1 * torch.randn(n, 256, generator=g)).to(DEV)
    d_n = (1e-2 * torch.randn(n, 3, generator=g)).to(DEV)
    res = {}
    for name in ("tc_mixed16", "tc_bf16x3"):
        x = x0.clone().requires_grad_(True)
        bt = bt0.to(DEV).requires_grad_(True)
        s, f, nn, xyz = H.ops.sdf_hand(sdf.packed(), x, bt, T0.to(DEV), precision=H.ops._PRECISIONS[name])
        ((s * d_sdf).sum() + (f * d_feat).sum() + (nn * d_n).sum()).backward()
        res[name] = (s.detach(), f.detach(), nn.detach(), x.grad, bt.grad)
    a, b = res["tc_mixed16"], res["tc_bf16x3"]
    errs = [max_abs(a[0], b[0]), max_abs(a[1], b[1]), rel_l2(a[2], b[2]), rel_l2(a[3], b[3]), rel_l2(a[4], b[4])]
    print("m16 vs per-layer at n=%d: sdf %.2e feat %.2e normal %.2e d_pts %.2e d_bt %.2e" % tuple([n] + errs))
    assert errs[0] < 5e-5 and errs[1] < 2e-4 and errs[2] < 1e-3 and errs[3] < 5e-3 and errs[4] < 5e-3


def test_batched_frames_match_the_per_layer_path():
    """use_batch=True: bt_inv [F,21,4,4], points [F,P,3] with P = 300 (frame boundaries inside tiles): chain kernels against
    the per-layer path, forward and the gradients to every frame's bone transforms."""
    import honerf_b200 as H
    sdf, _, _, _, _ = hand_modules(requires_grad=False, use_batch=True)
    Fn, P = 3, 300
    btF, TF_, JF = synth.hand_pose(n_frames=Fn)
    g = torch.Generator().manual_seed(21)
    x0 = torch.stack([JF[f][torch.randint(0, 21, (P,), generator=g)] + 0.03 * torch.randn(P, 3, generator=g) for f in range(Fn)]).to(DEV)
    d_sdf = torch.randn(Fn * P, 1, generator=g).to(DEV)
    d_n = (1e-2 * torch.randn(Fn * P, 3, generator=g)).to(DEV)
    res = {}
    for name in ("tc_mixed16", "tc_bf16x3"):
        x = x0.clone().requires_grad_(True)
        bt = btF.to(DEV).requires_grad_(True)
        s, f, nn, xyz = H.ops.sdf_hand(sdf.packed(), x, bt, TF_.to(DEV), precision=H.ops._PRECISIONS[name])
        ((s * d_sdf).sum() + (nn * d_n).sum() + 0.01 * f.sum()).backward()
        res[name] = (s.detach(), f.detach(), nn.detach(), x.grad, bt.grad)
    a, b = res["tc_mixed16"], res["tc_bf16x3"]
    errs = [max_abs(a[0], b[0]), max_abs(a[1], b[1]), rel_l2(a[2], b[2]), rel_l2(a[3], b[3])] + \
           [rel_l2(a[4][f], b[4][f]) for f in range(Fn)]
    print("batched frames, m16 vs per-layer: sdf %.2e feat %.2e normal %.2e d_pts %.2e d_bt per frame %s" % (
        errs[0], errs[1], errs[2], errs[3], ["%.2e" % e for e in errs[4:]]))
    assert errs[0] < 5e-5 and errs[1] < 2e-4 and errs[2] < 1e-3 and all(e < 5e-3 for e in errs[3:])


def test_hand_colour_render_path_matches_the_differentiable_path():
    """Forward-only rendering runs layers 1..3 + output of the hand colour net on the colour chain kernel
    (hn_color_hand_fwd_render); same colours as the differentiable per-layer call: 2e-5 abs."""
    import honerf_b200 as H
    _, col, _, _, _ = hand_modules(requires_grad=False)
    for n in (1, 130, 20000):
        g = torch.Generator().manual_seed(n)
        xyz = (0.3 * torch.randn(n, 1386, generator=g)).to(DEV)
        feat = (0.3 * torch.randn(n, 256, generator=g)).to(DEV)
        nrm = torch.nn.functional.normalize(torch.randn(n, 3, generator=g), dim=-1).to(DEV)
        b = H.ops.color_hand(col.packed(), xyz.clone().requires_grad_(True), feat, nrm).detach()
        with torch.no_grad():
            a = H.ops.color_hand(col.packed(), xyz, feat, nrm)          # contiguous rows (ld 1386): input row assembled
            # the way the renderer passes it: a 16-byte aligned view of the SDF operator's stash rows (ld 1644), which the
            # first layer then reads in place (two partial contractions, no 1672-wide input row)
            wide = torch.zeros(n, 1644, device=DEV)
            wide[:, 256:256 + 1386] = xyz
            a2 = H.ops.color_hand(col.packed(), wide[:, 256:256 + 1386], feat, nrm)
        assert a.shape == (n, 3) and max_abs(a, b) < 2e-5, (n, max_abs(a, b))
        assert max_abs(a2, b) < 2e-5, (n, max_abs(a2, b))
