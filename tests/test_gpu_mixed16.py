"""HN_TC_MIXED16 ('tc_mixed16'): the object SDF operator with the fp16x3 value trunk, single-16-bit-operand sweeps in
tensor memory and the 16-bit dW-ready activation stash (csrc/chain16*.cu), against fp64 (oracle/analytic.py, itself
pinned to the reference's autograd in tests/test_analytic_model.py).  Stated bounds (profiles/r02_precision_table.md has
the CPU emulation they were chosen from): sdf / feature 5e-5 abs (the trunk is unchanged arithmetic), normal 2e-3
relative, d_pts and every weight / bias gradient 1e-2 relative (L2) -- the north star's gradient tolerance."""
import ctypes

import pytest
import torch

import analytic as A
from golden_util import max_abs, rel_l2
from gpu_util import DEV, obj_modules

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("n,out,nin,two", [(1000, 256, 256, True), (300, 193, 256, True), (4096, 256, 63, True),
                                           (129, 256, 256, False), (70000, 256, 256, True), (64, 256, 64, False)])
def test_dw16_kernel(n, out, nin, two):
    """dW = P^T Q (+ P2^T Q2), db = colsum(P) from bf16 dW-ready tiles (un-swizzled MN-major UMMA operands fetched by
    32 KB bulk copies, one MMA per product) vs fp64 matmul of the SAME bf16-rounded operands: <= 2e-5 of the largest entry
    (fp32 accumulation only)."""
    from honerf_b200 import _lib
    g = torch.Generator().manual_seed(n + out)
    mk = lambda c: torch.randn(n, c, generator=g).to(DEV)
    P, Q, P2, Q2 = mk(out), mk(nin), mk(out), mk(nin)
    r16 = lambda t: t.bfloat16().double()
    ref = r16(P).T @ r16(Q) + ((r16(P2).T @ r16(Q2)) if two else 0)
    ldc = (nin + 3) // 4 * 4
    C = torch.zeros(out, ldc, device=DEV)
    db = torch.zeros(out, device=DEV)
    part = torch.empty(16 * 65536, device=DEV)
    npad = (n + 127) // 128 * 128
    tiles = torch.empty(4 * npad * 512, device=DEV, dtype=torch.uint8)
    ptr = lambda t: ctypes.c_void_p(t.data_ptr())
    _lib.check(_lib.lib.hn_dw16_test(ptr(P), out, ptr(Q), nin, ptr(P2) if two else None, ptr(Q2) if two else None, n, ptr(C),
                                     ldc, ptr(db), ptr(tiles), tiles.numel(), ptr(part), part.numel(),
                                     ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)), "hn_dw16_test")
    torch.cuda.synchronize()
    err = max_abs(C[:, :nin], ref) / float(ref.abs().max())
    errb = max_abs(db, r16(P).sum(0)) / float(r16(P).sum(0).abs().max())
    print("n=%d %dx%d: dW rel-to-max %.2e, db %.2e" % (n, out, nin, err, errb))
    assert err < 2e-5 and errb < 1e-5
    if ldc > nin:
        assert float(C[:, nin:].abs().max()) == 0.0


def _pts(n, seed=0):
    g = torch.Generator().manual_seed(seed)
    return 0.45 * torch.randn(n, 3, generator=g)


@pytest.mark.parametrize("n", [1, 130, 3000, 148 * 128 + 1])
def test_mixed16_operator_vs_fp64(n):
    import honerf_b200 as H
    prec = H.ops._PRECISIONS["tc_mixed16"]
    sdf, _, _, sp, _ = obj_modules()
    x = _pts(n, seed=200 + n)
    g = torch.Generator().manual_seed(n)
    d_sdf, d_feat, d_n = torch.randn(n, 1, generator=g), 0.1 * torch.randn(n, 256, generator=g), torch.randn(n, 3, generator=g)
    spd = {k: v.double().requires_grad_(True) for k, v in sp.items() if k != "se3_refine"}
    Ws, bs = A.effective_weights(spd)
    xd = x.double().requires_grad_(True)
    rs, rf, rn, _ = A.sdf_obj_fwd(Ws, bs, xd)
    L = (rs * d_sdf.double()).sum() + (rf * d_feat.double()).sum() + (rn * d_n.double()).sum()
    names = list(spd)
    ref_g = dict(zip(["pts"] + names, torch.autograd.grad(L, [xd] + [spd[k] for k in names])))
    xg = x.to(DEV).requires_grad_(True)
    s, f, nn = H.ops.sdf_obj(sdf.packed(), xg, 1.0, precision=prec)
    torch.cuda.synchronize()
    print("n=%d sdf %.2e feat %.2e normal rel %.2e" % (n, max_abs(s, rs), max_abs(f, rf), rel_l2(nn, rn)))
    assert torch.isfinite(s).all() and torch.isfinite(f).all() and torch.isfinite(nn).all()
    assert max_abs(s, rs) < 5e-5 and max_abs(f, rf) < 5e-5 and rel_l2(nn, rn) < 2e-3
    per_pt = (nn.detach().cpu().double() - rn.detach()).norm(dim=1) / rn.detach().norm(dim=1).clamp_min(1e-3)
    assert float(per_pt.max()) < 2e-2, float(per_pt.max())
    ((s * d_sdf.to(DEV)).sum() + (f * d_feat.to(DEV)).sum() + (nn * d_n.to(DEV)).sum()).backward()
    torch.cuda.synchronize()
    got = {"pts": xg.grad}
    got.update({k: p.grad for k, p in sdf.named_parameters() if p.grad is not None})
    worst = {k: rel_l2(got[k], ref_g[k]) for k in ref_g}
    print("worst:", sorted(worst.items(), key=lambda kv: -kv[1])[:5])
    assert all(v < 1e-2 for v in worst.values()), worst


def test_mixed16_matches_bf16x3_path_loosely_and_no_weight_grads():
    """Same inputs through both tensor-core paths: values to 5e-5, normals 2e-3; and the pose-fitting case (frozen
    weights: the kernel skips the weight-gradient operand stores) still returns d_pts."""
    import honerf_b200 as H
    sdf, _, _, _, _ = obj_modules(requires_grad=False)
    x = _pts(5000, seed=5).to(DEV).requires_grad_(True)
    a = H.ops.sdf_obj(sdf.packed(), x, 1.0, precision=H.ops._PRECISIONS["tc_mixed16"])
    b = H.ops.sdf_obj(sdf.packed(), x, 1.0, precision=H.ops._PRECISIONS["tc_bf16x3"])
    assert max_abs(a[0], b[0]) < 5e-5 and max_abs(a[1], b[1]) < 5e-5 and rel_l2(a[2], b[2]) < 2e-3
    w = torch.randn(5000, 3, device=DEV)
    ga, = torch.autograd.grad((a[2] * w).sum() + a[0].sum(), x)
    gb, = torch.autograd.grad((b[2] * w).sum() + b[0].sum(), x)
    assert rel_l2(ga, gb) < 1e-2


def test_colour_net_16bit_stash_matches_the_fp32_stash_path():
    """HN_TC_MIXED16 keeps the colour chain's arithmetic and only changes what is stashed (hi / lo bf16 T16 tile pairs) and
    which kernel forms the weight gradients (dw16_kernel, three MMAs per product on the pairs): forward and input gradients
    bit-identical to HN_TC_BF16X3, weight / bias gradients within 2e-5 relative (fp32 summation order only)."""
    import honerf_b200 as H
    n = 40000
    _, col, _, _, _ = obj_modules()
    g = torch.Generator().manual_seed(0)
    x = (0.45 * torch.randn(n, 3, generator=g)).to(DEV)
    d = torch.nn.functional.normalize(torch.randn(n, 3, generator=g), dim=-1).to(DEV)
    feat = (0.3 * torch.randn(n, 256, generator=g)).to(DEV)
    nrm = torch.nn.functional.normalize(torch.randn(n, 3, generator=g), dim=-1).to(DEV)
    go = (1e-3 * torch.randn(n, 3, generator=g)).to(DEV)
    res = {}
    for name in ("tc_bf16x3", "tc_mixed16"):
        for q in col.parameters():
            q.grad = None
        xs = [t.clone().requires_grad_(True) for t in (x, d, feat, nrm)]
        rgb = H.ops.color_obj(col.packed(), xs[0], xs[1], xs[2], xs[3], precision=H.ops._PRECISIONS[name])
        (rgb * go).sum().backward()
        res[name] = ({k: q.grad.clone() for k, q in col.named_parameters()}, [t.grad for t in xs], rgb.detach())
    a, b = res["tc_mixed16"], res["tc_bf16x3"]
    assert torch.equal(a[2], b[2]) and all(torch.equal(u, v) for u, v in zip(a[1], b[1]))
    worst = {k: rel_l2(a[0][k], b[0][k]) for k in a[0]}
    print("colour weight gradients, 16-bit pair stash vs fp32 stash:", sorted(worst.items(), key=lambda kv: -kv[1])[:3])
    assert all(v < 2e-5 for v in worst.values()), worst
