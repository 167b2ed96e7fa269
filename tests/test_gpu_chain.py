"""Fused tile-chain kernels (HN_TC_BF16X3: split bf16 operands on tcgen05, activations resident in shared
memory across layers) against the fp32 SIMT verification path and the fp64 oracle."""
import pytest
import torch

import analytic as A
import synth
from golden_util import max_abs
from gpu_util import DEV, obj_modules

pytestmark = pytest.mark.gpu


def _pts(n, seed=0):
    g = torch.Generator().manual_seed(seed)
    return 0.45 * torch.randn(n, 3, generator=g)


@pytest.mark.parametrize("n", [1, 127, 128, 129, 1000, 40000])
def test_sdf_only_chain_vs_simt_and_fp64(n):
    """SDFNetwork_OBJ.sdf through the chain kernel (fp16 hi/lo operand pairs): <= 5e-5 abs vs fp64 (observed ~1e-5,
    the floor set by the tensor core's fp32 accumulation; single-pass bf16 would be ~6e-3), ragged tile tails included."""
    import honerf_b200 as H
    sdf, _, _, sp, _ = obj_modules(requires_grad=False)
    x = _pts(n, seed=n)
    got = H.ops.sdf_obj_sdf_only(sdf.packed(), x.to(DEV), 1.0, precision=H.ops._PRECISIONS["tc_bf16x3"])
    simt = H.ops.sdf_obj_sdf_only(sdf.packed(), x.to(DEV), 1.0, precision=H.ops._PRECISIONS["simt_fp32"])
    spd = {k: v.double() for k, v in sp.items() if k != "se3_refine"}
    Ws, bs = A.effective_weights(spd)
    ref = A.sdf_obj_fwd(Ws, bs, x.double())[0]
    e_ref, e_simt = max_abs(got, ref), max_abs(got, simt)
    print("n=%d  chain vs fp64 %.2e  chain vs simt %.2e  simt vs fp64 %.2e" % (n, e_ref, e_simt, max_abs(simt, ref)))
    assert torch.isfinite(got).all()
    assert e_ref < 5e-5 and e_simt < 5e-5


def test_sdf_only_chain_geometric_init_and_scale():
    """Un-perturbed geometric-init weights (most of lin0 / skip columns are zero) and scale != 1."""
    import honerf_b200 as H
    import ref_conf
    torch.manual_seed(3)
    net = H.SDFNetwork_OBJ(H.Embedding(), 4, "real", **dict(ref_conf.OBJ_SDF_CONF, scale=2.0)).to(DEV)
    x = _pts(5000, seed=9).to(DEV)
    a = H.ops.sdf_obj_sdf_only(net.packed(), x, 0.5, precision=H.ops._PRECISIONS["tc_bf16x3"])
    b = H.ops.sdf_obj_sdf_only(net.packed(), x, 0.5, precision=H.ops._PRECISIONS["simt_fp32"])
    assert max_abs(a, b) < 2e-5


@pytest.mark.parametrize("n", [1, 130, 3000])
def test_fwd_chain_value_feature_normal_and_backward(n):
    """Fused forward (value trunk + feature head + normal sweep, 17 chained layers per tile) vs fp64:
    sdf / feature 5e-5 abs, normal 1e-4 relative; then the second-order backward fed by the chain kernel's
    stash vs fp64 autograd: every weight gradient and d_pts within 1e-3 relative (L2)."""
    import honerf_b200 as H
    from golden_util import rel_l2
    sdf, _, _, sp, _ = obj_modules()
    x = _pts(n, seed=100 + n)
    g = torch.Generator().manual_seed(n)
    d_sdf, d_feat, d_n = torch.randn(n, 1, generator=g), 0.1 * torch.randn(n, 256, generator=g), torch.randn(n, 3, generator=g)
    spd = {k: v.double().requires_grad_(True) for k, v in sp.items() if k != "se3_refine"}
    Ws, bs = A.effective_weights(spd)
    xd = x.double().requires_grad_(True)
    rs, rf, rn, _ = A.sdf_obj_fwd(Ws, bs, xd)
    # autograd through the analytic forward (itself validated against the reference in test_analytic_model.py)
    L = (rs * d_sdf.double()).sum() + (rf * d_feat.double()).sum() + (rn * d_n.double()).sum()
    names = list(spd)
    ref_g = dict(zip(["pts"] + names, torch.autograd.grad(L, [xd] + [spd[k] for k in names])))
    xg = x.to(DEV).requires_grad_(True)
    s, f, nn = H.ops.sdf_obj(sdf.packed(), xg, 1.0, precision=H.ops._PRECISIONS["tc_bf16x3"])
    print("n=%d sdf %.2e feat %.2e normal rel %.2e" % (n, max_abs(s, rs), max_abs(f, rf), rel_l2(nn, rn)))
    assert max_abs(s, rs) < 5e-5 and max_abs(f, rf) < 5e-5 and rel_l2(nn, rn) < 1e-4
    ((s * d_sdf.to(DEV)).sum() + (f * d_feat.to(DEV)).sum() + (nn * d_n.to(DEV)).sum()).backward()
    got = {"pts": xg.grad}
    got.update({k: p.grad for k, p in sdf.named_parameters() if p.grad is not None})
    worst = {k: rel_l2(got[k], ref_g[k]) for k in ref_g}
    print("worst:", sorted(worst.items(), key=lambda kv: -kv[1])[:4])
    assert all(v < 1e-3 for v in worst.values()), worst


def _tile(x):
    """[n, c<=256] -> the chain path's tiled layout [tile][col/4][128][4] (n padded to 128, columns to 256)."""
    n, c = x.shape
    npad = (n + 127) // 128 * 128
    full = torch.zeros(npad, 256, device=x.device)
    full[:n, :c] = x
    return full.reshape(npad // 128, 128, 64, 4).permute(0, 2, 1, 3).contiguous()


@pytest.mark.parametrize("n,out,nin,tiled,two", [(1000, 256, 256, 1, True), (300, 193, 256, 1, True), (4096, 256, 63, 0, True),
                                                  (129, 256, 256, 0, False), (1, 3, 17, 0, False), (70000, 256, 256, 1, True)])
def test_weight_gradient_kernel(n, out, nin, tiled, two):
    """dW = P^T Q (+ P2^T Q2) and db = colsum(P) from the MN-major bf16x3 split-K kernel vs fp64 matmul:
    <= 5e-5 of the largest entry (observed 3e-6 .. 3e-5 at 70 000 points)."""
    import ctypes
    import honerf_b200 as H
    from honerf_b200 import _lib
    g = torch.Generator().manual_seed(n + out)
    mk = lambda c: torch.randn(n, c, generator=g).to(DEV)
    P, Q, P2, Q2 = mk(out), mk(nin), mk(out), mk(nin)
    ref = P.double().T @ Q.double() + ((P2.double().T @ Q2.double()) if two else 0)
    ldc = (nin + 3) // 4 * 4
    C = torch.zeros(out, ldc, device=DEV)
    db = torch.zeros(out, device=DEV)
    part = torch.empty(16 * 65536, device=DEV)
    args = [_tile(t) if tiled else t for t in (P, Q, P2, Q2)]
    ptr = lambda t: ctypes.c_void_p(t.data_ptr())
    _lib.check(_lib.lib.hn_dw_test(ptr(args[0]), out, tiled, out, ptr(args[1]), nin, tiled, nin, ptr(args[2]) if two else None,
                                   ptr(args[3]) if two else None, n, ptr(C), ldc, ptr(db), ptr(part), part.numel(),
                                   ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)), "hn_dw_test")
    err = max_abs(C[:, :nin], ref) / float(ref.abs().max())
    errb = max_abs(db, P.double().sum(0)) / float(P.double().sum(0).abs().max())
    print("n=%d %dx%d tiled=%d: dW rel-to-max %.2e, db %.2e" % (n, out, nin, tiled, err, errb))
    assert err < 5e-5 and errb < 1e-5
    if ldc > nin:
        assert float(C[:, nin:].abs().max()) == 0.0


@pytest.mark.parametrize("n", [77, 4096 + 130])
def test_weight_gradient_kernel_column_major_operand(n):
    """The first SDF layer's job: P in the tiled stash layout, Q (the 39-wide encoding) in column-major [64][128]
    tiles; same bound as above."""
    import ctypes
    from honerf_b200 import _lib
    g = torch.Generator().manual_seed(n)
    mk = lambda c: torch.randn(n, c, generator=g).to(DEV)
    P, Q, P2, Q2 = mk(256), mk(39), mk(256), mk(39)
    ref = P.double().T @ Q.double() + P2.double().T @ Q2.double()

    def colmajor(x):
        npad = (n + 127) // 128 * 128
        full = torch.zeros(npad, 64, device=DEV)
        full[:n, :39] = x
        return full.reshape(npad // 128, 128, 64).permute(0, 2, 1).contiguous()

    C = torch.zeros(256, 40, device=DEV)
    db = torch.zeros(256, device=DEV)
    part = torch.empty(16 * 65536, device=DEV)
    args = [_tile(P), colmajor(Q), _tile(P2), colmajor(Q2)]
    ptr = lambda t: ctypes.c_void_p(t.data_ptr())
    _lib.check(_lib.lib.hn_dw_test(ptr(args[0]), 0, 1, 256, ptr(args[1]), 64, 2, 39, ptr(args[2]), ptr(args[3]), n, ptr(C), 40,
                                   ptr(db), ptr(part), part.numel(), ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)),
               "hn_dw_test")
    err = max_abs(C[:, :39], ref) / float(ref.abs().max())
    errb = max_abs(db, P.double().sum(0)) / float(P.double().sum(0).abs().max())
    print("n=%d column-major Q: dW rel-to-max %.2e, db %.2e" % (n, err, errb))
    assert err < 5e-5 and errb < 1e-5 and float(C[:, 39:].abs().max()) == 0.0


@pytest.mark.parametrize("n", [1, 200, 5000])
def test_color_chain_forward_backward(n):
    """RenderingNetwork_OBJ through the colour chain kernels (split 373-wide first layer, ReLU chain, N=16 output
    MMA) vs fp64 autograd of the oracle: rgb 2e-5 abs; every input / weight gradient 1e-3 relative (L2)."""
    import honerf_b200 as H
    import honerf_oracle as O
    from golden_util import rel_l2
    _, col, _, _, cp = obj_modules()
    g = torch.Generator().manual_seed(7 * n + 1)
    pts = 0.45 * torch.randn(n, 3, generator=g)
    dirs = torch.nn.functional.normalize(torch.randn(n, 3, generator=g), dim=-1)
    feat = 0.3 * torch.randn(n, 256, generator=g)
    nrm = torch.randn(n, 3, generator=g)
    w = torch.randn(n, 3, generator=g)
    cpd = {k: v.double().requires_grad_(True) for k, v in cp.items()}
    ins = [t.double().requires_grad_(True) for t in (pts, dirs, feat, nrm)]
    ref = O.color_obj_forward(cpd, *ins)
    names = list(cpd)
    ref_g = dict(zip(["pts", "dirs", "feat", "normal"] + names,
                     torch.autograd.grad((ref * w.double()).sum(), ins + [cpd[k] for k in names])))
    gi = [t.to(DEV).requires_grad_(True) for t in (pts, dirs, feat, nrm)]
    rgb = H.ops.color_obj(col.packed(), *gi, precision=H.ops._PRECISIONS["tc_bf16x3"])
    print("n=%d rgb %.2e" % (n, max_abs(rgb, ref)))
    assert max_abs(rgb, ref) < 2e-5
    (rgb * w.to(DEV)).sum().backward()
    got = dict(zip(["pts", "dirs", "feat", "normal"], [t.grad for t in gi]))
    got.update({k: p.grad for k, p in col.named_parameters() if p.grad is not None})
    # A ReLU whose pre-activation is within the forward's rounding error of zero may take the other branch than
    # the fp64 reference (its derivative is discontinuous there; the reference's own fp32 arithmetic does the
    # same, 100x less often).  Such a flip changes that ONE point's gradient by a few percent, so: per point,
    # the input gradients must agree to 1e-3 for >= 97 % of the points with a median <= 1e-4; tensors summed over
    # points (weight gradients) inherit the few flipped points: 3e-2.
    per_pt = ((got["feat"].double().cpu() - ref_g["feat"]).norm(dim=1) / ref_g["feat"].norm(dim=1))
    print("per-point d_feat error: median %.2e, frac < 1e-3: %.3f" % (per_pt.median(), (per_pt < 1e-3).float().mean()))
    assert per_pt.median() < 1e-4 and (per_pt < 1e-3).float().mean() >= 0.97
    worst = {k: rel_l2(got[k], ref_g[k]) for k in ref_g}
    print("worst:", sorted(worst.items(), key=lambda kv: -kv[1])[:4])
    assert all(v < 3e-2 for v in worst.values()), worst
