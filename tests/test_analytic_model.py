"""The CPU model of the kernel algorithm (oracle/analytic.py: explicit forward / normal sweep /
tangent + reverse sweeps) against fp64 autograd double-backward through the oracle."""
import torch

import analytic as A
import honerf_oracle as O
import synth


def test_second_order_backward_formulas_fp64():
    sp, _ = synth.obj_states()
    N = 96
    g = torch.Generator().manual_seed(0)
    x = (0.45 * torch.randn(N, 3, generator=g)).double()
    spd = {k: v.double().requires_grad_(True) for k, v in sp.items() if k != "se3_refine"}
    xg = x.clone().requires_grad_(True)
    out = O.sdf_obj_forward(spd, xg)
    nrm = O.sdf_gradient(lambda q: O.sdf_obj_forward(spd, q)[:, :1], xg)
    d_sdf = torch.randn(N, 1, generator=g).double()
    d_feat = 0.1 * torch.randn(N, 256, generator=g).double()
    d_n = torch.randn(N, 3, generator=g).double()
    L = (out[:, :1] * d_sdf).sum() + (out[:, 1:] * d_feat).sum() + (nrm * d_n).sum()
    names = list(spd)
    gr = torch.autograd.grad(L, [xg] + [spd[k] for k in names])
    gd = dict(zip(names, gr[1:]))
    Ws, bs = A.effective_weights(spd)
    sdf, feat, normal, st = A.sdf_obj_fwd(Ws, bs, x)
    assert (sdf - out[:, :1]).abs().max() < 1e-12 and (feat - out[:, 1:]).abs().max() < 1e-12
    assert (normal - nrm).abs().max() < 1e-7
    dx, dW, db = A.sdf_obj_bwd(Ws, bs, st, d_sdf, d_feat, d_n)
    assert ((dx - gr[0]).abs().max() / gr[0].abs().max()) < 1e-7
    for l in range(9):
        dg, dv = A.wn_backward(spd["lin%d.weight_v" % l], spd["lin%d.weight_g" % l], dW[l])
        for a, b in ((dg, gd["lin%d.weight_g" % l]), (dv, gd["lin%d.weight_v" % l]), (db[l], gd["lin%d.bias" % l])):
            assert ((a - b).abs().max() / b.abs().max()) < 1e-7, l


def test_operand_rounding_emulation_orders_precisions():
    """Single-pass 11-bit operands sit near 1e-3 on the SDF, BF16 above it, split operands far below:
    the measurement behind DESIGN.md section 4's precision plan."""
    sp, _ = synth.obj_states()
    spf = {k: v for k, v in sp.items() if k != "se3_refine"}
    x = 0.45 * torch.randn(256, 3, generator=torch.Generator().manual_seed(1))
    Ws, bs = A.effective_weights(spf)
    ref = A.sdf_obj_fwd([w.double() for w in Ws], [b.double() for b in bs], x.double())[0]
    err = {}
    for name, mm in (("tf32", A.tf32_mm), ("bf16", A.bf16_mm), ("bf16x3", A.bf16x3_mm)):
        err[name] = float((A.sdf_obj_fwd(Ws, bs, x, mm=mm)[0].double() - ref).abs().max())
    assert err["bf16x3"] < 1e-4 < err["tf32"] < err["bf16"]
    assert err["bf16"] > 1e-3
