"""CPU emulation of the operand formats the chain kernels could run the object SDF field in: which of
   {three, two, one} 16-bit MMAs per product, and which stash width (fp32 pair / single 16-bit), stay inside the north
   star's bounds (sdf / colour 1e-3 abs, gradients 1e-2 relative).  Uses oracle/analytic.py's formulas with the
   contraction and the stash rounding swapped per sweep.  Prints one row per variant; profiles/r02_precision_table.md is
   the committed copy, next to the GPU measurements of the variants that were built.

   python tools/precision_table.py [n_points]
"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import analytic as A   # noqa: E402
import synth           # noqa: E402


def r16(x, kind):
    return x.half().float() if kind == "f16" else x.bfloat16().float()


def make_mm(kind, passes):
    """passes: 'x3' hi*hi + lo*hi + hi*lo | 'a2' (a_hi + a_lo) * b_hi | 'b2' a_hi * (b_hi + b_lo) | 'x1' | 'exact'"""
    if passes == "exact":
        return A.exact_mm

    def mm(a, b):
        ah = r16(a, kind)
        bh = r16(b, kind)
        if passes == "x1":
            return ah @ bh
        al = r16(a - ah, kind)
        bl = r16(b - bh, kind)
        if passes == "a2":
            return ah @ bh + al @ bh
        if passes == "b2":
            return ah @ bh + ah @ bl
        return ah @ bh + (al @ bh + ah @ bl)
    return mm


def run(Ws, bs, x, cot, cfg):
    """cfg: trunk=(kind, passes), normal=..., tangent=..., reverse=..., dw=..., store_h / store_c = None | 'f16' | 'bf16'"""
    d_sdf, d_feat, d_n = cot
    mm_t, mm_n = make_mm(*cfg["trunk"]), make_mm(*cfg["normal"])
    sh = (lambda t: t) if cfg.get("store_h") is None else (lambda t: r16(t, cfg["store_h"]))
    sc = (lambda t: t) if cfg.get("store_c") is None else (lambda t: r16(t, cfg["store_c"]))
    # forward: the trunk and the normal sweep use different contractions -> run analytic's forward twice is wasteful;
    # restate with two callables
    e, s, c, freq = A.enc_obj(x)
    H, a_in = [], []
    a = e
    for l in range(9):
        if l == 4:
            a = torch.cat([a, e], 1) * A.SQRT1_2
        a_in.append(a)
        z = mm_t(a, Ws[l].t()) + bs[l]
        if l < 8:
            a = torch.nn.functional.softplus(z, beta=100.0)
            H.append(a)
    sdf, feat = z[:, :1], z[:, 1:]
    Hs = [sh(h) for h in H]                     # what the stash holds (the next layer's operand stays on chip, unrounded)
    Hn = H if cfg.get("h_pair_for_normal") else Hs   # the normal sweep may read a hi + lo pair of the same quantity
    D = [None] * 8
    hb = Ws[8][0][None, :].expand(x.shape[0], -1)
    for l in range(7, -1, -1):
        D[l] = A.sp_prime_from_h(Hn[l]) * hb
        ab = mm_n(D[l], Ws[l])
        if l == 4:
            eb_skip, hb = ab[:, 193:] * A.SQRT1_2, ab[:, :193] * A.SQRT1_2
        else:
            hb = ab
    eb = hb + eb_skip
    normal = A.enc_jt(eb, s, c, freq)
    Ds = [sc(d) for d in D]
    # backward
    mm_u, mm_r, mm_w = make_mm(*cfg["tangent"]), make_mm(*cfg["reverse"]), make_mm(*cfg["dw"])
    ue = A.enc_j(d_n, s, c, freq)
    u = ue
    au_in, X = [], []
    for l in range(8):
        if l == 4:
            u = torch.cat([u, ue], 1) * A.SQRT1_2
        au_in.append(u)
        q = mm_u(u, Ws[l].t())
        sp1 = A.sp_prime_from_h(Hs[l])
        u = sp1 * q
        X.append(100.0 * (1.0 - sp1) * Ds[l] * q)
    u_last = u
    Xs = [sc(t) for t in X]
    Us = [sc(t) for t in au_in]
    a_s = [a_in[0]] + [sh(t) for t in a_in[1:]]
    dz = torch.cat([d_sdf, d_feat], 1)
    dW, db = [None] * 9, [None] * 9
    for l in range(8, -1, -1):
        dzs = sc(dz) if l < 8 else dz
        dW[l] = mm_w(dzs.t(), a_s[l])
        if l < 8:
            dW[l] = dW[l] + mm_w(Ds[l].t(), Us[l])
        else:
            dW[l] = dW[l].clone()
            dW[l][0] += u_last.sum(0)
        db[l] = dz.sum(0)
        da = mm_r(dz, Ws[l])
        if l == 4:
            de_skip, da = da[:, 193:] * A.SQRT1_2, da[:, :193] * A.SQRT1_2
        if l > 0:
            dz = A.sp_prime_from_h(Hs[l - 1]) * da + Xs[l - 1]
    de = da + de_skip
    d_x = A.enc_jt(de, s, c, freq) + A.enc_hess(eb, d_n, s, c, freq)
    return sdf, feat, normal, d_x, dW, db


def rel_l2(a, b):
    return float((a.double() - b.double()).norm() / (b.double().norm() + 1e-300))


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
    torch.manual_seed(0)
    sp, _ = synth.obj_states()
    spd = {k: v.double() for k, v in sp.items() if k != "se3_refine"}
    Wd, bd = A.effective_weights(spd)
    Wf, bf = [w.float() for w in Wd], [b.float() for b in bd]
    g = torch.Generator().manual_seed(5)
    x = 0.45 * torch.randn(n, 3, generator=g)
    cot = (torch.randn(n, 1, generator=g), 0.1 * torch.randn(n, 256, generator=g), torch.randn(n, 3, generator=g))
    ex = ("f16", "exact")
    ref = run(Wd, bd, x.double(), tuple(c.double() for c in cot),
              dict(trunk=ex, normal=ex, tangent=ex, reverse=ex, dw=ex))
    F3, F2a, F2b, F1 = ("f16", "x3"), ("f16", "a2"), ("f16", "b2"), ("f16", "x1")
    B3, B2a, B2b, B1 = ("bf16", "x3"), ("bf16", "a2"), ("bf16", "b2"), ("bf16", "x1")
    variants = [
        ("r01: trunk f16x3, sweeps bf16x3, fp32 stash", dict(trunk=F3, normal=B3, tangent=B3, reverse=B3, dw=B3)),
        ("trunk f16 a-split (weights 1x f16)", dict(trunk=F2a, normal=B3, tangent=B3, reverse=B3, dw=B3)),
        ("trunk f16 b-split (activations 1x f16)", dict(trunk=F2b, normal=B3, tangent=B3, reverse=B3, dw=B3)),
        ("trunk f16 x1", dict(trunk=F1, normal=B3, tangent=B3, reverse=B3, dw=B3)),
        ("normal sweep bf16 b-split", dict(trunk=F3, normal=B2b, tangent=B3, reverse=B3, dw=B3)),
        ("normal sweep bf16 x1", dict(trunk=F3, normal=B1, tangent=B3, reverse=B3, dw=B3)),
        ("normal sweep f16 b-split (scaled cotangent)", dict(trunk=F3, normal=F2b, tangent=B3, reverse=B3, dw=B3)),
        ("tangent+reverse bf16 b-split", dict(trunk=F3, normal=B3, tangent=B2b, reverse=B2b, dw=B3)),
        ("tangent+reverse bf16 x1", dict(trunk=F3, normal=B3, tangent=B1, reverse=B1, dw=B3)),
        ("dw bf16 a-split", dict(trunk=F3, normal=B3, tangent=B3, reverse=B3, dw=B2a)),
        ("dw bf16 x1", dict(trunk=F3, normal=B3, tangent=B3, reverse=B3, dw=B1)),
        ("stash H f16, cotangents bf16 (x3 everywhere)", dict(trunk=F3, normal=B3, tangent=B3, reverse=B3, dw=B3, store_h="f16", store_c="bf16")),
        ("stash H f16, cotangents f16", dict(trunk=F3, normal=B3, tangent=B3, reverse=B3, dw=B3, store_h="f16", store_c="f16")),
        ("ALL b-split, 16-bit stash, dw x1", dict(trunk=F2b, normal=B2b, tangent=B2b, reverse=B2b, dw=B1, store_h="f16", store_c="bf16")),
        ("trunk x3; sweeps b-split, 16-bit stash, dw x1", dict(trunk=F3, normal=B2b, tangent=B2b, reverse=B2b, dw=B1, store_h="f16", store_c="bf16")),
        ("trunk b-split; sweeps x1, 16-bit stash, dw x1", dict(trunk=F2b, normal=B1, tangent=B1, reverse=B1, dw=B1, store_h="f16", store_c="bf16")),
        ("r02 default: x3 everywhere, 16-bit stash (normal sweep reads s' as a hi+lo pair), dw on stored bf16", dict(trunk=F3, normal=F3, tangent=B3, reverse=B3, dw=B1, store_h="f16", store_c="bf16", h_pair_for_normal=True)),
        ("ALL x1, 16-bit stash", dict(trunk=F1, normal=B1, tangent=B1, reverse=B1, dw=B1, store_h="f16", store_c="bf16")),
    ]
    print("| variant | sdf max-abs | feat max-abs | normal rel-L2 | d_pts rel-L2 | worst dW rel-L2 | worst db rel-L2 |")
    print("|---|---:|---:|---:|---:|---:|---:|")
    for name, cfg in variants:
        got = run(Wf, bf, x, cot, cfg)
        e_sdf = float((got[0].double() - ref[0]).abs().max())
        e_feat = float((got[1].double() - ref[1]).abs().max())
        e_n = rel_l2(got[2], ref[2])
        e_x = rel_l2(got[3], ref[3])
        e_w = max(rel_l2(a, b) for a, b in zip(got[4], ref[4]))
        e_b = max(rel_l2(a, b) for a, b in zip(got[5], ref[5]))
        print("| %s | %.1e | %.1e | %.1e | %.1e | %.1e | %.1e |" % (name, e_sdf, e_feat, e_n, e_x, e_w, e_b))


if __name__ == "__main__":
    main()
