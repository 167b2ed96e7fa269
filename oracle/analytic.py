"""CPU model of the *kernel algorithm*: SDF value + feature + analytic normal as ONE operator with a
hand-derived second-order backward, written with explicit
matrix products so that (a) the formulas the CUDA kernels implement are validated against autograd
on CPU before any GPU time is spent, and (b) the effect of tensor-core operand rounding
(TF32 / BF16 / split-BF16) can be emulated by swapping the ``mm`` callable.

TEST INFRASTRUCTURE ONLY (same rules as oracle/honerf_oracle.py).  Restates utils/fields.py:316-347
(object SDF forward + gradient) and the backward autograd derives for it.
"""
import math

import torch

SQRT1_2 = 1.0 / math.sqrt(2.0)


def exact_mm(a, b):
    return a @ b


def round_tf32(x):
    """round-to-nearest-even to 10 explicit mantissa bits (what cvt.rna.tf32.f32 produces)."""
    i = x.contiguous().view(torch.int32)
    r = ((i + 0x0FFF + ((i >> 13) & 1)) & ~0x1FFF)
    return r.view(torch.float32)


def trunc_tf32(x):
    return (x.contiguous().view(torch.int32) & ~0x1FFF).view(torch.float32)


def tf32_mm(a, b):
    return round_tf32(a) @ round_tf32(b)


def bf16_mm(a, b):
    return a.bfloat16().float() @ b.bfloat16().float()


def bf16x3_mm(a, b):
    ah = a.bfloat16().float()
    al = (a - ah).bfloat16().float()
    bh = b.bfloat16().float()
    bl = (b - bh).bfloat16().float()
    return ah @ bh + (ah @ bl + al @ bh)


def bf16x2_mm(a, b):
    """activation split only (a_hi + a_lo) x b_hi"""
    ah = a.bfloat16().float()
    al = (a - ah).bfloat16().float()
    bh = b.bfloat16().float()
    return ah @ bh + al @ bh


def enc_obj(x, L=10):
    """[x, channel-major sin/cos] (utils/fields.py:13-20, 318-319) and the sin/cos themselves."""
    freq = (2.0 ** torch.arange(L, dtype=torch.float32)).to(x.dtype)
    spec = x[:, :, None] * freq                       # [N,3,L]
    s, c = spec.sin(), spec.cos()
    e = torch.cat([x, torch.stack([s, c], dim=2).reshape(x.shape[0], -1)], dim=1)
    return e, s, c, freq


def enc_jt(g, s, c, freq):
    """J_e^T g : [N,3+6L] -> [N,3]."""
    N, _, L = s.shape
    ge = g[:, 3:].reshape(N, 3, 2, L)
    return g[:, :3] + (freq * (c * ge[:, :, 0] - s * ge[:, :, 1])).sum(-1)


def enc_j(t, s, c, freq):
    """J_e t : [N,3] -> [N,3+6L]."""
    N, _, L = s.shape
    ds = freq * c * t[:, :, None]
    dc = -freq * s * t[:, :, None]
    return torch.cat([t, torch.stack([ds, dc], dim=2).reshape(N, -1)], dim=1)


def enc_hess(g, t, s, c, freq):
    """sum_k d2 e_k/dx2 * g_k * t  (diagonal per coordinate): [N,3]."""
    N, _, L = s.shape
    ge = g[:, 3:].reshape(N, 3, 2, L)
    return t * (freq * freq * (-s * ge[:, :, 0] - c * ge[:, :, 1])).sum(-1)


def effective_weights(p, n_lin=9):
    Ws, bs = [], []
    for l in range(n_lin):
        v, g = p["lin%d.weight_v" % l], p["lin%d.weight_g" % l]
        Ws.append(v * (g / v.norm(dim=1, keepdim=True)))
        bs.append(p["lin%d.bias" % l])
    return Ws, bs


def sp_prime_from_h(h, beta=100.0):
    """softplus'(z) = sigmoid(beta z) = 1 - exp(-beta h) with h = softplus(z)."""
    return -torch.expm1(-beta * h)


def sdf_obj_fwd(Ws, bs, x, scale=1.0, mm=exact_mm, skip=4, beta=100.0):
    """Returns sdf [N,1], feat [N,256], normal [N,3] and the stash the backward needs."""
    e, s, c, freq = enc_obj(x)
    H, a_in = [], []
    a = e
    n_lin = len(Ws)
    for l in range(n_lin):
        if l == skip:
            a = torch.cat([a, e], dim=1) * SQRT1_2
        a_in.append(a)
        z = mm(a, Ws[l].t()) + bs[l]
        if l < n_lin - 1:
            a = torch.nn.functional.softplus(z, beta=beta)
            H.append(a)
    sdf = z[:, :1] / scale
    feat = z[:, 1:]
    # normal sweep (reverse mode with the one-hot seed on the sdf column)
    D = [None] * (n_lin - 1)
    hb = (Ws[n_lin - 1][0] / scale)[None, :].expand(x.shape[0], -1)
    eb_skip = None
    for l in range(n_lin - 2, -1, -1):
        D[l] = sp_prime_from_h(H[l], beta) * hb
        ab = mm(D[l], Ws[l])
        if l == skip:
            n_h = Ws[l].shape[1] - e.shape[1]
            eb_skip = ab[:, n_h:] * SQRT1_2
            hb = ab[:, :n_h] * SQRT1_2
        else:
            hb = ab
    eb = hb + eb_skip
    normal = enc_jt(eb, s, c, freq)
    stash = dict(e=e, s=s, c=c, freq=freq, H=H, D=D, a_in=a_in, eb=eb)
    return sdf, feat, normal, stash


def sdf_obj_bwd(Ws, bs, stash, d_sdf, d_feat, d_normal, scale=1.0, mm=exact_mm, skip=4,
                beta=100.0, need_dx=True):
    """Second-order backward: returns (d_x, dW list, db list)."""
    e, s, c, freq, H, D, a_in = (stash[k] for k in ("e", "s", "c", "freq", "H", "D", "a_in"))
    n_lin = len(Ws)
    # tangent sweep along d_normal
    ue = enc_j(d_normal, s, c, freq)
    u = ue
    au_in, X = [], []
    for l in range(n_lin - 1):
        if l == skip:
            u = torch.cat([u, ue], dim=1) * SQRT1_2
        au_in.append(u)
        q = mm(u, Ws[l].t())
        sp1 = sp_prime_from_h(H[l], beta)
        u = sp1 * q
        X.append(beta * (1.0 - sp1) * D[l] * q)       # s''(z) * hb * q with D = s' * hb
    u_last = u                                        # tangent entering the output layer
    # reverse sweep
    dz = torch.cat([d_sdf / scale, d_feat], dim=1)
    dW, db = [None] * n_lin, [None] * n_lin
    de_skip = None
    for l in range(n_lin - 1, -1, -1):
        dW[l] = mm(dz.t(), a_in[l])
        if l < n_lin - 1:
            dW[l] = dW[l] + mm(D[l].t(), au_in[l])
        else:
            # the normal sweep is seeded with row 0 of the output layer: <dn, n> = W_out[0].u / scale
            dW[l] = dW[l].clone()
            dW[l][0] += u_last.sum(0) / scale
        db[l] = dz.sum(0)
        da = mm(dz, Ws[l])
        if l == skip:
            n_h = Ws[l].shape[1] - e.shape[1]
            de_skip = da[:, n_h:] * SQRT1_2
            da = da[:, :n_h] * SQRT1_2
        if l > 0:
            dz = sp_prime_from_h(H[l - 1], beta) * da + X[l - 1]
    de = da + de_skip
    d_x = None
    if need_dx:
        d_x = enc_jt(de, s, c, freq) + enc_hess(stash["eb"], d_normal, s, c, freq)
    return d_x, dW, db


def wn_backward(v, g, dW):
    """(dg, dv) of W = g v / ||v||_row (SURVEY E-2)."""
    n = v.norm(dim=1, keepdim=True)
    dot = (dW * v).sum(1, keepdim=True)
    dg = dot / n
    dv = (g / n) * (dW - dot * v / (n * n))
    return dg, dv


# ------------------------------------------------------------------------------------------------
# HALO per-bone feature (utils/fields.py:22-36, 142-148) and its derivatives, per (point, joint):
#   q = R x + t - T,  v = |q|,  r = q / v,  h = 1 - sigmoid(200 (v - cutoff))
#   F(q) = h * [v, sin(2^k v), cos(2^k v) (k<10), r (3), sin(2^k r_a), cos(2^k r_a) (k<7, a<3)]   (66)
# For a cotangent c (66) define s(q) = <c, F(q)> = h(v) (A(v) + B(r)).  The CUDA kernels need
#   grad   g  = d s / d q            (normal sweep, first-order backward)
#   jvp    dF = F'(q) w              (tangent sweep)
#   hvp    Hw = d/dq (g . w)         (second-order backward: d<dn, normal>/d(x, R, t, T))
# ------------------------------------------------------------------------------------------------
HALO_CUTOFF = (0.08, 0.03, 0.03, 0.02, 0.02, 0.03, 0.02, 0.02, 0.02, 0.03, 0.02, 0.02, 0.02, 0.03,
               0.02, 0.02, 0.02, 0.03, 0.02, 0.02, 0.02)


def _halo_parts(q, cutoff, Lv=10, Lr=7, tau=200.0):
    """q [..,3], cutoff [..] -> dict of the shared scalars."""
    v = q.norm(dim=-1)
    r = q / v[..., None]
    sg = torch.sigmoid(tau * (v - cutoff))
    h = 1.0 - sg
    h1 = -tau * sg * (1.0 - sg)
    h2 = -tau * tau * sg * (1.0 - sg) * (1.0 - 2.0 * sg)
    fv = (2.0 ** torch.arange(Lv, dtype=q.dtype))
    fr = (2.0 ** torch.arange(Lr, dtype=q.dtype))
    av = v[..., None] * fv                    # [.., Lv]
    ar = r[..., None] * fr                    # [.., 3, Lr]
    return dict(v=v, r=r, h=h, h1=h1, h2=h2, fv=fv, fr=fr, sv=av.sin(), cv=av.cos(), sr=ar.sin(), cr=ar.cos())


def halo_feature_block(q, cutoff):
    p = _halo_parts(q, cutoff)
    phi = torch.cat([p["v"][..., None], p["sv"], p["cv"], p["r"],
                     torch.cat([p["sr"], p["cr"]], dim=-1).flatten(-2)], dim=-1)   # [.., 66]
    return phi * p["h"][..., None], p, phi


def _split_c(c, Lv=10, Lr=7):
    c0 = c[..., 0]
    cs, cc = c[..., 1:1 + Lv], c[..., 1 + Lv:1 + 2 * Lv]
    cr = c[..., 1 + 2 * Lv:4 + 2 * Lv]
    ang = c[..., 4 + 2 * Lv:].reshape(*c.shape[:-1], 3, 2 * Lr)
    return c0, cs, cc, cr, ang[..., :Lr], ang[..., Lr:]


def halo_grad_hvp(q, cutoff, c, w=None):
    """Returns g = d<c,F>/dq [..,3] and, when w is given, (jvp [..,66], hvp [..,3])."""
    F, p, phi = halo_feature_block(q, cutoff)
    c0, cs, cc, cr, cas, cac = _split_c(c)
    v, r, h, h1, h2 = p["v"], p["r"], p["h"], p["h1"], p["h2"]
    fv, fr = p["fv"], p["fr"]
    A = c0 * v + (cs * p["sv"] + cc * p["cv"]).sum(-1)
    A1 = c0 + (fv * (cs * p["cv"] - cc * p["sv"])).sum(-1)
    A2 = (fv * fv * (-cs * p["sv"] - cc * p["cv"])).sum(-1)
    B = (cr * r).sum(-1) + (cas * p["sr"] + cac * p["cr"]).sum((-1, -2))
    b1 = cr + (fr * (cas * p["cr"] - cac * p["sr"])).sum(-1)            # [..,3]
    b2 = (fr * fr * (-cas * p["sr"] - cac * p["cr"])).sum(-1)          # [..,3]
    G = A + B
    rb1 = (r * b1).sum(-1)
    Pb1 = b1 - r * rb1[..., None]
    g = (h1 * G + h * A1)[..., None] * r + (h / v)[..., None] * Pb1
    if w is None:
        return g
    dv = (r * w).sum(-1)
    dr = (w - r * dv[..., None]) / v[..., None]
    # jvp of the 66 features
    dphi = torch.cat([dv[..., None], fv * p["cv"] * dv[..., None], -fv * p["sv"] * dv[..., None], dr,
                      torch.cat([fr * p["cr"] * dr[..., None], -fr * p["sr"] * dr[..., None]], dim=-1).flatten(-2)],
                     dim=-1)
    jvp = dphi * h[..., None] + phi * (h1 * dv)[..., None]
    # hvp
    dG = A1 * dv + (b1 * dr).sum(-1)
    db1 = b2 * dr
    dPb1 = -dr * rb1[..., None] - r * (dr * b1).sum(-1)[..., None] + (db1 - r * (r * db1).sum(-1)[..., None])
    hvp = (h2 * dv * G + h1 * dG + h1 * dv * A1 + h * A2 * dv)[..., None] * r \
        + (h1 * G + h * A1)[..., None] * dr \
        + (h1 * dv / v)[..., None] * Pb1 + (h / v)[..., None] * dPb1 - (h * dv / (v * v))[..., None] * Pb1
    return g, jvp, hvp


# --------------------------------------------------------------------------------------------
# Closed forms used by csrc/loss.cu and csrc/rays.cu (no autograd), checked against the reference's golden vectors
# on the CPU before the kernels are trusted on the GPU (tests/test_oracle_8f.py).
# --------------------------------------------------------------------------------------------
def render_loss_closed_form(color, wsum, true_rgb, mask, grad_err, color_div, color_w, mask_w, igr_w, g=1.0):
    """Forward sums and the backward formulas of render_loss_fwd/bwd_kernel."""
    n = color.shape[0]
    m = mask.reshape(n, 1)
    err = (color - true_rgb) * m
    mask_sum = m.sum() + 1e-5
    div = color_div if color_div > 0 else mask_sum
    p = wsum.reshape(n, 1).clip(1e-3, 1.0 - 1e-3)
    bce = (m - 1.0) * torch.log1p(-p).clamp_min(-100.0) - m * torch.log(p).clamp_min(-100.0)
    color_loss = err.abs().sum() / div
    mask_loss = bce.sum() / n
    ge = grad_err if grad_err is not None else 0.0
    total = color_w * color_loss + mask_w * mask_loss + igr_w * ge
    d_color = g * color_w / div * torch.sign(err) * m
    w = wsum.reshape(n, 1)
    passes = ((w >= 1e-3) & (w <= 1.0 - 1e-3)).to(color.dtype)
    d_wsum = passes * g * mask_w * (p - m) / ((1.0 - p) * p).clamp_min(1e-12) / n
    return total, color_loss, mask_loss, d_color, d_wsum, g * igr_w


def interaction_closed_form(sdf_h, sdf_o, thr, w_c, w_p, g=1.0):
    h, o = sdf_h[:, 0], sdf_o[:, 0]
    a = h.abs() + o.abs()
    contact = a < thr
    pen = (o < 0) & (h < 0)
    cnum = contact.to(h.dtype).sum() + 1e-9
    pnum = pen.to(h.dtype).sum() + 1e-9
    c_loss, p_loss = (a * contact).sum() / cnum, (a * pen).sum() / pnum
    k = g * (contact.to(h.dtype) * w_c / cnum + pen.to(h.dtype) * w_p / pnum)
    return w_c * c_loss + w_p * p_loss, c_loss, p_loss, k * torch.sign(h), k * torch.sign(o)


def rays_closed_form(R, T, focal, pp, xy):
    """csrc/rays.cu: adjugate inverse of R, per-ray un-projection of the depth-1 / depth-2 points (single camera)."""
    r = R.reshape(9)
    c00, c01, c02 = r[4] * r[8] - r[5] * r[7], r[5] * r[6] - r[3] * r[8], r[3] * r[7] - r[4] * r[6]
    det = r[0] * c00 + r[1] * c01 + r[2] * c02
    inv = torch.stack([c00, r[2] * r[7] - r[1] * r[8], r[1] * r[5] - r[2] * r[4],
                       c01, r[0] * r[8] - r[2] * r[6], r[2] * r[3] - r[0] * r[5],
                       c02, r[1] * r[6] - r[0] * r[7], r[0] * r[4] - r[1] * r[3]]).reshape(3, 3) / det

    def unproject(depth):
        v = torch.stack([(xy[..., 0] - pp[0]) / focal[0] * depth - T[0], (xy[..., 1] - pp[1]) / focal[1] * depth - T[1],
                         torch.full_like(xy[..., 0], depth) - T[2]], dim=-1)
        return v @ inv
    p1, p2 = unproject(1.0), unproject(2.0)
    d = p2 - p1
    d = d / d.norm(dim=-1, keepdim=True).clamp_min(1e-12)
    return p1 - d, d


def nn_select_bruteforce(pts, in_mask, out_mask):
    """CPU model of csrc/neighbors.cu: fp64 distances, lowest index on exact ties."""
    T, P = in_mask.shape
    d2 = ((pts.double()[:, None, :] - pts.double()[None, :, :]) ** 2).sum(-1)          # [P(p), P(q)]
    flag = torch.zeros(T, P, dtype=torch.bool)
    nearest = torch.full((T, P), -1, dtype=torch.int64)
    for t in range(T):
        if not out_mask[t].any():
            continue
        dd = d2.clone()
        dd[:, ~out_mask[t]] = float("inf")
        q = dd.argmin(dim=1)                       # torch argmin returns the first minimum
        nearest[t, in_mask[t]] = q[in_mask[t]]
        flag[t, q[in_mask[t]]] = True
    return flag, nearest


def stable_loss_closed_form(hand_sdf, pts0, fixed=False, nn=None):
    """The device formulation of honerf_b200.ops.stable_loss_from_sdf (masks + two matrix-vector products), with the
    nearest-neighbour flags from `nn` (default: the brute-force model above)."""
    nn = nn or (lambda p, i, o: nn_select_bruteforce(p, i, o)[0])
    F_, P = hand_sdf.shape
    neg = hand_sdf.detach() < 0
    valid = neg.any(dim=1)
    in_time = valid.sum()
    if fixed:
        out_mask = ~neg
    else:
        out_mask = torch.ones_like(neg)
        out_mask[:, 0] = ~(~neg).any(dim=1)
        out_mask[:, 1] = ~neg.any(dim=1)
    flag = nn(pts0, neg & valid[:, None], out_mask)
    vf = valid.to(hand_sdf.dtype)
    s_pos = (hand_sdf.clip(0, 1e7) * vf[:, None]).sum(0)
    s_neg = (hand_sdf.clip(-1e7, 0).abs() * vf[:, None]).sum(0)
    n_in = neg.sum(dim=1).to(hand_sdf.dtype)
    denom = ((in_time - 1).to(hand_sdf.dtype) * n_in).clamp_min(1.0)
    in_err = (neg.to(hand_sdf.dtype) @ s_pos) / denom
    out_err = (flag.to(hand_sdf.dtype) @ s_neg) / denom
    total = ((in_err + 0.05 * out_err) * vf).sum() / in_time.clamp_min(1).to(hand_sdf.dtype)
    return torch.where(in_time > 1, total, torch.zeros_like(total))
