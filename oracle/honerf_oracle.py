"""CPU oracle: an independent restatement of HO-NeRF's NeuS volume-rendering hot path.

TEST INFRASTRUCTURE ONLY.  Only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` /
``--impl reference`` legs of ``bench.py`` may import this module, and only as the checker (or as the
timed CPU baseline) -- never from the product package ``honerf_b200``.

It is written from the behaviour specification in SURVEY.md (section 8a, appendices A-E) as plain
functional PyTorch on CPU (fp32, or fp64 for gradient checks) over ``state_dict``-style parameter
mappings.  Every function cites the reference ``file:line`` it restates.  It is pinned against the
real reference two ways (see ``oracle/make_golden.py`` and ``tests/test_oracle_golden.py``):
golden input/output vectors produced by importing ``/root/reference`` in the build container are
committed under ``tests/golden/``; and, when the reference tree is present, the oracle is compared
live against the reference modules.

Parity status: PINNED (golden vectors generated from the reference's own code; the reference ships
no tests or fixtures of its own -- SURVEY.md section 4).  The SURVEY 8f rows (losses, contact stability, ray
bundles) are pinned the same way by ``oracle/make_golden_8f.py`` / ``tests/test_oracle_8f.py``, with ONE exception that
is PARITY UNPINNED: ``unproject_ndc`` restates the NDC camera model of pytorch3d, an un-vendored and un-pinned
dependency of the reference that is not importable here (the ray-bundle construction around it is the reference's own
function and is pinned).
"""
from __future__ import annotations

import math
from typing import Dict, Mapping, Optional, Sequence

import torch
import torch.nn.functional as F

Tensor = torch.Tensor
Params = Mapping[str, Tensor]

# utils/fields.py:24 and :40 -- per-joint cut-off distances of the HALO embedding
HALO_CUTOFF = (0.08, 0.03, 0.03, 0.02, 0.02, 0.03, 0.02, 0.02, 0.02, 0.03, 0.02, 0.02, 0.02, 0.03,
               0.02, 0.02, 0.02, 0.03, 0.02, 0.02, 0.02)
HALO_TAU = 200.0
SQRT2 = math.sqrt(2.0)


# --------------------------------------------------------------------------------------------
# encodings
# --------------------------------------------------------------------------------------------
def embed(x: Tensor, L: int) -> Tensor:
    """BARF-style encoding without pi, channel-major (utils/fields.py:13-20).

    out[..., n*2L + s*L + k] = (sin if s == 0 else cos)(2^k * x[..., n])
    """
    freq = 2.0 ** torch.arange(L, dtype=torch.float32, device=x.device)
    freq = freq.to(x.dtype)
    spec = x[..., None] * freq                      # [..., C, L]
    enc = torch.stack([spec.sin(), spec.cos()], dim=-2)   # [..., C, 2, L]
    return enc.reshape(*x.shape[:-1], -1)


def embed_with_input(x: Tensor, L: int) -> Tensor:
    """cat([x, embed(x)]) as every caller does (utils/fields.py:318-319, 389-394)."""
    return torch.cat([x, embed(x, L)], dim=-1)


def halo_embed(pts: Tensor, bt_inv: Tensor, T_pose_21: Tensor):
    """A-NeRF/HALO per-bone coordinates (utils/fields.py:22-36; batched :38-52).

    pts [N,3] with bt_inv [21,4,4], T_pose_21 [21,3]; or pts [F,P,3] with bt_inv [F,21,4,4],
    T_pose_21 [F,21,3].  Returns v [..,21,1], r [..,21,3], h [..,21,1].
    """
    cutoff = torch.tensor(HALO_CUTOFF, dtype=torch.float32, device=pts.device).to(pts.dtype)
    R = bt_inv[..., :3, :3]
    t = bt_inv[..., :3, 3]
    # same broadcast matmul as the reference: q = R x + t - T suffers cancellation (|R x + t| ~ 1 m,
    # |q| ~ cm), so a different contraction order already moves the 2^9-frequency encodings by 1e-5
    if pts.dim() == 2:
        q = torch.matmul(R[None], pts[:, None, :, None])[..., 0] + t[None]
        q = q - T_pose_21[None]
    else:
        q = torch.matmul(R[:, None], pts[:, :, None, :, None])[..., 0] + t[:, None]
        q = q - T_pose_21[:, None]
    v = torch.linalg.norm(q, dim=-1, keepdim=True)
    r = q / v
    h = 1.0 - torch.sigmoid(HALO_TAU * (v - cutoff[:, None]))
    return v, r, h


def halo_feature(pts: Tensor, bt_inv: Tensor, T_pose_21: Tensor, v_multires=10, r_multires=7):
    """The 21*66 = 1386-wide hand input feature (utils/fields.py:142-148, SURVEY A-7)."""
    v, r, h = halo_embed(pts, bt_inv, T_pose_21)
    v = v.reshape(-1, 21, 1)
    r = r.reshape(-1, 21, 3)
    h = h.reshape(-1, 21, 1)
    feat = torch.cat([v, embed(v, v_multires), r, embed(r, r_multires)], dim=-1) * h
    return feat.flatten(-2, -1), r, h


# --------------------------------------------------------------------------------------------
# networks (functional, over state_dict-style mappings)
# --------------------------------------------------------------------------------------------
def wn_weight(p: Params, l: int) -> Tensor:
    """Effective weight of a weight-normalised Linear: g * v / ||v||_row
    (nn.utils.weight_norm, dim=0; utils/fields.py:120-121, 216-217, 307-308, 382-383)."""
    v = p["lin%d.weight_v" % l]
    g = p["lin%d.weight_g" % l]
    # the hook installed by nn.utils.weight_norm calls exactly this ATen op; using it (rather than
    # v * (g / ||v||) spelled out) keeps the oracle's weights bit-identical to the reference's
    return torch._weight_norm(v, g, 0)


def _linear(p: Params, l: int, x: Tensor) -> Tensor:
    return F.linear(x, wn_weight(p, l), p["lin%d.bias" % l])


def _num_linear(p: Params) -> int:
    n = 0
    while ("lin%d.weight_v" % n) in p:
        n += 1
    return n


def _sdf_trunk(p: Params, inputs: Tensor, skip_in: Sequence[int]) -> Tensor:
    n_lin = _num_linear(p)
    x = inputs
    for l in range(n_lin):
        if l in skip_in:
            x = torch.cat([x, inputs], dim=1) / SQRT2
        x = _linear(p, l, x)
        if l < n_lin - 1:
            x = F.softplus(x, beta=100)
    return x


def sdf_obj_forward(p: Params, x: Tensor, multires=10, skip_in=(4,), scale=1.0) -> Tensor:
    """SDFNetwork_OBJ.forward (utils/fields.py:316-328): [N,3] -> [N,257] = [sdf/scale, feature]."""
    inputs = embed_with_input(x, multires)
    out = _sdf_trunk(p, inputs, skip_in)
    return torch.cat([out[:, :1] / scale, out[:, 1:]], dim=-1)


def sdf_hand_forward(p: Params, x: Tensor, bt_inv: Tensor, T_pose_21: Tensor, v_multires=10,
                     r_multires=7, skip_in=(4,)):
    """SDFNetwork.forward (utils/fields.py:132-156): returns (out [N,257], xyz_feature [N,1386],
    r, h).  No division by ``scale`` here (SURVEY A-9)."""
    feat, r, h = halo_feature(x, bt_inv, T_pose_21, v_multires, r_multires)
    out = _sdf_trunk(p, feat, skip_in)
    return out, feat, r, h


def sdf_gradient(sdf_fn, x: Tensor, create_graph=True) -> Tensor:
    """d sdf / d x through autograd, graph kept (utils/fields.py:165-177, 336-347). [N,3]."""
    if not x.requires_grad:
        x.requires_grad_(True)
    with torch.enable_grad():
        y = sdf_fn(x)
        (g,) = torch.autograd.grad(y, x, torch.ones_like(y), create_graph=create_graph,
                                   retain_graph=True)
    return g


def color_obj_forward(p: Params, pts, dirs, feat, normals, v_multires=10, r_multires=4,
                      grad_multires=4) -> Tensor:
    """RenderingNetwork_OBJ.forward (utils/fields.py:387-405)."""
    x = torch.cat([embed_with_input(pts, v_multires), embed_with_input(dirs, r_multires), feat,
                   embed_with_input(normals, grad_multires)], dim=-1)
    n_lin = _num_linear(p)
    for l in range(n_lin):
        x = _linear(p, l, x)
        if l < n_lin - 1:
            x = F.relu(x)
    return torch.sigmoid(x)


def color_hand_forward(p: Params, xyz_feature, feat, normals, grad_multires=4) -> Tensor:
    """RenderingNetwork.forward (utils/fields.py:222-240); d, h, index are ignored there."""
    x = torch.cat([xyz_feature, feat, embed_with_input(normals, grad_multires)], dim=-1)
    n_lin = _num_linear(p)
    for l in range(n_lin):
        x = _linear(p, l, x)
        if l < n_lin - 1:
            x = F.relu(x)
    return torch.sigmoid(x)


def inv_s_from_variance(variance: Tensor) -> Tensor:
    """SingleVarianceNetwork + the clip at the call site (utils/fields.py:248-249,
    utils/renderer.py:144)."""
    return torch.exp(variance * 10.0).clip(1e-6, 1e6)


# --------------------------------------------------------------------------------------------
# hierarchical sampling
# --------------------------------------------------------------------------------------------
def sample_pdf_cdf(weights: Tensor) -> Tensor:
    """cdf of sample_pdf (utils/renderer.py:13-16): [..., m-1] -> [..., m]."""
    w = weights + 1e-5
    pdf = w / torch.sum(w, -1, keepdim=True)
    cdf = torch.cumsum(pdf, -1)
    return torch.cat([torch.zeros_like(cdf[..., :1]), cdf], -1)


def inverse_cdf(bins: Tensor, cdf: Tensor, n_samples: int):
    """Deterministic inverse-CDF sampling (utils/renderer.py:18-35).  Returns (samples, below,
    above); indices are int64."""
    u = torch.linspace(0.5 / n_samples, 1.0 - 0.5 / n_samples, n_samples, device=cdf.device)
    u = u.to(cdf.dtype).expand(list(cdf.shape[:-1]) + [n_samples]).contiguous()
    inds = torch.searchsorted(cdf, u, right=True)
    below = torch.clamp(inds - 1, min=0)
    above = torch.clamp(inds, max=cdf.shape[-1] - 1)
    cdf_b = torch.gather(cdf, -1, below)
    cdf_a = torch.gather(cdf, -1, above)
    bin_b = torch.gather(bins, -1, below)
    bin_a = torch.gather(bins, -1, above)
    denom = cdf_a - cdf_b
    denom = torch.where(denom < 1e-5, torch.ones_like(denom), denom)
    t = (u - cdf_b) / denom
    return bin_b + t * (bin_a - bin_b), below, above


def up_sample_weights(z_vals: Tensor, sdf: Tensor, inv_s: float) -> Tensor:
    """Section weights of NeuSRenderer.up_sample (utils/renderer.py:64-83): [..., m] -> [..., m-1]."""
    prev_sdf, next_sdf = sdf[..., :-1], sdf[..., 1:]
    prev_z, next_z = z_vals[..., :-1], z_vals[..., 1:]
    mid_sdf = (prev_sdf + next_sdf) * 0.5
    cos_val = (next_sdf - prev_sdf) / (next_z - prev_z + 1e-5)
    prev_cos = torch.cat([torch.zeros_like(cos_val[..., :1]), cos_val[..., :-1]], dim=-1)
    cos_val = torch.minimum(prev_cos, cos_val).clip(-1e3, 0.0)
    dist = next_z - prev_z
    prev_est = mid_sdf - cos_val * dist * 0.5
    next_est = mid_sdf + cos_val * dist * 0.5
    prev_cdf = torch.sigmoid(prev_est * inv_s)
    next_cdf = torch.sigmoid(next_est * inv_s)
    alpha = (prev_cdf - next_cdf + 1e-5) / (prev_cdf + 1e-5)
    trans = torch.cumprod(torch.cat([torch.ones_like(alpha[..., :1]), 1.0 - alpha + 1e-7], -1), -1)
    return alpha * trans[..., :-1]


def up_sample(z_vals: Tensor, sdf: Tensor, n_importance: int, inv_s: float) -> Tensor:
    """NeuSRenderer.up_sample (utils/renderer.py:60-86)."""
    w = up_sample_weights(z_vals, sdf.reshape(z_vals.shape), inv_s)
    return inverse_cdf(z_vals, sample_pdf_cdf(w), n_importance)[0].detach()


def merge_sorted(z_vals: Tensor, new_z: Tensor):
    """The sort of cat_z_vals (utils/renderer.py:92-93). Returns (sorted z, permutation)."""
    return torch.sort(torch.cat([z_vals, new_z], dim=-1), dim=-1)


def ray_points(rays_o: Tensor, rays_d: Tensor, z: Tensor) -> Tensor:
    """o + d * z with separately rounded multiply and add (utils/renderer.py:91,124,216)."""
    return rays_o[..., None, :] + rays_d[..., None, :] * z[..., :, None]


def hierarchical_z(sdf_fn, rays_o, rays_d, z_vals, n_importance, up_sample_steps):
    """The no-grad loop of NeuSRenderer.render (utils/renderer.py:214-234).  ``sdf_fn`` maps
    [M,3] points to [M,1] sdf."""
    B, n = z_vals.shape
    with torch.no_grad():
        sdf = sdf_fn(ray_points(rays_o, rays_d, z_vals).reshape(-1, 3)).reshape(B, n)
        for i in range(up_sample_steps):
            new_z = up_sample(z_vals, sdf, n_importance // up_sample_steps, 64 * 2 ** i)
            last = i + 1 == up_sample_steps
            merged, index = merge_sorted(z_vals, new_z)
            if not last:
                new_sdf = sdf_fn(ray_points(rays_o, rays_d, new_z).reshape(-1, 3)).reshape(B, -1)
                sdf = torch.gather(torch.cat([sdf, new_sdf], dim=-1), -1, index)
            z_vals = merged
    return z_vals


# --------------------------------------------------------------------------------------------
# compositing
# --------------------------------------------------------------------------------------------
def mid_points(rays_o, rays_d, z_vals, sample_dist):
    """Section lengths and mid-point samples (utils/renderer.py:119-127)."""
    dists = z_vals[..., 1:] - z_vals[..., :-1]
    last = torch.full_like(dists[..., :1], float(torch.tensor(sample_dist, dtype=torch.float32)))
    dists = torch.cat([dists, last], -1)
    mid_z = z_vals + dists * 0.5
    pts = ray_points(rays_o, rays_d, mid_z)
    dirs = rays_d[..., None, :].expand(pts.shape)
    return dists, pts.reshape(-1, 3), dirs.reshape(-1, 3)


def neus_alpha(sdf, normals, dirs, dists, inv_s):
    """s-density alpha (utils/renderer.py:147-161).  sdf [N,1], normals/dirs [N,3], dists [B,n].
    Returns alpha [B,n] (clipped to [0,1]) and c = prev_cdf [B,n]."""
    true_cos = (dirs * normals).sum(-1, keepdim=True)
    iter_cos = -F.relu(-true_cos)                    # cos_anneal_ratio is hard-coded to 1.0
    d = dists.reshape(-1, 1)
    est_next = sdf + iter_cos * d * 0.5
    est_prev = sdf - iter_cos * d * 0.5
    prev_cdf = torch.sigmoid(est_prev * inv_s)
    next_cdf = torch.sigmoid(est_next * inv_s)
    alpha = ((prev_cdf - next_cdf + 1e-5) / (prev_cdf + 1e-5)).reshape(dists.shape).clip(0.0, 1.0)
    return alpha, prev_cdf.reshape(dists.shape)


def eikonal(normals: Tensor) -> Tensor:
    """mean((||n|| - 1)^2) over all samples (utils/renderer.py:166-169)."""
    return ((torch.linalg.norm(normals, ord=2, dim=-1) - 1.0) ** 2).mean()


def render_core_obj(sdf_p: Params, color_p: Params, variance: Tensor, rays_o, rays_d, z_vals,
                    sample_dist, scale=1.0) -> Dict[str, Tensor]:
    """NeuSRenderer.render_core, object branch (utils/renderer.py:107-177)."""
    B, n = z_vals.shape
    dists, pts, dirs = mid_points(rays_o, rays_d, z_vals, sample_dist)
    if not pts.requires_grad:
        pts.requires_grad_(True)
    out = sdf_obj_forward(sdf_p, pts, scale=scale)
    sdf, feat = out[:, :1], out[:, 1:]
    normals = sdf_gradient(lambda q: sdf_obj_forward(sdf_p, q, scale=scale)[:, :1], pts)
    rgb = color_obj_forward(color_p, pts, dirs, feat, normals).reshape(B, n, 3)
    inv_s = inv_s_from_variance(variance)
    alpha, c = neus_alpha(sdf, normals, dirs, dists, inv_s)
    # quirk D-1: transmittance is seeded with c[:, :1], not with ones (utils/renderer.py:163)
    trans = torch.cumprod(torch.cat([c[:, :1], 1.0 - alpha + 1e-7], -1), -1)[:, :-1]
    weights = alpha * trans
    color = (rgb * weights[:, :, None]).sum(dim=1)
    return dict(color=color, s_val=(1.0 / inv_s).expand(B * n, 1), weights=weights, cdf=c,
                gradient_error=eikonal(normals.reshape(B, n, 3)), sdf=sdf, normals=normals,
                rgb=rgb)


def render_core_hand(sdf_p: Params, color_p: Params, variance: Tensor, rays_o, rays_d, z_vals,
                     sample_dist, bt_inv, T_pose_21, r_multires=7) -> Dict[str, Tensor]:
    """NeuSRenderer.render_core, hand branch (utils/renderer.py:136-141 + :144-177)."""
    B, n = z_vals.shape
    dists, pts, dirs = mid_points(rays_o, rays_d, z_vals, sample_dist)
    if not pts.requires_grad:
        pts.requires_grad_(True)
    out, xyz_feature, _, _ = sdf_hand_forward(sdf_p, pts, bt_inv, T_pose_21, r_multires=r_multires)
    sdf, feat = out[:, :1], out[:, 1:]
    normals = sdf_gradient(
        lambda q: sdf_hand_forward(sdf_p, q, bt_inv, T_pose_21, r_multires=r_multires)[0][:, :1], pts)
    rgb = color_hand_forward(color_p, xyz_feature, feat, normals).reshape(B, n, 3)
    inv_s = inv_s_from_variance(variance)
    alpha, c = neus_alpha(sdf, normals, dirs, dists, inv_s)
    trans = torch.cumprod(torch.cat([c[:, :1], 1.0 - alpha + 1e-7], -1), -1)[:, :-1]
    weights = alpha * trans
    color = (rgb * weights[:, :, None]).sum(dim=1)
    return dict(color=color, s_val=(1.0 / inv_s).expand(B * n, 1), weights=weights, cdf=c,
                gradient_error=eikonal(normals.reshape(B, n, 3)), sdf=sdf, normals=normals,
                rgb=rgb)


def rays_to_local(rays_o, rays_d, Ro, To, repeat=False):
    """convert_obj_to_local (utils/renderer.py:180-188, 424-432; batched renderer_batch.py:176-182).
    Un-batched: Ro [3,3], To [3], rays [B,3].  Batched: Ro [F,3,3], To [F,3], rays [F,P,3]."""
    if Ro.dim() == 2:
        # the un-batched fitting renderer materialises Ro per ray (utils/renderer.py:427), which
        # routes through bmm and rounds differently from the broadcast matmul of :183
        Rb = Ro[None].repeat(rays_o.shape[0], 1, 1) if repeat else Ro[None]
        o = torch.matmul(Rb, (rays_o - To[None])[..., None])[..., 0]
        d = torch.matmul(Rb, rays_d[..., None])[..., 0]
    else:
        o = torch.matmul(Ro[:, None], (rays_o - To[:, None])[..., None])[..., 0]
        d = torch.matmul(Ro[:, None], rays_d[..., None])[..., 0]
    return o, d


def coarse_z(near, far, n_samples, batch_shape, t_rand: Optional[Tensor], device="cpu"):
    """Coarse samples + one jitter per ray (utils/renderer.py:203-212; SURVEY A-1)."""
    sample_dist = (far - near) / n_samples
    z = torch.linspace(0.0, 1.0, n_samples, device=device)
    z = near + (far - near) * z
    z = z.expand(*batch_shape, n_samples)
    if t_rand is not None:
        z = z + t_rand * sample_dist
    return z.contiguous(), sample_dist


def render_obj(sdf_p, color_p, variance, rays_o, rays_d, near, far, Ro, To, t_rand=None,
               n_samples=64, n_importance=64, up_sample_steps=4, scale=1.0):
    """NeuSRenderer.render with model_type == 'obj' (utils/renderer.py:190-258).  ``t_rand`` is
    the [B,1] jitter ``rand - 0.5`` (None == perturb 0)."""
    rays_o, rays_d = rays_to_local(rays_o, rays_d, Ro, To)
    B = rays_o.shape[0]
    z_vals, sample_dist = coarse_z(near, far, n_samples, (B,), t_rand, rays_o.device)
    if n_importance > 0:
        z_vals = hierarchical_z(lambda q: sdf_obj_forward(sdf_p, q, scale=scale)[:, :1],
                                rays_o.detach(), rays_d.detach(), z_vals, n_importance,
                                up_sample_steps)
    core = render_core_obj(sdf_p, color_p, variance, rays_o, rays_d, z_vals, sample_dist, scale)
    return _finish_render(core, z_vals)


def render_hand(sdf_p, color_p, variance, rays_o, rays_d, near, far, bt_inv, T_pose_21,
                t_rand=None, n_samples=64, n_importance=64, up_sample_steps=4, r_multires=7):
    """NeuSRenderer.render with model_type == 'hand'."""
    B = rays_o.shape[0]
    z_vals, sample_dist = coarse_z(near, far, n_samples, (B,), t_rand, rays_o.device)
    if n_importance > 0:
        z_vals = hierarchical_z(
            lambda q: sdf_hand_forward(sdf_p, q, bt_inv.detach(), T_pose_21.detach(),
                                       r_multires=r_multires)[0][:, :1],
            rays_o.detach(), rays_d.detach(), z_vals, n_importance, up_sample_steps)
    core = render_core_hand(sdf_p, color_p, variance, rays_o, rays_d, z_vals, sample_dist, bt_inv,
                            T_pose_21, r_multires)
    return _finish_render(core, z_vals)


def _finish_render(core, z_vals):
    """Output dict of NeuSRenderer.render (utils/renderer.py:246-258; SURVEY A-6)."""
    B, n = z_vals.shape
    w = core["weights"]
    return {
        "color_fine": core["color"],
        "s_val": core["s_val"].reshape(B, n).mean(dim=-1, keepdim=True),
        "cdf_fine": core["cdf"],
        "weight_sum": w.sum(dim=-1, keepdim=True),
        "weight_max": torch.max(w, dim=-1, keepdim=True)[0],
        "gradient_error": core["gradient_error"],
        "z_vals": z_vals, "weights": w, "sdf": core["sdf"], "normals": core["normals"],
        "rgb": core["rgb"],
    }


def training_loss(out, true_rgb, true_mask, igr_weight=1.0, mask_weight=1.0):
    """Object/hand training loss without VGG (exp_runner.py:206-227)."""
    mask = (true_mask > 0.5).to(out["color_fine"].dtype)
    mask_sum = mask.sum() + 1e-5
    color_error = (out["color_fine"] - true_rgb) * mask
    color_loss = F.l1_loss(color_error, torch.zeros_like(color_error), reduction="sum") / mask_sum
    mask_loss = F.binary_cross_entropy(out["weight_sum"].clip(1e-3, 1.0 - 1e-3), mask)
    return color_loss + mask_loss * mask_weight + out["gradient_error"] * igr_weight


# --------------------------------------------------------------------------------------------
# two-field (hand + object) fitting renderers
# --------------------------------------------------------------------------------------------
def fit_field_alpha(kind, sdf_p, color_p, variance, rays_o, rays_d, z_vals, sample_dist,
                    bt_inv=None, T_pose_21=None, scale=1.0):
    """NeuSRenderer_fitting.get_alpha_sample_color (utils/renderer.py:360-422;
    utils/renderer_batch.py:115-174).  rays/z may carry a leading frame dim."""
    lead = z_vals.shape[:-1]
    n = z_vals.shape[-1]
    dists, pts, dirs = mid_points(rays_o, rays_d, z_vals, sample_dist)
    if not pts.requires_grad:
        pts.requires_grad_(True)
    if kind == "obj":
        out = sdf_obj_forward(sdf_p, pts, scale=scale)
        sdf, feat = out[:, :1], out[:, 1:]
        normals = sdf_gradient(lambda q: sdf_obj_forward(sdf_p, q, scale=scale)[:, :1], pts)
        rgb = color_obj_forward(color_p, pts, dirs, feat, normals)
    else:
        if z_vals.dim() == 3:
            Fn = z_vals.shape[0]
            hp = pts.reshape(Fn, -1, 3)
            fwd = lambda q: sdf_hand_forward(sdf_p, q.reshape(Fn, -1, 3), bt_inv, T_pose_21)
        else:
            hp = pts
            fwd = lambda q: sdf_hand_forward(sdf_p, q, bt_inv, T_pose_21)
        out, xyz_feature, _, _ = fwd(pts)
        sdf, feat = out[:, :1], out[:, 1:]
        normals = sdf_gradient(lambda q: fwd(q)[0][:, :1], pts)
        rgb = color_hand_forward(color_p, xyz_feature, feat, normals)
    inv_s = inv_s_from_variance(variance)
    alpha, _ = neus_alpha(sdf, normals, dirs, dists.reshape(-1, n), inv_s)
    return (alpha.reshape(*lead, n), rgb.reshape(*lead, n, 3), sdf.reshape(-1, 1),
            eikonal(normals), normals.reshape(-1, 3))


def fit_composite(alpha_h, rgb_h, alpha_o, rgb_o):
    """Joint transmittance of the two fields (utils/renderer.py:512-524;
    utils/renderer_batch.py:258-270).  Leading cumprod entry is ONE here (SURVEY D-1)."""
    final = (1.0 - alpha_h + 1e-7) * (1.0 - alpha_o + 1e-7)
    T = torch.cumprod(torch.cat([torch.ones_like(final[..., :1]), final], -1), -1)[..., :-1]
    w_h = alpha_h * T
    w_o = alpha_o * T
    color = (rgb_h * w_h[..., None]).sum(dim=-2) + (rgb_o * w_o[..., None]).sum(dim=-2)
    wsum = w_h.sum(dim=-1, keepdim=True) + w_o.sum(dim=-1, keepdim=True)
    return color, wsum, w_h, w_o


def fit_hierarchical_z(hand_sdf_fn, obj_sdf_fn, ro_h, rd_h, ro_o, rd_o, z_vals, n_importance,
                       up_sample_steps, batched_quirk=False):
    """The no-grad loop of NeuSRenderer_fitting.render (utils/renderer.py:460-498;
    utils/renderer_batch.py:207-243): each field up-samples on its own growing list; all new
    z's of both fields are appended to the shared list, which is sorted at the end.

    ``batched_quirk`` reproduces SURVEY D-7: in the frame-batched renderer the re-ordered SDF of
    frames >= 1 is gathered from frame 0's rows (utils/renderer_batch.py:108-111)."""
    lead = z_vals.shape[:-1]
    with torch.no_grad():
        zs = {"h": z_vals, "o": z_vals}
        sdf = {"h": hand_sdf_fn(ray_points(ro_h, rd_h, z_vals)).reshape(*lead, -1),
               "o": obj_sdf_fn(ray_points(ro_o, rd_o, z_vals)).reshape(*lead, -1)}
        rays = {"h": (ro_h, rd_h, hand_sdf_fn), "o": (ro_o, rd_o, obj_sdf_fn)}
        shared = z_vals
        for i in range(up_sample_steps):
            last = i + 1 == up_sample_steps
            new = {}
            for k in ("h", "o"):
                o, d, fn = rays[k]
                new[k] = up_sample(zs[k], sdf[k], n_importance // up_sample_steps, 64 * 2 ** i)
                merged, index = merge_sorted(zs[k], new[k])
                if not last:
                    new_sdf = fn(ray_points(o, d, new[k])).reshape(*lead, -1)
                    cat = torch.cat([sdf[k], new_sdf], dim=-1)
                    if batched_quirk and cat.dim() == 3:
                        cat = cat[:1].expand_as(cat)
                    sdf[k] = torch.gather(cat, -1, index)
                zs[k] = merged
            shared = torch.cat([shared, new["h"], new["o"]], dim=-1)
        shared, _ = torch.sort(shared, dim=-1)
    return shared


def fit_render(hand, obj, rays_o, rays_d, near, far, bt_inv, T_pose_21, Ro, To, t_rand=None,
               n_samples=64, n_importance=64, up_sample_steps=4, scale=1.0):
    """NeuSRenderer_fitting.render, un-batched (utils/renderer.py:434-535) and frame-batched
    (utils/renderer_batch.py:184-281).  ``hand`` / ``obj`` are (sdf_params, color_params, variance)."""
    batched = rays_o.dim() == 3
    ro_o, rd_o = rays_to_local(rays_o, rays_d, Ro, To, repeat=True)
    lead = rays_o.shape[:-1]
    z_vals, sample_dist = coarse_z(near, far, n_samples, lead, t_rand, rays_o.device)
    if batched:
        Fn = rays_o.shape[0]
        hand_sdf = lambda q: sdf_hand_forward(hand[0], q.reshape(Fn, -1, 3), bt_inv, T_pose_21)[0][:, :1]
    else:
        hand_sdf = lambda q: sdf_hand_forward(hand[0], q.reshape(-1, 3), bt_inv, T_pose_21)[0][:, :1]
    obj_sdf = lambda q: sdf_obj_forward(obj[0], q.reshape(-1, 3), scale=scale)[:, :1]
    if n_importance > 0:
        z_vals = fit_hierarchical_z(hand_sdf, obj_sdf, rays_o.detach(), rays_d.detach(),
                                    ro_o.detach(), rd_o.detach(), z_vals, n_importance,
                                    up_sample_steps, batched_quirk=batched)
    else:
        z_vals, _ = torch.sort(z_vals, dim=-1)
    a_h, c_h, sdf_h, ge_h, n_h = fit_field_alpha("hand", hand[0], hand[1], hand[2], rays_o, rays_d,
                                                 z_vals, sample_dist, bt_inv, T_pose_21)
    a_o, c_o, sdf_o, ge_o, n_o = fit_field_alpha("obj", obj[0], obj[1], obj[2], ro_o, rd_o, z_vals,
                                                 sample_dist, scale=scale)
    color, wsum, w_h, w_o = fit_composite(a_h, c_h, a_o, c_o)
    return {"color_fine": color, "weight_sum": wsum, "sdf_hand": sdf_h, "sdf_obj": sdf_o,
            "gradient_error_hand": ge_h, "gradient_error_obj": ge_o, "gradient_hand": n_h,
            "gradient_obj": n_o, "z_vals": z_vals}


# --------------------------------------------------------------------------------------------
# SDF lattice (mesh extraction input)
# --------------------------------------------------------------------------------------------
def sdf_grid(sdf_fn, bound_min: Tensor, bound_max: Tensor, resolution: int, chunk=64) -> Tensor:
    """The ``u`` lattice of extract_geometry (utils/renderer.py:260-278): linspace per axis, ij
    meshgrid, SDF at each node.  Returns [res,res,res] float32."""
    xs = torch.linspace(float(bound_min[0]), float(bound_max[0]), resolution)
    ys = torch.linspace(float(bound_min[1]), float(bound_max[1]), resolution)
    zs = torch.linspace(float(bound_min[2]), float(bound_max[2]), resolution)
    u = torch.zeros(resolution, resolution, resolution)
    with torch.no_grad():
        for xi in range(0, resolution, chunk):
            xx, yy, zz = torch.meshgrid(xs[xi:xi + chunk], ys, zs, indexing="ij")
            pts = torch.stack([xx, yy, zz], dim=-1).reshape(-1, 3)
            u[xi:xi + chunk] = sdf_fn(pts).reshape(xx.shape)
    return u


# --------------------------------------------------------------------------------------------
# caller-side losses and ray generation
# --------------------------------------------------------------------------------------------
def training_loss_terms(out, true_rgb, true_mask):
    """exp_runner.py:205-224: (color_fine_loss, mask_loss, psnr)."""
    mask = (true_mask > 0.5).to(out["color_fine"].dtype)
    mask_sum = mask.sum() + 1e-5
    color_error = (out["color_fine"] - true_rgb) * mask
    color_loss = F.l1_loss(color_error, torch.zeros_like(color_error), reduction="sum") / mask_sum
    psnr = 20.0 * torch.log10(1.0 / (((out["color_fine"] - true_rgb) ** 2 * mask).sum() / (mask_sum * 3.0)).sqrt())
    mask_loss = F.binary_cross_entropy(out["weight_sum"].clip(1e-3, 1.0 - 1e-3), mask)
    return color_loss, mask_loss, psnr


def fitting_render_loss(out, true_rgb, true_mask, scale=1.0):
    """fitting_single.py:253-256 (scale=1); fitting_video.py:287-291 is the same on [F*P] flattened rays with
    scale=0.5 (its divisor true_mask.shape[0] * true_mask.shape[1] is the ray count)."""
    color_error = (out["color_fine"] - true_rgb) * true_mask
    n = out["weight_sum"].numel()
    color_loss = F.l1_loss(color_error, torch.zeros_like(color_error), reduction="sum") / n
    mask_loss = F.binary_cross_entropy(out["weight_sum"].clip(1e-3, 1.0 - 1e-3), true_mask)
    return scale * (color_loss + 0.5 * mask_loss)


def interaction_loss(sdf_hand, sdf_obj, w_contact=30.0, w_penet=20.0):
    """fitting_single.py:268-283 / fitting_video.py:295-309: returns (total, contact_loss, penet_loss)."""
    sdf_hand, sdf_obj = sdf_hand[:, 0], sdf_obj[:, 0]
    sdf_abs_sum = torch.abs(sdf_hand) + torch.abs(sdf_obj)
    contact_id = sdf_abs_sum < 1e-2
    contact_num = contact_id.to(sdf_hand.dtype).sum() + 1e-9
    contact_loss = torch.sum(sdf_abs_sum[contact_id]) / contact_num
    inner = sdf_obj < 0
    hs, os_ = sdf_hand[inner], sdf_obj[inner]
    pen = hs < 0
    penet_num = pen.to(sdf_hand.dtype).sum() + 1e-9
    penet_loss = torch.sum(torch.abs(hs[pen]) + torch.abs(os_[pen])) / penet_num
    return w_contact * contact_loss + w_penet * penet_loss, contact_loss, penet_loss


def unproject_ndc(R, T, focal, pp, xy_depth):
    """pytorch3d (un-vendored dependency of the reference; version not pinned upstream) PerspectiveCameras
    .unproject_points(xy_depth, from_ndc=True, world_coordinates=True), restated from its published camera model
    (row vectors): X_view = X_world R + T; x_ndc = fx X/Z + px, y_ndc = fy Y/Z + py.  PARITY UNPINNED for this function
    (pytorch3d is not importable here); the bundle construction around it IS pinned (tests/golden/rays.npz)."""
    x, y, depth = xy_depth[..., 0], xy_depth[..., 1], xy_depth[..., 2]
    view = torch.stack([(x - pp[0]) * depth / focal[0], (y - pp[1]) * depth / focal[1], depth], dim=-1)
    return (view - T) @ torch.linalg.inv(R)


def rays_from_ndc(R, T, focal, pp, xy):
    """utils/utils.py:78-107: depth-1 and depth-2 planes -> unit directions and origins one unit behind plane 1."""
    ones = torch.ones_like(xy[..., :1])
    p1 = unproject_ndc(R, T, focal, pp, torch.cat([xy, ones], dim=-1))
    p2 = unproject_ndc(R, T, focal, pp, torch.cat([xy, 2.0 * ones], dim=-1))
    d = F.normalize(p2 - p1, dim=-1)
    return p1 - d, d


def ndc_grid_xy(H, W):
    """exp_runner.py:338-350: [H*W, 2] NDC coordinates of the full image (pixel = row * W + col)."""
    if W >= H:
        range_x, range_y = W / H, 1.0
    else:
        range_x, range_y = 1.0, H / W
    img_x = torch.linspace(range_x, -range_x, W).unsqueeze(0).repeat(H, 1).reshape(-1, 1)
    img_y = torch.linspace(range_y, -range_y, H).unsqueeze(1).repeat(1, W).reshape(-1, 1)
    return torch.cat((img_x, img_y), -1)


def stable_loss_from_sdf(hand_sdf, pts0):
    """utils/renderer_batch.py:328-369 given hand_sdf [F,P] (the hand SDF of every frame at the strided object
    vertices) and pts0 [P,3] = pts[0]: the reference's host logic line by line, scipy cKDTree included -- and
    including its quirk that np.setdiff1d receives the BOOLEAN in-mask."""
    import numpy as np
    from scipy import spatial
    p_num = pts0.shape[0]
    vert_id_all = range(p_num)
    hand_sdf_list, in_id_list = [], []
    for f in range(hand_sdf.shape[0]):
        cur = hand_sdf[f].reshape(-1)
        penet_id = cur < 0
        if penet_id.float().sum() > 0:
            in_id_list.append(penet_id)
            hand_sdf_list.append(cur)
    stable_loss = 0
    if len(in_id_list) > 1:
        hand_sdf_list = torch.stack(hand_sdf_list, 0)
        in_time = hand_sdf_list.shape[0]
        for cid in range(in_time):
            cur_in_id = in_id_list[cid].clone().cpu()
            cur_out_id = np.setdiff1d(vert_id_all, cur_in_id)
            in_points = pts0[cur_in_id].detach().cpu()
            out_points = pts0[cur_out_id].detach().cpu()
            n_in = in_points.shape[0]
            _, near = spatial.cKDTree(out_points.numpy()).query(in_points.numpy(), k=1)
            near = np.unique(near.reshape(-1))
            in_err = hand_sdf_list[:, cur_in_id].clip(0, 1e7).sum() / ((in_time - 1) * n_in)
            sel = hand_sdf_list[:, cur_out_id]
            out_err = torch.abs(sel[:, near].clip(-1e7, 0)).sum() / ((in_time - 1) * n_in)
            stable_loss = stable_loss + in_err + 0.05 * out_err
        stable_loss /= in_time
    return stable_loss


# ------------------------------------------------------------------------------------------------
# render_core_outside (background NeRF branch).  NOT in the HO-NeRF reference: utils/renderer.py:47,56 only store
# n_outside and every config sets 0.  The north star names it, so it is restated from the semantics of the NeuS renderer
# the reference derives from -- PARITY UNPINNED: no reference code, golden vector or test exists to pin this function to.
# ------------------------------------------------------------------------------------------------
def render_core_outside(rays_o, rays_d, z_vals, sample_dist, nerf, n_outside=1, background_rgb=None):
    """z_vals [B,n] -> dict(color [B,3], sampled_color [B,n,3], alpha [B,n], weights [B,n]).  `nerf(pts, dirs)` returns
    (density [B*n,1], raw rgb [B*n,3]); pts are the inverted-sphere coordinates (p / r, 1 / r), r = clip(|p|, 1, 1e10)."""
    B, n = z_vals.shape
    dists = z_vals[..., 1:] - z_vals[..., :-1]
    dists = torch.cat([dists, torch.full_like(dists[..., :1], float(sample_dist))], -1)
    mid_z_vals = z_vals + dists * 0.5
    pts = rays_o[:, None, :] + rays_d[:, None, :] * mid_z_vals[..., :, None]
    dis_to_center = torch.linalg.norm(pts, ord=2, dim=-1, keepdim=True).clip(1.0, 1e10)
    pts = torch.cat([pts / dis_to_center, 1.0 / dis_to_center], dim=-1)
    dirs = rays_d[:, None, :].expand(B, n, 3)
    pts = pts.reshape(-1, 4)[:, : 3 + int(n_outside > 0)]
    dirs = dirs.reshape(-1, 3)
    density, sampled_color = nerf(pts, dirs)
    sampled_color = torch.sigmoid(sampled_color)
    alpha = 1.0 - torch.exp(-F.softplus(density.reshape(B, n)) * dists)
    weights = alpha * torch.cumprod(torch.cat([torch.ones_like(alpha[:, :1]), 1.0 - alpha + 1e-7], -1), -1)[:, :-1]
    sampled_color = sampled_color.reshape(B, n, 3)
    color = (weights[:, :, None] * sampled_color).sum(dim=1)
    if background_rgb is not None:
        color = color + background_rgb * (1.0 - weights.sum(dim=-1, keepdim=True))
    return {"color": color, "sampled_color": sampled_color, "alpha": alpha, "weights": weights}
