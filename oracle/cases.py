"""Seeded inputs of the golden cases, shared by oracle/make_golden.py (which feeds them to the
reference) and by tests/ (which feed them to the oracle and to the CUDA path).
TEST INFRASTRUCTURE ONLY."""
import torch

import synth


def embed_case():
    g = torch.Generator().manual_seed(100)
    return dict(x=torch.randn(7, 3, generator=g))


def obj_fields_case():
    g = torch.Generator().manual_seed(101)
    pts = 0.45 * torch.randn(96, 3, generator=g)
    dirs = torch.nn.functional.normalize(torch.randn(96, 3, generator=g), dim=-1)
    return dict(pts=pts, dirs=dirs)


def sampling_case():
    R = synth.object_rays(24)
    z = 0.4 + 1.1 * torch.linspace(0.0, 1.0, 64)[None, :] + R["t_rand"] * (1.1 / 64)
    w = torch.rand(24, 63, generator=torch.Generator().manual_seed(102))
    bins = torch.sort(torch.rand(24, 64, generator=torch.Generator().manual_seed(103)), -1)[0]
    return dict(R=R, z0=z, pdf_w=w, pdf_bins=bins)


def obj_render_case():
    B = 24
    R = synth.object_rays(B, seed=5)
    g = torch.Generator().manual_seed(104)
    true_rgb = torch.rand(B, 3, generator=g)
    true_mask = (torch.rand(B, 1, generator=g) > 0.5).float()
    return dict(R=R, true_rgb=true_rgb, true_mask=true_mask)


def hand_fields_case():
    bt, T, J = synth.hand_pose()
    HR = synth.hand_rays(12, J)
    zz = torch.linspace(0.75, 1.05, 8)
    pts = (HR["rays_o"][:, None] + HR["rays_d"][:, None] * zz[None, :, None]).reshape(-1, 3)
    return dict(bt_inv=bt, T_pose_21=T, J=J, pts=pts)


def hand_render_case():
    bt, T, J = synth.hand_pose()
    B = 12
    HR = synth.hand_rays(B, J, seed=6)
    g = torch.Generator().manual_seed(105)
    return dict(bt_inv=bt, T_pose_21=T, R=HR, true_rgb=torch.rand(B, 3, generator=g))


def hand_render_loss(out, true_rgb):
    return (out["color_fine"] - true_rgb).abs().mean() + 0.1 * out["weight_sum"].mean() \
        + 1e-4 * out["gradient_error"]


def fit_render_case():
    bt, T, J = synth.hand_pose()
    B = 10
    HR = synth.hand_rays(B, J, seed=7)
    gq = torch.Generator().manual_seed(106)
    Ro = synth.random_rotation(gq)
    To = J.mean(0) + 0.02 * torch.randn(3, generator=gq)
    true_rgb = torch.rand(B, 3, generator=gq)
    return dict(bt_inv=bt, T_pose_21=T, R=HR, Ro=Ro, To=To, true_rgb=true_rgb)


def fit_render_batch_case():
    Fn, P = 2, 5
    gq = torch.Generator().manual_seed(107)
    btF, TF, JF = synth.hand_pose(n_frames=Fn)
    ro, rd, tr = [], [], []
    for f in range(Fn):
        h = synth.hand_rays(P, JF[f], seed=8 + f)
        ro.append(h["rays_o"]); rd.append(h["rays_d"]); tr.append(h["t_rand"])
    RoF = torch.stack([synth.random_rotation(gq) for _ in range(Fn)])
    ToF = JF.mean(1) + 0.02 * torch.randn(Fn, 3, generator=gq)
    true_rgb = torch.rand(Fn, P, 3, generator=gq)
    return dict(bt_inv=btF, T_pose_21=TF, rays_o=torch.stack(ro), rays_d=torch.stack(rd),
                t_rand=torch.stack(tr), Ro=RoF, To=ToF, true_rgb=true_rgb, near=0.4, far=1.5)


def fit_loss(out, true_rgb):
    return (out["color_fine"] - true_rgb).abs().mean() + 0.5 * out["weight_sum"].mean() \
        + out["sdf_hand"].clip(-1, 0).abs().mean() + out["sdf_obj"].clip(-1, 0).abs().mean()


def sdf_grid_case():
    return dict(res=12, lo=-0.6, hi=0.6)


def loss_case(n=200, seed=301):
    """Render outputs as the loss epilogues see them: colours in (0,1), weight sums that hit both clip bounds,
    SDF pairs with contact (|h|+|o| < 1e-2) and penetration (h<0, o<0) populations."""
    g = torch.Generator().manual_seed(seed)
    color = torch.rand(n, 3, generator=g)
    wsum = (torch.rand(n, 1, generator=g) * 1.2 - 0.1).clip(0.0, 1.05)
    wsum[:5] = torch.tensor([[0.0], [1e-3], [1.0 - 1e-3], [1.0], [0.5]])
    true_rgb = torch.rand(n, 3, generator=g)
    true_rgb[7] = color[7]                       # exact zero error: sign(0) = 0
    true_mask = (torch.rand(n, 1, generator=g) > 0.4).float()
    grad_err = torch.tensor(0.137)
    m = 6 * n
    sdf_h = 0.02 * torch.randn(m, 1, generator=g)
    sdf_o = 0.02 * torch.randn(m, 1, generator=g)
    sdf_h[:3] = torch.tensor([[0.0], [-0.001], [0.004]])
    sdf_o[:3] = torch.tensor([[-0.002], [0.0], [0.0]])
    return dict(color=color, wsum=wsum, true_rgb=true_rgb, true_mask=true_mask, grad_err=grad_err, sdf_h=sdf_h,
                sdf_o=sdf_o)


def rays_case(seed=302):
    """Two cameras on a ring looking roughly at the origin (pytorch3d row-vector convention), NDC intrinsics like
    the reference's labels (fx_ndc ~ 2-3, small principal-point offsets), random NDC points and a 7 x 5 image."""
    import synth
    g = torch.Generator().manual_seed(seed)
    R = torch.stack([synth.random_rotation(g) for _ in range(2)])
    T = torch.tensor([[0.02, -0.03, 0.95], [-0.05, 0.01, 1.10]]) + 0.01 * torch.randn(2, 3, generator=g)
    focal = torch.tensor([[2.3, 2.4], [2.9, 2.8]])
    pp = torch.tensor([[0.03, -0.02], [-0.04, 0.05]])
    xy = torch.rand(2, 33, 2, generator=g) * 2.4 - 1.2
    return dict(R=R, T=T, focal=focal, pp=pp, xy=xy, H=5, W=7)


def stable_case(n_frames=4, n_pts=160, seed=303):
    """get_stable_loss_cross inputs: object vertices (object frame) scattered around the hand joints, per-frame hand
    poses and object poses.  `pts` is [F, 10 * n_pts, 3] so that the reference's `[:, ::10]` stride picks `sel`."""
    import synth
    g = torch.Generator().manual_seed(seed)
    bt, T, J = synth.hand_pose(seed=5, n_frames=n_frames)
    sel = J[0][torch.randint(0, 21, (n_pts,), generator=g)] + 0.012 * torch.randn(n_pts, 3, generator=g)
    Ro = torch.stack([torch.eye(3) + 0.01 * torch.randn(3, 3, generator=g) for _ in range(n_frames)])
    To = 0.004 * torch.randn(n_frames, 3, generator=g)
    return dict(bt_inv=bt, T_pose_21=T, sel=sel, Ro=Ro, To=To)


def stable_pts(sel, n_frames):
    return sel.repeat_interleave(10, dim=0)[None].repeat(n_frames, 1, 1).contiguous()
