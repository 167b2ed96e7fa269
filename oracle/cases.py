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
