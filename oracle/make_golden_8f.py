"""Golden vectors for the SURVEY.md 8f rows (loss epilogues, ray generation), produced by running the reference's
OWN source lines (build container only; /root/reference does not travel).

    python oracle/make_golden_8f.py        # writes tests/golden/losses.npz, rays.npz, stable.npz

TEST INFRASTRUCTURE ONLY.
* Losses: the reference's loss code is inline in its drivers (exp_runner.py, fitting_single.py), not a function, so
  the exact source lines are read from the reference tree, dedented and exec'd on the seeded inputs of
  oracle/cases.py:loss_case; gradients come from torch autograd through those very lines.
* Rays: utils/utils.py is imported unmodified (pytorch3d and matplotlib stubbed: they are absent here) and its
  `_xy_to_ray_bundle` is called with a camera object whose `unproject_points` restates pytorch3d's published NDC camera
  model through the 4x4 projection-matrix inverse, the way pytorch3d composes it.  The bundle construction is therefore
  the reference's; the camera model is a restatement (pytorch3d is an un-vendored, un-pinned dependency).
"""
import importlib.util
import os
import sys
import textwrap
import types

import numpy as np
import torch
import torch.nn.functional as F

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import cases  # noqa: E402

REF = os.environ.get("HONERF_REFERENCE_ROOT", "/root/reference")
OUT = os.path.join(os.path.dirname(HERE), "tests", "golden")


def ref_lines(path, first, last, must_start_with):
    """Lines first..last (1-based, inclusive) of a reference file, dedented; the first line is checked so a moved
    line range fails loudly instead of silently running other code."""
    with open(os.path.join(REF, path)) as f:
        src = f.readlines()[first - 1:last]
    assert src[0].strip().startswith(must_start_with), (path, first, src[0])
    return textwrap.dedent("".join(src))


def np_(t):
    return t.detach().cpu().numpy()


def losses():
    c = cases.loss_case()
    out = {}
    # ---- training (exp_runner.py:206-227) --------------------------------------------------------
    color = c["color"].clone().requires_grad_(True)
    wsum = c["wsum"].clone().requires_grad_(True)
    ge = c["grad_err"].clone().requires_grad_(True)
    env = dict(torch=torch, F=F, true_mask=c["true_mask"].clone(), true_rgb=c["true_rgb"],
               render_out=dict(color_fine=color, weight_sum=wsum, gradient_error=ge, s_val=None, cdf_fine=None,
                               weight_max=None),
               self=types.SimpleNamespace(mask_weight=0.7, igr_weight=0.3))
    exec(ref_lines("exp_runner.py", 206, 207, "true_mask = (true_mask > 0.5)"), env)
    exec(ref_lines("exp_runner.py", 214, 227, "color_fine = render_out['color_fine']"), env)
    g = torch.autograd.grad(env["loss"], [color, wsum, ge])
    out.update({"train:loss": np_(env["loss"]), "train:color_loss": np_(env["color_fine_loss"]),
                "train:mask_loss": np_(env["mask_loss"]), "train:psnr": np_(env["psnr"]),
                "train:d_color": np_(g[0]), "train:d_wsum": np_(g[1]), "train:d_grad_err": np_(g[2])})
    # ---- fitting render loss (fitting_single.py:251-256) -----------------------------------------
    color = c["color"].clone().requires_grad_(True)
    wsum = c["wsum"].clone().requires_grad_(True)
    env = dict(torch=torch, F=F, true_mask=c["true_mask"], true_rgb=c["true_rgb"],
               render_out=dict(color_fine=color, weight_sum=wsum))
    exec(ref_lines("fitting_single.py", 251, 256, "color_fine = render_out['color_fine']"), env)
    g = torch.autograd.grad(env["render_loss"], [color, wsum])
    out.update({"fit:loss": np_(env["render_loss"]), "fit:color_loss": np_(env["color_fine_loss"]),
                "fit:mask_loss": np_(env["mask_loss"]), "fit:d_color": np_(g[0]), "fit:d_wsum": np_(g[1])})
    # ---- contact / penetration (fitting_single.py:268-282) ---------------------------------------
    sh = c["sdf_h"].clone().requires_grad_(True)
    so = c["sdf_o"].clone().requires_grad_(True)
    env = dict(torch=torch, render_out=dict(sdf_hand=sh, sdf_obj=so))
    exec(ref_lines("fitting_single.py", 268, 282, "sdf_hand = render_out['sdf_hand'][:,0]"), env)
    g = torch.autograd.grad(env["interaction_loss"], [sh, so])
    out.update({"int:loss": np_(env["interaction_loss"]), "int:contact": np_(env["contact_loss"]),
                "int:penet": np_(env["penet_loss"]), "int:contact_num": np_(env["contact_num"]),
                "int:penet_num": np_(env["penet_num"]), "int:d_h": np_(g[0]), "int:d_o": np_(g[1])})
    np.savez_compressed(os.path.join(OUT, "losses.npz"), **out)
    print("losses.npz:", {k: v.shape for k, v in out.items()})


class RestatedPerspectiveCameras:
    """pytorch3d PerspectiveCameras (NDC) restated: K = [[fx,0,px,0],[0,fy,py,0],[0,0,0,1],[0,0,1,0]] applied to row
    vectors (so the matrix used is K^T), world->view = [R 0; T 1]; unproject = inverse of the composed 4x4 applied to
    (x, y, 1/depth, 1) followed by the homogeneous divide."""

    def __init__(self, R, T, focal_length, principal_point):
        self.R, self.T, self.f, self.p = R, T, focal_length, principal_point

    def unproject_points(self, xy_depth, from_ndc=True):
        assert from_ndc
        n = self.R.shape[0]
        K = torch.zeros(n, 4, 4)
        K[:, 0, 0], K[:, 1, 1] = self.f[:, 0], self.f[:, 1]
        K[:, 0, 2], K[:, 1, 2] = self.p[:, 0], self.p[:, 1]
        K[:, 2, 3] = 1.0
        K[:, 3, 2] = 1.0
        w2v = torch.zeros(n, 4, 4)
        w2v[:, :3, :3] = self.R
        w2v[:, 3, :3] = self.T
        w2v[:, 3, 3] = 1.0
        full = w2v @ K.transpose(1, 2)
        inv = torch.inverse(full)
        pts = torch.cat([xy_depth[..., :2], 1.0 / xy_depth[..., 2:3], torch.ones_like(xy_depth[..., :1])], dim=-1)
        hom = pts @ inv
        return hom[..., :3] / hom[..., 3:]


def load_ref_utils():
    def stub(name, **attrs):
        m = sys.modules.get(name) or types.ModuleType(name)
        for k, v in attrs.items():
            setattr(m, k, v)
        sys.modules[name] = m
        return m

    class RayBundle:
        def __init__(self, origins, directions, lengths, xys):
            self.origins, self.directions, self.lengths, self.xys = origins, directions, lengths, xys
    stub("matplotlib"); stub("matplotlib.pyplot")
    stub("pytorch3d"); stub("pytorch3d.renderer"); stub("pytorch3d.renderer.cameras", CamerasBase=object)
    stub("pytorch3d.renderer.implicit"); stub("pytorch3d.renderer.implicit.utils", RayBundle=RayBundle)
    spec = importlib.util.spec_from_file_location("honerf_ref_utils_utils", os.path.join(REF, "utils", "utils.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def rays():
    c = cases.rays_case()
    U = load_ref_utils()
    cam = RestatedPerspectiveCameras(c["R"], c["T"], c["focal"], c["pp"])
    b = U._xy_to_ray_bundle(cam, c["xy"], 0.4, 1.5, 64)
    out = {"o": np_(b.origins), "d": np_(b.directions), "lengths0": np_(b.lengths[0, 0])}
    # full image through camera 0, the exp_runner.py:338-352 lines themselves
    env = dict(torch=torch, self=types.SimpleNamespace(W=c["W"], H=c["H"], device="cpu", near=0.4, far=1.5),
               _xy_to_ray_bundle=U._xy_to_ray_bundle,
               test_camera=RestatedPerspectiveCameras(c["R"][:1], c["T"][:1], c["focal"][:1], c["pp"][:1]))
    exec(ref_lines("exp_runner.py", 338, 353, "if self.W >= self.H:"), env)
    out.update({"grid_xy": np_(env["rays_xy"][0]), "grid_o": np_(env["rays_o"]), "grid_d": np_(env["rays_d"])})
    np.savez_compressed(os.path.join(OUT, "rays.npz"), **out)
    print("rays.npz:", {k: v.shape for k, v in out.items()})


def stable():
    """get_stable_loss_cross (utils/renderer_batch.py:318-371) of the UNMODIFIED reference class, called on a stub
    `self` that carries the reference's own hand SDF network (CPU, use_batch=True) with the seeded synthetic weights.
    lin8.bias[0] is shifted by the median SDF so the zero level set passes through the vertex cloud (stored as `shift`);
    vertices closer than 3e-4 to the surface in any frame are dropped (stored as `keep`) so a 1e-5 SDF difference on
    the GPU cannot move a vertex between the in and out sets."""
    import ref_loader
    import synth
    ref = ref_loader.load_reference()
    c = cases.stable_case()
    sp, _ = synth.hand_states()
    net = ref.fields.SDFNetwork(ref.fields.Embedding(), 4, "real", use_batch=True, **ref_loader.HAND_SDF_CONF)
    net.load_state_dict(sp)
    Fn = c["Ro"].shape[0]

    def sdf_at(sel):
        pw = (c["Ro"].unsqueeze(1) @ sel[None].repeat(Fn, 1, 1).unsqueeze(-1))[..., 0] + c["To"].unsqueeze(1)
        with torch.no_grad():
            return net.sdf(pw, c["bt_inv"], c["T_pose_21"]).reshape(Fn, -1)
    shift = sdf_at(c["sel"]).median()
    with torch.no_grad():
        net.lin8.bias[0] -= shift
    s = sdf_at(c["sel"])
    keep = (s.abs() > 3e-4).all(dim=0)
    sel = c["sel"][keep]
    pts = cases.stable_pts(sel, Fn)
    stub = types.SimpleNamespace(sdf_network_hand=net)
    cls = ref.renderer_batch.NeuSRenderer_fitting
    bt = c["bt_inv"].clone().requires_grad_(True)
    loss = cls.get_stable_loss_cross(stub, pts, bt, c["T_pose_21"], c["Ro"], c["To"])
    d_bt, = torch.autograd.grad(loss, [bt])
    hand_sdf = sdf_at(sel)
    out = {"shift": np_(shift), "keep": np_(keep), "hand_sdf": np_(hand_sdf), "loss": np_(loss), "d_bt_inv": np_(d_bt),
           "n_in": np_((hand_sdf < 0).sum(1))}
    # second fixture: only one frame penetrates -> the reference returns the int 0
    one = hand_sdf.clone()
    one[1:] = one[1:].abs() + 1e-3
    out["loss_one_frame"] = np.float32(cls.get_stable_loss_cross(
        types.SimpleNamespace(sdf_network_hand=types.SimpleNamespace(sdf=lambda p, b, t: one.reshape(-1, 1))),
        pts, bt, c["T_pose_21"], c["Ro"], c["To"]))
    np.savez_compressed(os.path.join(OUT, "stable.npz"), **out)
    print("stable.npz:", {k: v.shape for k, v in out.items()}, "loss", float(loss), "n_in", out["n_in"], "kept", int(keep.sum()))


if __name__ == "__main__":
    which = sys.argv[1:] or ["losses", "rays", "stable"]
    for w in which:
        {"losses": losses, "rays": rays, "stable": stable}[w]()
