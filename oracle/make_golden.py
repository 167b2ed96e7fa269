"""Generate golden input/output vectors from the UNMODIFIED reference (run in the build container).

    python oracle/make_golden.py            # writes tests/golden/*.npz

TEST INFRASTRUCTURE ONLY.  The reference is Python and cannot travel to the GPU box, so its
outputs on seeded synthetic inputs (oracle/synth.py) are committed as small fixtures.  The inputs
are NOT stored: tests rebuild them from the same seeds (same torch build in this image), which also
keeps the fixtures small.  Every array here comes from calling reference code
(utils/fields.py, utils/renderer.py, utils/renderer_batch.py) -- none from this repository's
restatement.
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import ref_loader  # noqa: E402
import synth  # noqa: E402
import cases  # noqa: E402

OUT = os.path.join(os.path.dirname(HERE), "tests", "golden")


class fixed_rand:
    """Make the reference's unseeded ``torch.rand`` jitter (utils/renderer.py:211) return a given
    tensor: rand() := t_rand + 0.5."""

    def __init__(self, t_rand):
        self.val = t_rand + 0.5

    def __enter__(self):
        self.orig = torch.rand
        torch.rand = lambda *a, **k: self.val.clone()

    def __exit__(self, *a):
        torch.rand = self.orig


def np_(t):
    return t.detach().cpu().numpy()


def build_obj(ref):
    sp, cp = synth.obj_states()
    emb = ref.fields.Embedding()
    sdf = ref.fields.SDFNetwork_OBJ(emb, 4, "real", **ref_loader.OBJ_SDF_CONF)
    col = ref.fields.RenderingNetwork_OBJ(emb, "real", **ref_loader.OBJ_COLOR_CONF)
    dev = ref.fields.SingleVarianceNetwork(ref_loader.VARIANCE_INIT)
    sdf.load_state_dict(sp)
    col.load_state_dict(cp)
    return sdf, col, dev


def build_hand(ref, use_batch=False):
    sp, cp = synth.hand_states()
    emb = ref.fields.Embedding()
    sdf = ref.fields.SDFNetwork(emb, 4, "real", use_batch=use_batch, **ref_loader.HAND_SDF_CONF)
    col = ref.fields.RenderingNetwork(emb, "real", **ref_loader.HAND_COLOR_CONF)
    dev = ref.fields.SingleVarianceNetwork(ref_loader.VARIANCE_INIT)
    sdf.load_state_dict(sp)
    col.load_state_dict(cp)
    return sdf, col, dev


def grads_of(loss, named):
    names, tensors = zip(*named)
    gs = torch.autograd.grad(loss, tensors, allow_unused=True)
    return {n: (np_(g) if g is not None else None) for n, g in zip(names, gs)}


def select_params(prefix, module):
    """bias + weight_g of every layer (full) and weight_v of every layer (full)."""
    return [(prefix + n, p) for n, p in module.named_parameters() if n != "se3_refine"]


def compress_grads(g):
    """Keep the fixtures small: weight_v gradients are stored as their first 4 rows plus the
    row-sum and column-sum vectors (any layout or scaling error moves at least one of those)."""
    out = {}
    for k, v in g.items():
        if v is None:
            continue
        if k.endswith("weight_v"):
            out[k + ":rows4"] = v[:4].astype(np.float32)
            out[k + ":rowsum"] = v.sum(1).astype(np.float32)
            out[k + ":colsum"] = v.sum(0).astype(np.float32)
        else:
            out[k] = v.astype(np.float32)
    return out


def main():
    os.makedirs(OUT, exist_ok=True)
    ref = ref_loader.load_reference()
    torch.manual_seed(0)
    zeros_bt, zeros_T = torch.zeros(21, 4, 4), torch.zeros(21, 3)

    # ---- A: encoding layout (utils/fields.py:13-20) -----------------------------------------
    x = cases.embed_case()["x"]
    emb = ref.fields.Embedding()
    np.savez(os.path.join(OUT, "embed.npz"), enc10=np_(emb(x, 10)), enc4=np_(emb(x, 4)),
             enc7=np_(emb(x, 7)))

    # ---- B: object fields (utils/fields.py:316-347, 387-405) --------------------------------
    sdf, col, dev = build_obj(ref)
    c = cases.obj_fields_case()
    pts, dirs = c["pts"], c["dirs"]
    out = sdf(pts)
    grad = sdf.gradient(pts.clone()).squeeze()
    rgb = col(pts, dirs, out[:, 1:], grad, 0)
    np.savez(os.path.join(OUT, "obj_fields.npz"), sdf_out=np_(out), gradient=np_(grad), rgb=np_(rgb))

    # ---- C: hierarchical sampling (utils/renderer.py:10-37, 60-105) -------------------------
    r = ref.renderer.NeuSRenderer(sdf, dev, col, "obj", **ref_loader.RENDERER_CONF)
    c = cases.sampling_case()
    R, z = c["R"], c["z0"]
    lo, ld = r.convert_obj_to_local(R["rays_o"], R["rays_d"], R["Ro"], R["To"])
    with torch.no_grad():
        s = sdf.sdf((lo[:, None, :] + ld[:, None, :] * z[..., :, None]).reshape(-1, 3)).reshape(24, 64)
        rec = {"local_o": np_(lo), "local_d": np_(ld), "z0": np_(z), "sdf0": np_(s)}
        for i in range(4):
            new_z = r.up_sample(lo, ld, z, s, 16, 64 * 2 ** i)
            rec["new_z%d" % i] = np_(new_z)
            z, s = r.cat_z_vals(lo, ld, z, new_z, s, zeros_bt, zeros_T, last=(i == 3))
            rec["z%d" % (i + 1)] = np_(z)
            if i < 3:
                rec["sdf%d" % (i + 1)] = np_(s)
        # sample_pdf in isolation, with its internals re-derived by the same torch calls
        w, bins = c["pdf_w"], c["pdf_bins"]
        rec["pdf_samples"] = np_(ref.renderer.sample_pdf(bins, w, 16, det=True))
    np.savez(os.path.join(OUT, "sampling.npz"), **rec)

    # ---- D: object render + training-loss gradients (utils/renderer.py:190-258) -------------
    c = cases.obj_render_case()
    R, true_rgb, true_mask = c["R"], c["true_rgb"], c["true_mask"]
    Ro = R["Ro"].clone().requires_grad_(True)
    To = R["To"].clone().requires_grad_(True)
    with fixed_rand(R["t_rand"]):
        out = r.render(R["rays_o"], R["rays_d"], R["near"], R["far"], zeros_bt, zeros_T, None, Ro, To, 0)
    mask_sum = true_mask.sum() + 1e-5
    color_error = (out["color_fine"] - true_rgb) * true_mask
    loss = torch.nn.functional.l1_loss(color_error, torch.zeros_like(color_error), reduction="sum") / mask_sum \
        + torch.nn.functional.binary_cross_entropy(out["weight_sum"].clip(1e-3, 1.0 - 1e-3), true_mask) \
        + out["gradient_error"]
    named = select_params("sdf.", sdf) + select_params("color.", col) + \
        [("variance", dev.variance), ("Ro", Ro), ("To", To)]
    gr = compress_grads(grads_of(loss, named))
    rec = {k: np_(v) for k, v in out.items()}
    rec["loss"] = np_(loss)
    rec.update({"grad:" + k: v for k, v in gr.items()})
    np.savez(os.path.join(OUT, "obj_render.npz"), **rec)

    # ---- E: hand fields (utils/fields.py:22-52, 132-177, 222-240) ---------------------------
    hsdf, hcol, hdev = build_hand(ref)
    c = cases.hand_fields_case()
    bt, T, J, pts = c["bt_inv"], c["T_pose_21"], c["J"], c["pts"]
    out, xyz_feature, rr, hh = hsdf(pts, bt, T)
    grad = hsdf.gradient(pts.clone(), bt, T).squeeze()
    rgb = hcol(None, xyz_feature, out[:, 1:], hh, grad, 0)
    np.savez(os.path.join(OUT, "hand_fields.npz"), sdf_out=np_(out), xyz_feature=np_(xyz_feature),
             r=np_(rr), h=np_(hh), gradient=np_(grad), rgb=np_(rgb))

    # ---- F: hand render + gradients to bt_inv (utils/renderer.py:190-258, hand branch) -------
    hr = ref.renderer.NeuSRenderer(hsdf, hdev, hcol, "hand", **ref_loader.RENDERER_CONF)
    c = cases.hand_render_case()
    HR, true_rgb = c["R"], c["true_rgb"]
    btg = bt.clone().requires_grad_(True)
    Tg = T.clone().requires_grad_(True)
    with fixed_rand(HR["t_rand"]):
        out = hr.render(HR["rays_o"], HR["rays_d"], HR["near"], HR["far"], btg, Tg, None, None, None, 0)
    loss = cases.hand_render_loss(out, true_rgb)
    named = select_params("sdf.", hsdf) + select_params("color.", hcol) + \
        [("variance", hdev.variance), ("bt_inv", btg), ("T_pose_21", Tg)]
    gr = compress_grads(grads_of(loss, named))
    rec = {k: np_(v) for k, v in out.items()}
    rec["loss"] = np_(loss)
    rec.update({"grad:" + k: v for k, v in gr.items()})
    np.savez(os.path.join(OUT, "hand_render.npz"), **rec)

    # ---- G: two-field fitting renderers (utils/renderer.py:434-535; renderer_batch.py:184-281)
    fr = ref.renderer.NeuSRenderer_fitting(hsdf, hdev, hcol, sdf, dev, col, **ref_loader.RENDERER_CONF)
    c = cases.fit_render_case()
    HR, true_rgb = c["R"], c["true_rgb"]
    Ro = c["Ro"].clone().requires_grad_(True)
    To = c["To"].clone().requires_grad_(True)
    btg = bt.clone().requires_grad_(True)
    with fixed_rand(HR["t_rand"]):
        out = fr.render(HR["rays_o"], HR["rays_d"], HR["near"], HR["far"], btg, T, None, Ro, To)
    loss = cases.fit_loss(out, true_rgb)
    gr = compress_grads(grads_of(loss, [("bt_inv", btg), ("Ro", Ro), ("To", To)]))
    rec = {k: np_(v) for k, v in out.items()}
    rec["loss"] = np_(loss)
    rec.update({"grad:" + k: v for k, v in gr.items()})
    np.savez(os.path.join(OUT, "fit_render.npz"), **rec)

    # batched (frame dim) -- use_batch=True hand net
    hsdf_b, hcol_b, hdev_b = build_hand(ref, use_batch=True)
    frb = ref.renderer_batch.NeuSRenderer_fitting(hsdf_b, hdev_b, hcol_b, sdf, dev, col,
                                                  **ref_loader.RENDERER_CONF)
    c = cases.fit_render_batch_case()
    btF, TF, ro, rd, tr, true_rgb = c["bt_inv"], c["T_pose_21"], c["rays_o"], c["rays_d"], c["t_rand"], c["true_rgb"]
    RoF = c["Ro"].clone().requires_grad_(True)
    ToF = c["To"].clone().requires_grad_(True)
    btFg = btF.clone().requires_grad_(True)
    with fixed_rand(tr):
        out = frb.render(ro, rd, 0.4, 1.5, btFg, TF, None, RoF, ToF)
    loss = cases.fit_loss(out, true_rgb)
    gr = compress_grads(grads_of(loss, [("bt_inv", btFg), ("Ro", RoF), ("To", ToF)]))
    rec = {k: np_(v) for k, v in out.items()}
    rec["loss"] = np_(loss)
    rec.update({"grad:" + k: v for k, v in gr.items()})
    np.savez(os.path.join(OUT, "fit_render_batch.npz"), **rec)

    # ---- H: SDF lattice (utils/renderer.py:260-278) -----------------------------------------
    c = cases.sdf_grid_case()
    res = c["res"]
    xs = torch.linspace(c["lo"], c["hi"], res)
    xx, yy, zz3 = torch.meshgrid(xs, xs, xs)
    gp = torch.cat([xx.reshape(-1, 1), yy.reshape(-1, 1), zz3.reshape(-1, 1)], dim=-1)
    with torch.no_grad():
        u = sdf.sdf(gp).reshape(res, res, res)
    np.savez(os.path.join(OUT, "sdf_grid.npz"), u=np_(u))
    for f in sorted(os.listdir(OUT)):
        print(f, os.path.getsize(os.path.join(OUT, f)))


if __name__ == "__main__":
    main()
