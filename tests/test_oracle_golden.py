"""The CPU oracle (oracle/honerf_oracle.py) against golden vectors produced by the real reference
(oracle/make_golden.py).  CPU only."""
import torch

import cases
import honerf_oracle as O
import synth
from golden_util import check_grads_against_golden, load_golden, max_abs, rel_err

VAR = torch.tensor(0.3)


def test_embed_layout():
    g = load_golden("embed")
    x = cases.embed_case()["x"]
    for L in (10, 4, 7):
        assert torch.equal(O.embed(x, L), g["enc%d" % L])
    # explicit layout statement: channel-major, sin block then cos block per channel
    e = O.embed(x, 4)
    assert torch.equal(e[:, 1 * 8 + 0 * 4 + 2], torch.sin(4.0 * x[:, 1]))
    assert torch.equal(e[:, 2 * 8 + 1 * 4 + 3], torch.cos(8.0 * x[:, 2]))


def test_obj_fields():
    g = load_golden("obj_fields")
    c = cases.obj_fields_case()
    sp, cp = synth.obj_states()
    out = O.sdf_obj_forward(sp, c["pts"])
    assert max_abs(out, g["sdf_out"]) < 2e-6
    n = O.sdf_gradient(lambda q: O.sdf_obj_forward(sp, q)[:, :1], c["pts"].clone())
    assert rel_err(n, g["gradient"]) < 1e-5
    rgb = O.color_obj_forward(cp, c["pts"], c["dirs"], out[:, 1:], n)
    assert max_abs(rgb, g["rgb"]) < 2e-6


def test_sampling_chain():
    g = load_golden("sampling")
    c = cases.sampling_case()
    sp, _ = synth.obj_states()
    R = c["R"]
    lo, ld = O.rays_to_local(R["rays_o"], R["rays_d"], R["Ro"], R["To"])
    assert max_abs(lo, g["local_o"]) < 1e-6 and max_abs(ld, g["local_d"]) < 1e-6
    lo, ld = g["local_o"], g["local_d"]
    z, s = g["z0"], g["sdf0"]
    assert torch.equal(z, c["z0"])
    for i in range(4):
        new_z = O.up_sample(z, s, 16, 64 * 2 ** i)
        assert torch.equal(new_z, g["new_z%d" % i]), "up_sample step %d" % i
        z, index = O.merge_sorted(z, new_z)
        assert torch.equal(z, g["z%d" % (i + 1)])
        if i < 3:
            s = g["sdf%d" % (i + 1)]
    samples = O.inverse_cdf(c["pdf_bins"], O.sample_pdf_cdf(c["pdf_w"]), 16)[0]
    assert torch.equal(samples, g["pdf_samples"])


def _params_named(prefix, sd):
    return {prefix + k: v for k, v in sd.items() if k != "se3_refine"}


def test_obj_render_and_grads():
    g = load_golden("obj_render")
    c = cases.obj_render_case()
    R = c["R"]
    sp, cp = synth.obj_states()
    sp = {k: v.clone().requires_grad_(True) for k, v in sp.items()}
    cp = {k: v.clone().requires_grad_(True) for k, v in cp.items()}
    var = VAR.clone().requires_grad_(True)
    Ro = R["Ro"].clone().requires_grad_(True)
    To = R["To"].clone().requires_grad_(True)
    out = O.render_obj(sp, cp, var, R["rays_o"], R["rays_d"], R["near"], R["far"], Ro, To, R["t_rand"])
    for k in ("color_fine", "s_val", "cdf_fine", "weight_sum", "weight_max"):
        assert max_abs(out[k], g[k]) < 1e-5, k
    assert rel_err(out["gradient_error"], g["gradient_error"]) < 1e-5
    loss = O.training_loss(out, c["true_rgb"], c["true_mask"])
    assert rel_err(loss, g["loss"]) < 1e-5
    named = {**_params_named("sdf.", sp), **_params_named("color.", cp), "variance": var, "Ro": Ro, "To": To}
    grads = torch.autograd.grad(loss, list(named.values()), allow_unused=True)
    check_grads_against_golden(g, {k: v for k, v in zip(named, grads) if v is not None}, 2e-3)


def test_hand_fields():
    g = load_golden("hand_fields")
    c = cases.hand_fields_case()
    sp, cp = synth.hand_states()
    out, feat, r, h = O.sdf_hand_forward(sp, c["pts"], c["bt_inv"], c["T_pose_21"])
    assert max_abs(feat, g["xyz_feature"]) < 1e-5
    assert max_abs(r, g["r"]) < 1e-5 and max_abs(h, g["h"]) < 1e-5
    assert max_abs(out, g["sdf_out"]) < 1e-4
    n = O.sdf_gradient(lambda q: O.sdf_hand_forward(sp, q, c["bt_inv"], c["T_pose_21"])[0][:, :1],
                       c["pts"].clone())
    assert rel_err(n, g["gradient"]) < 1e-4
    rgb = O.color_hand_forward(cp, feat, out[:, 1:], n)
    assert max_abs(rgb, g["rgb"]) < 1e-4


def test_hand_render_and_grads():
    g = load_golden("hand_render")
    c = cases.hand_render_case()
    R = c["R"]
    sp, cp = synth.hand_states()
    sp = {k: v.clone().requires_grad_(True) for k, v in sp.items()}
    cp = {k: v.clone().requires_grad_(True) for k, v in cp.items()}
    var = VAR.clone().requires_grad_(True)
    bt = c["bt_inv"].clone().requires_grad_(True)
    T = c["T_pose_21"].clone().requires_grad_(True)
    out = O.render_hand(sp, cp, var, R["rays_o"], R["rays_d"], R["near"], R["far"], bt, T, R["t_rand"])
    for k in ("color_fine", "cdf_fine", "weight_sum", "weight_max"):
        assert max_abs(out[k], g[k]) < 2e-4, k
    loss = cases.hand_render_loss(out, c["true_rgb"])
    assert rel_err(loss, g["loss"]) < 1e-4
    named = {**_params_named("sdf.", sp), **_params_named("color.", cp), "variance": var,
             "bt_inv": bt, "T_pose_21": T}
    grads = torch.autograd.grad(loss, list(named.values()), allow_unused=True)
    check_grads_against_golden(g, {k: v for k, v in zip(named, grads) if v is not None}, 2e-2)


def test_fit_render():
    g = load_golden("fit_render")
    c = cases.fit_render_case()
    R = c["R"]
    hs, hc = synth.hand_states()
    os_, oc = synth.obj_states()
    bt = c["bt_inv"].clone().requires_grad_(True)
    Ro = c["Ro"].clone().requires_grad_(True)
    To = c["To"].clone().requires_grad_(True)
    out = O.fit_render((hs, hc, VAR), (os_, oc, VAR), R["rays_o"], R["rays_d"], R["near"], R["far"],
                       bt, c["T_pose_21"], Ro, To, R["t_rand"])
    # 1-ulp differences in the ray transform can move an importance sample across a cdf knot, so
    # end-to-end two-field renders are compared at the north-star tolerance (1e-3), not at 1e-5
    for k in ("color_fine", "weight_sum", "sdf_hand", "sdf_obj"):
        assert max_abs(out[k], g[k]) < 1e-3, k
    assert rel_err(out["gradient_obj"], g["gradient_obj"]) < 1e-2
    assert rel_err(out["gradient_hand"], g["gradient_hand"]) < 1e-2
    loss = cases.fit_loss(out, c["true_rgb"])
    grads = torch.autograd.grad(loss, [bt, Ro, To])
    check_grads_against_golden(g, dict(zip(["bt_inv", "Ro", "To"], grads)), 2e-2)


def test_fit_render_batch_with_frame0_gather_quirk():
    g = load_golden("fit_render_batch")
    c = cases.fit_render_batch_case()
    hs, hc = synth.hand_states()
    os_, oc = synth.obj_states()
    bt = c["bt_inv"].clone().requires_grad_(True)
    Ro = c["Ro"].clone().requires_grad_(True)
    To = c["To"].clone().requires_grad_(True)
    out = O.fit_render((hs, hc, VAR), (os_, oc, VAR), c["rays_o"], c["rays_d"], c["near"], c["far"],
                       bt, c["T_pose_21"], Ro, To, c["t_rand"])
    for k in ("color_fine", "weight_sum", "sdf_hand", "sdf_obj"):
        assert max_abs(out[k], g[k]) < 1e-3, k
    loss = cases.fit_loss(out, c["true_rgb"])
    grads = torch.autograd.grad(loss, [bt, Ro, To])
    check_grads_against_golden(g, dict(zip(["bt_inv", "Ro", "To"], grads)), 2e-2)


def test_sdf_grid():
    g = load_golden("sdf_grid")
    c = cases.sdf_grid_case()
    sp, _ = synth.obj_states()
    lo = torch.full((3,), c["lo"])
    hi = torch.full((3,), c["hi"])
    u = O.sdf_grid(lambda q: O.sdf_obj_forward(sp, q)[:, :1], lo, hi, c["res"])
    assert max_abs(u, g["u"]) < 2e-6
