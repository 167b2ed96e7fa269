"""The CPU oracle against the UNMODIFIED reference, live, on inputs the golden fixtures do NOT hold (other seeds for the
network states, the rays, the pose and the jitter).  CPU only; skipped where the reference tree is absent (the GPU box):
there the committed fixtures of tests/golden (same generator, oracle/make_golden.py) carry the pin."""
import pytest
import torch

import honerf_oracle as O
import ref_loader
import synth
from golden_util import max_abs, rel_err

pytestmark = pytest.mark.skipif(not ref_loader.reference_available(), reason="reference tree not present")


class _fixed_rand:
    """the reference's unseeded jitter (utils/renderer.py:211): rand() := t_rand + 0.5, as oracle/make_golden.py does"""

    def __init__(self, t_rand):
        self.val = t_rand + 0.5

    def __enter__(self):
        self.orig = torch.rand
        torch.rand = lambda *a, **k: self.val.clone()

    def __exit__(self, *a):
        torch.rand = self.orig


def _obj_nets(ref, seed):
    sp, cp = synth.obj_states(seed=seed)
    emb = ref.fields.Embedding()
    sdf = ref.fields.SDFNetwork_OBJ(emb, 4, "real", **ref_loader.OBJ_SDF_CONF)
    col = ref.fields.RenderingNetwork_OBJ(emb, "real", **ref_loader.OBJ_COLOR_CONF)
    dev = ref.fields.SingleVarianceNetwork(ref_loader.VARIANCE_INIT)
    sdf.load_state_dict(sp)
    col.load_state_dict(cp)
    return sp, cp, sdf, col, dev


@pytest.mark.parametrize("seed", [31, 32])
def test_object_fields_fresh_seed(seed):
    """SDFNetwork_OBJ.forward / .gradient and RenderingNetwork_OBJ.forward (utils/fields.py:316-347, 387-405)"""
    ref = ref_loader.load_reference()
    sp, cp, sdf, col, _ = _obj_nets(ref, seed)
    g = torch.Generator().manual_seed(1000 + seed)
    pts = 0.5 * torch.randn(64, 3, generator=g)
    dirs = torch.nn.functional.normalize(torch.randn(64, 3, generator=g), dim=-1)
    want = sdf(pts)
    want_n = sdf.gradient(pts.clone()).squeeze()
    want_rgb = col(pts, dirs, want[:, 1:], want_n, 0)
    out = O.sdf_obj_forward(sp, pts)
    n = O.sdf_gradient(lambda q: O.sdf_obj_forward(sp, q)[:, :1], pts.clone())
    assert max_abs(out, want) < 2e-6
    assert rel_err(n, want_n) < 1e-5
    assert max_abs(O.color_obj_forward(cp, pts, dirs, out[:, 1:], n), want_rgb) < 2e-6


def test_object_render_and_training_gradients_fresh_seed():
    """NeuSRenderer.render for the object field, 8 rays x (64 + 64) samples, and the gradients of the training loss of
    exp_runner.py:206-227 through the second-order path (utils/renderer.py:107-258)"""
    ref = ref_loader.load_reference()
    sp, cp, sdf, col, dev = _obj_nets(ref, 33)
    B = 8
    R = synth.object_rays(B, seed=77)
    g = torch.Generator().manual_seed(78)
    true_rgb = torch.rand(B, 3, generator=g)
    true_mask = (torch.rand(B, 1, generator=g) > 0.4).float()
    r = ref.renderer.NeuSRenderer(sdf, dev, col, "obj", **ref_loader.RENDERER_CONF)
    Ro = R["Ro"].clone().requires_grad_(True)
    To = R["To"].clone().requires_grad_(True)
    with _fixed_rand(R["t_rand"]):
        want = r.render(R["rays_o"], R["rays_d"], R["near"], R["far"], torch.zeros(21, 4, 4), torch.zeros(21, 3), None, Ro, To, 0)
    mask_sum = true_mask.sum() + 1e-5
    ce = (want["color_fine"] - true_rgb) * true_mask
    want_loss = torch.nn.functional.l1_loss(ce, torch.zeros_like(ce), reduction="sum") / mask_sum \
        + torch.nn.functional.binary_cross_entropy(want["weight_sum"].clip(1e-3, 1.0 - 1e-3), true_mask) + want["gradient_error"]
    ref_named = [("sdf." + n, p) for n, p in sdf.named_parameters()] + [("color." + n, p) for n, p in col.named_parameters()] \
        + [("variance", dev.variance), ("Ro", Ro), ("To", To)]
    want_g = dict(zip([n for n, _ in ref_named], torch.autograd.grad(want_loss, [p for _, p in ref_named], allow_unused=True)))

    sp = {k: v.clone().requires_grad_(True) for k, v in sp.items()}
    cp = {k: v.clone().requires_grad_(True) for k, v in cp.items()}
    var = torch.tensor(ref_loader.VARIANCE_INIT).requires_grad_(True)
    Ro2 = R["Ro"].clone().requires_grad_(True)
    To2 = R["To"].clone().requires_grad_(True)
    out = O.render_obj(sp, cp, var, R["rays_o"], R["rays_d"], R["near"], R["far"], Ro2, To2, R["t_rand"])
    for k in ("color_fine", "s_val", "cdf_fine", "weight_sum", "weight_max"):
        assert max_abs(out[k], want[k]) < 1e-5, k
    assert rel_err(out["gradient_error"], want["gradient_error"]) < 1e-5
    loss = O.training_loss(out, true_rgb, true_mask)
    assert rel_err(loss, want_loss) < 1e-5
    named = {**{"sdf." + k: v for k, v in sp.items()}, **{"color." + k: v for k, v in cp.items()}, "variance": var, "Ro": Ro2,
             "To": To2}
    got = dict(zip(named, torch.autograd.grad(loss, list(named.values()), allow_unused=True)))
    checked = 0
    for k, w in want_g.items():
        if w is None or k.endswith("se3_refine"):
            continue
        assert got.get(k) is not None, k
        assert rel_err(got[k].reshape(w.shape), w) < 2e-3, k      # same bound as the golden test (fp32 autograd on both sides)
        checked += 1
    assert checked >= 40


def test_hand_fields_fresh_pose():
    """anerf_emb_point + SDFNetwork + RenderingNetwork of the hand (utils/fields.py:22-52, 132-177, 222-240) on another
    pose, other network states and other points than the fixture's"""
    ref = ref_loader.load_reference()
    sp, cp = synth.hand_states(seed=41)
    emb = ref.fields.Embedding()
    hsdf = ref.fields.SDFNetwork(emb, 4, "real", use_batch=False, **ref_loader.HAND_SDF_CONF)
    hcol = ref.fields.RenderingNetwork(emb, "real", **ref_loader.HAND_COLOR_CONF)
    hsdf.load_state_dict(sp)
    hcol.load_state_dict(cp)
    bt, T, J = synth.hand_pose(seed=9)
    HR = synth.hand_rays(6, J, seed=10)
    zz = torch.linspace(0.7, 1.1, 6)
    pts = (HR["rays_o"][:, None] + HR["rays_d"][:, None] * zz[None, :, None]).reshape(-1, 3)
    want, want_feat, want_r, want_h = hsdf(pts, bt, T)
    want_n = hsdf.gradient(pts.clone(), bt, T).squeeze()
    want_rgb = hcol(None, want_feat, want[:, 1:], want_h, want_n, 0)
    out, feat, r, h = O.sdf_hand_forward(sp, pts, bt, T)
    assert max_abs(feat, want_feat) < 1e-5 and max_abs(r, want_r) < 1e-5 and max_abs(h, want_h) < 1e-5
    assert max_abs(out, want) < 1e-4
    n = O.sdf_gradient(lambda q: O.sdf_hand_forward(sp, q, bt, T)[0][:, :1], pts.clone())
    assert rel_err(n, want_n) < 1e-4
    assert max_abs(O.color_hand_forward(cp, feat, out[:, 1:], n), want_rgb) < 1e-4


def _hand_nets(ref, seed, use_batch=False):
    sp, cp = synth.hand_states(seed=seed)
    emb = ref.fields.Embedding()
    hsdf = ref.fields.SDFNetwork(emb, 4, "real", use_batch=use_batch, **ref_loader.HAND_SDF_CONF)
    hcol = ref.fields.RenderingNetwork(emb, "real", **ref_loader.HAND_COLOR_CONF)
    hdev = ref.fields.SingleVarianceNetwork(ref_loader.VARIANCE_INIT)
    hsdf.load_state_dict(sp)
    hcol.load_state_dict(cp)
    return sp, cp, hsdf, hcol, hdev


def test_hand_render_and_pose_gradients_fresh_seed():
    """NeuSRenderer.render, hand branch (utils/renderer.py:190-258), 6 rays, gradients to the bone transforms and the
    T-pose joints (what pose fitting differentiates)"""
    import cases
    ref = ref_loader.load_reference()
    sp, cp, hsdf, hcol, hdev = _hand_nets(ref, 43)
    bt, T, J = synth.hand_pose(seed=11)
    B = 6
    HR = synth.hand_rays(B, J, seed=12)
    true_rgb = torch.rand(B, 3, generator=torch.Generator().manual_seed(13))
    hr = ref.renderer.NeuSRenderer(hsdf, hdev, hcol, "hand", **ref_loader.RENDERER_CONF)
    btg, Tg = bt.clone().requires_grad_(True), T.clone().requires_grad_(True)
    with _fixed_rand(HR["t_rand"]):
        want = hr.render(HR["rays_o"], HR["rays_d"], HR["near"], HR["far"], btg, Tg, None, None, None, 0)
    want_loss = cases.hand_render_loss(want, true_rgb)
    want_g = torch.autograd.grad(want_loss, [btg, Tg])
    bt2, T2 = bt.clone().requires_grad_(True), T.clone().requires_grad_(True)
    var = torch.tensor(ref_loader.VARIANCE_INIT)
    out = O.render_hand(sp, cp, var, HR["rays_o"], HR["rays_d"], HR["near"], HR["far"], bt2, T2, HR["t_rand"])
    for k in ("color_fine", "cdf_fine", "weight_sum", "weight_max"):
        assert max_abs(out[k], want[k]) < 2e-4, k
    loss = cases.hand_render_loss(out, true_rgb)
    assert rel_err(loss, want_loss) < 1e-4
    got = torch.autograd.grad(loss, [bt2, T2])
    for a, b, name in zip(got, want_g, ("bt_inv", "T_pose_21")):
        assert rel_err(a, b) < 2e-2, name                     # the golden test's bound (steep field, fp32 on both sides)


def test_fitting_render_fresh_seed():
    """NeuSRenderer_fitting.render (utils/renderer.py:434-535): two fields, shared samples, gradients to bt_inv, Ro, To"""
    import cases
    ref = ref_loader.load_reference()
    hs, hc, hsdf, hcol, hdev = _hand_nets(ref, 44)
    os_, oc, sdf, col, dev = _obj_nets(ref, 34)
    bt, T, J = synth.hand_pose(seed=14)
    B = 6
    HR = synth.hand_rays(B, J, seed=15)
    gq = torch.Generator().manual_seed(16)
    Ro0 = synth.random_rotation(gq)
    To0 = J.mean(0) + 0.02 * torch.randn(3, generator=gq)
    true_rgb = torch.rand(B, 3, generator=gq)
    fr = ref.renderer.NeuSRenderer_fitting(hsdf, hdev, hcol, sdf, dev, col, **ref_loader.RENDERER_CONF)
    btg, Ro, To = bt.clone().requires_grad_(True), Ro0.clone().requires_grad_(True), To0.clone().requires_grad_(True)
    with _fixed_rand(HR["t_rand"]):
        want = fr.render(HR["rays_o"], HR["rays_d"], HR["near"], HR["far"], btg, T, None, Ro, To)
    want_loss = cases.fit_loss(want, true_rgb)
    want_g = torch.autograd.grad(want_loss, [btg, Ro, To])
    var = torch.tensor(ref_loader.VARIANCE_INIT)
    bt2, Ro2, To2 = bt.clone().requires_grad_(True), Ro0.clone().requires_grad_(True), To0.clone().requires_grad_(True)
    out = O.fit_render((hs, hc, var), (os_, oc, var), HR["rays_o"], HR["rays_d"], HR["near"], HR["far"], bt2, T, Ro2, To2,
                       HR["t_rand"])
    # an importance sample may cross a cdf knot on a 1-ulp difference of the ray transform: north-star tolerance
    for k in ("color_fine", "weight_sum", "sdf_hand", "sdf_obj"):
        assert max_abs(out[k], want[k]) < 1e-3, k
    loss = cases.fit_loss(out, true_rgb)
    got = torch.autograd.grad(loss, [bt2, Ro2, To2])
    for a, b, name in zip(got, want_g, ("bt_inv", "Ro", "To")):
        assert rel_err(a, b) < 2e-2, name
