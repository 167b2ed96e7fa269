"""render_core_outside (csrc/outside.cu, NeuSRenderer.render_core_outside): the background branch the north star names.  The
HO-NeRF reference has no such method (utils/renderer.py:47,56 only store n_outside, every config sets 0): PARITY UNPINNED, the
oracle (oracle/honerf_oracle.render_core_outside) restates the NeuS semantics.  Checked against it in fp64: outputs 1e-5 abs,
gradients to the NeRF's parameters (through all four outputs, random cotangents: fp32 transmittance products against fp64) 2e-3
relative (observed 4e-4)."""
import pytest
import torch

import honerf_oracle as O
import synth
from golden_util import max_abs, rel_l2
from gpu_util import DEV

pytestmark = pytest.mark.gpu


class _TinyNerf(torch.nn.Module):
    """Stand-in for the caller's background NeRF: any module mapping (pts4, dirs) -> (density [N,1], raw rgb [N,3])."""

    def __init__(self, d_in=4):
        super().__init__()
        g = torch.Generator().manual_seed(3)
        self.w1 = torch.nn.Parameter(torch.randn(d_in + 3, 32, generator=g) * 0.7)
        self.w2 = torch.nn.Parameter(torch.randn(32, 4, generator=g) * 0.7)

    def forward(self, pts, dirs):
        h = torch.tanh(torch.cat([pts, dirs], -1) @ self.w1) @ self.w2
        return h[:, :1] * 3.0, h[:, 1:]


@pytest.mark.parametrize("n_rays,n,background", [(37, 160, True), (5, 32, False), (300, 97, True), (2, 2, False)])
def test_render_core_outside_vs_oracle(n_rays, n, background):
    import honerf_b200 as H
    import ref_conf
    from gpu_util import obj_modules
    sdf, col, var, _, _ = obj_modules(requires_grad=False)
    r = H.NeuSRenderer(sdf, var, col, "obj", **dict(ref_conf.RENDERER_CONF, n_outside=32))
    R = synth.object_rays(n_rays, seed=5)
    g = torch.Generator().manual_seed(n)
    z = torch.sort(0.5 + 4.0 * torch.rand(n_rays, n, generator=g), dim=-1).values
    bg = torch.tensor([0.2, 0.5, 0.9]) if background else None
    gc, gs = torch.randn(n_rays, 3, generator=g), torch.randn(n_rays, n, 3, generator=g)
    ga, gw = torch.randn(n_rays, n, generator=g), torch.randn(n_rays, n, generator=g)
    # fp64 oracle
    nerf64 = _TinyNerf().double()
    ref = O.render_core_outside(R["rays_o"].double(), R["rays_d"].double(), z.double(), 0.03, nerf64, 32,
                                bg.double() if background else None)
    L = (ref["color"] * gc.double()).sum() + (ref["sampled_color"] * gs.double()).sum() + (ref["alpha"] * ga.double()).sum() + \
        (ref["weights"] * gw.double()).sum()
    gref = torch.autograd.grad(L, list(nerf64.parameters()))
    # product
    nerf = _TinyNerf().to(DEV)
    out = r.render_core_outside(R["rays_o"].to(DEV), R["rays_d"].to(DEV), z.to(DEV), 0.03, nerf, bg.to(DEV) if background else None)
    for k in ("color", "sampled_color", "alpha", "weights"):
        assert out[k].shape == ref[k].shape and max_abs(out[k], ref[k]) < 1e-5, (k, max_abs(out[k], ref[k]))
    Lg = (out["color"] * gc.to(DEV)).sum() + (out["sampled_color"] * gs.to(DEV)).sum() + (out["alpha"] * ga.to(DEV)).sum() + \
        (out["weights"] * gw.to(DEV)).sum()
    got = torch.autograd.grad(Lg, list(nerf.parameters()))
    for a, b in zip(got, gref):
        assert rel_l2(a, b) < 2e-3, rel_l2(a, b)
    # only the colour is differentiated (the other cotangents arrive as None)
    out = r.render_core_outside(R["rays_o"].to(DEV), R["rays_d"].to(DEV), z.to(DEV), 0.03, nerf, bg.to(DEV) if background else None)
    g1 = torch.autograd.grad((out["color"] * gc.to(DEV)).sum(), list(nerf.parameters()))
    r1 = torch.autograd.grad((O.render_core_outside(R["rays_o"].double(), R["rays_d"].double(), z.double(), 0.03, nerf64, 32,
                                                    bg.double() if background else None)["color"] * gc.double()).sum(),
                             list(nerf64.parameters()))
    for a, b in zip(g1, r1):
        assert rel_l2(a, b) < 2e-3


def test_weights_are_a_sub_probability_and_saturated_density_is_opaque():
    import honerf_b200 as H
    B, n = 16, 64
    dists = torch.full((B, n), 0.05, device=DEV)
    dens = torch.full((B, n), 50.0, device=DEV)
    raw = torch.zeros(B, n, 3, device=DEV)
    color, sampled, alpha, weights = H.ops.outside_composite(dens, raw, dists, torch.tensor([1.0, 1.0, 1.0], device=DEV))
    assert float(weights.sum(-1).max()) <= 1.0 + 1e-5 and float(weights.sum(-1).min()) > 0.99
    assert max_abs(color, torch.full((B, 3), 0.5)) < 1e-2          # sigmoid(0) = 0.5, no background shows through
    color, _, _, weights = H.ops.outside_composite(torch.full((B, n), -50.0, device=DEV), raw, dists,
                                                   torch.tensor([0.1, 0.2, 0.3], device=DEV))
    assert float(weights.abs().max()) < 1e-6 and max_abs(color, torch.tensor([0.1, 0.2, 0.3]).expand(B, 3)) < 1e-5
