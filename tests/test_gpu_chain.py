"""Fused tile-chain kernels (HN_TC_BF16X3: split bf16 operands on tcgen05, activations resident in shared
memory across layers) against the fp32 SIMT verification path and the fp64 oracle."""
import pytest
import torch

import analytic as A
import synth
from golden_util import max_abs
from gpu_util import DEV, obj_modules

pytestmark = pytest.mark.gpu


def _pts(n, seed=0):
    g = torch.Generator().manual_seed(seed)
    return 0.45 * torch.randn(n, 3, generator=g)


@pytest.mark.parametrize("n", [1, 127, 128, 129, 1000, 40000])
def test_sdf_only_chain_vs_simt_and_fp64(n):
    """SDFNetwork_OBJ.sdf through the chain kernel: <= 5e-5 abs vs fp64 (observed ~1e-5; single-pass bf16
    would be ~6e-3), ragged tile tails included."""
    import honerf_b200 as H
    sdf, _, _, sp, _ = obj_modules(requires_grad=False)
    x = _pts(n, seed=n)
    got = H.ops.sdf_obj_sdf_only(sdf.packed(), x.to(DEV), 1.0, precision=H.ops._PRECISIONS["tc_bf16x3"])
    simt = H.ops.sdf_obj_sdf_only(sdf.packed(), x.to(DEV), 1.0, precision=H.ops._PRECISIONS["simt_fp32"])
    spd = {k: v.double() for k, v in sp.items() if k != "se3_refine"}
    Ws, bs = A.effective_weights(spd)
    ref = A.sdf_obj_fwd(Ws, bs, x.double())[0]
    e_ref, e_simt = max_abs(got, ref), max_abs(got, simt)
    print("n=%d  chain vs fp64 %.2e  chain vs simt %.2e  simt vs fp64 %.2e" % (n, e_ref, e_simt, max_abs(simt, ref)))
    assert torch.isfinite(got).all()
    assert e_ref < 5e-5 and e_simt < 5e-5


def test_sdf_only_chain_geometric_init_and_scale():
    """Un-perturbed geometric-init weights (most of lin0 / skip columns are zero) and scale != 1."""
    import honerf_b200 as H
    import ref_conf
    torch.manual_seed(3)
    net = H.SDFNetwork_OBJ(H.Embedding(), 4, "real", **dict(ref_conf.OBJ_SDF_CONF, scale=2.0)).to(DEV)
    x = _pts(5000, seed=9).to(DEV)
    a = H.ops.sdf_obj_sdf_only(net.packed(), x, 0.5, precision=H.ops._PRECISIONS["tc_bf16x3"])
    b = H.ops.sdf_obj_sdf_only(net.packed(), x, 0.5, precision=H.ops._PRECISIONS["simt_fp32"])
    assert max_abs(a, b) < 2e-5
