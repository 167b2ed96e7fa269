"""FlatAdam (hn_adam_flat) against torch.optim.Adam on the same parameters and gradients."""
import pytest
import torch

from gpu_util import DEV

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("weight_decay", [0.0, 1e-2])
def test_flat_adam_matches_torch_adam(weight_decay):
    """10 steps on 5 tensors of odd sizes, one of which never receives a gradient (torch skips it entirely):
    parameters equal to 1e-6 relative (fp32 rounding of a different but equivalent operation order)."""
    from honerf_b200.optim import FlatAdam
    g = torch.Generator().manual_seed(3)
    shapes = [(257, 39), (257,), (1,), (64, 3), (5, 1)]
    init = [torch.randn(*s, generator=g) for s in shapes]
    pa = [torch.nn.Parameter(t.clone().to(DEV)) for t in init]
    pb = [torch.nn.Parameter(t.clone().to(DEV)) for t in init]
    oa = FlatAdam(pa, lr=1e-2, weight_decay=weight_decay)
    ob = torch.optim.Adam(pb, lr=1e-2, weight_decay=weight_decay)
    for p, t in zip(pa, init):                       # re-homed, values intact
        assert torch.equal(p.detach().cpu(), t) and p.data_ptr() >= oa.flat.data_ptr()
    for step in range(10):
        grads = [torch.randn(*s, generator=g).to(DEV) * (10.0 ** (step % 3 - 1)) for s in shapes]
        oa.zero_grad(); ob.zero_grad()
        for i, (a, b, gr) in enumerate(zip(pa, pb, grads)):
            if i == 2:
                continue                             # this parameter has no gradient
            a.grad = gr.clone(); b.grad = gr.clone()
        v0 = pa[0]._version
        oa.step(); ob.step()
        assert pa[0]._version > v0                   # in-place update is visible to version-keyed caches
    for i, (a, b) in enumerate(zip(pa, pb)):
        err = float((a - b).abs().max() / b.abs().max())
        assert err < 1e-6, (i, err)
    assert torch.equal(pa[2].detach().cpu(), init[2])


def test_flat_adam_grad_scale_and_graph_replay():
    """grad_scale folds the 1/world averaging; the whole step replays from a CUDA graph (device-side step count)."""
    from honerf_b200.optim import FlatAdam
    g = torch.Generator().manual_seed(5)
    init = torch.randn(1000, generator=g)
    pa, pb = torch.nn.Parameter(init.clone().to(DEV)), torch.nn.Parameter(init.clone().to(DEV))
    oa, ob = FlatAdam([pa], lr=1e-3), torch.optim.Adam([pb], lr=1e-3)
    gstatic = torch.zeros(1000, device=DEV)
    pa.grad = gstatic
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        runs = oa.gather_grads()
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            oa.gather_grads()
            oa.step(runs, grad_scale=0.5)
    torch.cuda.current_stream().wait_stream(side)
    for step in range(5):
        gr = torch.randn(1000, generator=g).to(DEV)
        gstatic.copy_(2.0 * gr)
        graph.replay()
        pb.grad = gr
        ob.step()
    torch.cuda.synchronize()
    assert float((pa - pb).abs().max() / pb.abs().max()) < 1e-6
