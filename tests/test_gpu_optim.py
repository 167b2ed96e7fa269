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


def test_flat_adam_lr_schedule_takes_effect_under_graph_replay():
    """exp_runner.py:update_learning_rate rewrites optimizer.param_groups[i]['lr'] every iteration.  The learning rate
    is a device scalar: a step captured in a CUDA graph follows the schedule when sync_lr() runs before each replay."""
    from honerf_b200.optim import FlatAdam
    g = torch.Generator().manual_seed(6)
    init = torch.randn(777, generator=g)
    pa, pb = torch.nn.Parameter(init.clone().to(DEV)), torch.nn.Parameter(init.clone().to(DEV))
    oa, ob = FlatAdam([pa], lr=1e-3), torch.optim.Adam([pb], lr=1e-3)
    assert oa.param_groups[0]["lr"] == 1e-3 and oa.param_groups[0]["params"][0] is pa
    gstatic = torch.zeros(777, device=DEV)
    pa.grad = gstatic
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        runs = oa.gather_grads()
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            oa.gather_grads()
            oa.step(runs)
    torch.cuda.current_stream().wait_stream(side)
    for step in range(6):
        lr = 1e-3 * (0.5 ** step)                      # the caller's schedule
        for grp in oa.param_groups:
            grp["lr"] = lr
        for grp in ob.param_groups:
            grp["lr"] = lr
        oa.sync_lr()
        gr = torch.randn(777, generator=g).to(DEV)
        gstatic.copy_(gr)
        graph.replay()
        pb.grad = gr
        ob.step()
    torch.cuda.synchronize()
    assert float((pa - pb).abs().max() / pb.abs().max()) < 1e-6


def test_flat_adam_per_parameter_step_and_state_dict_round_trip():
    """A parameter that only sometimes receives a gradient keeps its own step count (torch's per-parameter `step`), and
    state_dict() / load_state_dict() speak torch.optim.Adam's format in both directions."""
    from honerf_b200.optim import FlatAdam
    g = torch.Generator().manual_seed(8)
    shapes = [(33, 7), (5,), (64,)]
    init = [torch.randn(*s, generator=g) for s in shapes]
    pa = [torch.nn.Parameter(t.clone().to(DEV)) for t in init]
    pb = [torch.nn.Parameter(t.clone().to(DEV)) for t in init]
    oa, ob = FlatAdam(pa, lr=3e-3), torch.optim.Adam(pb, lr=3e-3)

    def steps(oa, ob, pa, pb, n, first):
        for step in range(first, first + n):
            oa.zero_grad(); ob.zero_grad()
            for i, (a, b, s) in enumerate(zip(pa, pb, shapes)):
                if i == 1 and step % 3 != 0:
                    continue                           # parameter 1 has a gradient every third step only
                gr = torch.randn(*s, generator=g).to(DEV)
                a.grad = gr.clone(); b.grad = gr.clone()
            oa.step(); ob.step()
    steps(oa, ob, pa, pb, 7, 0)
    for a, b in zip(pa, pb):
        assert float((a - b).abs().max() / b.abs().max()) < 1e-6
    sd = oa.state_dict()
    ref = ob.state_dict()
    assert set(sd["state"]) == set(ref["state"]) and sd["param_groups"][0]["lr"] == ref["param_groups"][0]["lr"]
    for i in sd["state"]:
        assert float(sd["state"][i]["step"]) == float(ref["state"][i]["step"])
        assert torch.allclose(sd["state"][i]["exp_avg"], ref["state"][i]["exp_avg"], rtol=1e-5, atol=1e-8)
        assert torch.allclose(sd["state"][i]["exp_avg_sq"], ref["state"][i]["exp_avg_sq"], rtol=1e-5, atol=1e-10)
    # FlatAdam -> torch.optim.Adam and torch.optim.Adam -> FlatAdam, then keep training on both sides
    pc = [torch.nn.Parameter(a.detach().clone()) for a in pa]
    pd = [torch.nn.Parameter(b.detach().clone()) for b in pb]
    oc = torch.optim.Adam(pc, lr=1.0)
    oc.load_state_dict(sd)
    od = FlatAdam(pd, lr=1.0)
    od.load_state_dict(ref)
    assert oc.param_groups[0]["lr"] == 3e-3 and od.param_groups[0]["lr"] == 3e-3
    steps(od, oc, pd, pc, 5, 7)
    for a, b in zip(pd, pc):
        assert float((a - b).abs().max() / b.abs().max()) < 1e-6
