"""world_size-2 gloo test of the multi-GPU plumbing (ray sharding + flat gradient all-reduce)."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from honerf_b200 import dist as hdist
    torch.manual_seed(0)
    params = [torch.nn.Parameter(torch.randn(5, 3)), torch.nn.Parameter(torch.randn(7)),
              torch.nn.Parameter(torch.randn(2))]
    # rank-dependent gradients on the first two, none on the third
    params[0].grad = torch.full((5, 3), float(rank + 1))
    params[1].grad = torch.arange(7.0) * (rank + 1)
    n = hdist.allreduce_gradients(params, world)
    lo, hi = hdist.shard_rays(11, rank, world)
    out[rank] = (n, params[0].grad.clone(), params[1].grad.clone(), params[2].grad, (lo, hi))
    dist.destroy_process_group()


def test_flat_gradient_allreduce_gloo_world2():
    world = 2
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), out), nprocs=world, join=True)
    for r in range(world):
        n, g0, g1, g2, shard = out[r]
        assert n == 24                     # EVERY parameter of the list is sent (zeros where a rank has no gradient)
        assert torch.allclose(g0, torch.full((5, 3), 1.5))
        assert torch.allclose(g1, torch.arange(7.0) * 1.5)
        assert g2 is not None and float(g2.abs().max()) == 0.0
    assert out[0][4] == (0, 6) and out[1][4] == (6, 11)


def _worker_split(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from honerf_b200 import dist as hdist
    params = [torch.nn.Parameter(torch.zeros(4, 2)), torch.nn.Parameter(torch.zeros(3)), torch.nn.Parameter(torch.zeros(1))]
    params[0].grad = torch.full((4, 2), float(rank + 1))
    params[2].grad = torch.tensor([10.0 * (rank + 1)])
    flat = hdist.flatten_gradients(params)          # what bench.py captures in graph A
    hdist.allreduce_flat(flat)                      # the eager collective
    n = hdist.unflatten_gradients(params, flat, world)      # graph B
    out[rank] = (n, params[0].grad.clone(), params[1].grad, params[2].grad.clone())
    dist.destroy_process_group()


def test_split_flatten_allreduce_unflatten_gloo_world2():
    """The three-phase form bench.py uses at N > 1 (pack in graph A, eager all-reduce, unpack + Adam in graph B)
    gives the same averaged gradients as the one-call form."""
    world = 2
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker_split, args=(world, _free_port(), out), nprocs=world, join=True)
    for r in range(world):
        n, g0, g1, g2 = out[r]
        assert n == 12 and g1 is not None and float(g1.abs().max()) == 0.0
        assert torch.allclose(g0, torch.full((4, 2), 1.5)) and torch.allclose(g2, torch.tensor([15.0]))


def _fit_worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from honerf_b200 import dist as hdist
    n_views = 8
    lo, hi = hdist.shard_views(n_views, rank, world)
    # pose parameters of fitting: 45 hand-pose values per frame + object rotation / translation; every view adds a
    # known gradient, so the SUM over ranks must equal the sum over all views whatever the sharding
    pose = torch.nn.Parameter(torch.zeros(45))
    obj = torch.nn.Parameter(torch.zeros(9))
    pose.grad = sum(torch.full((45,), float(v + 1)) for v in range(lo, hi))
    obj.grad = sum(torch.arange(9.0) * (v + 1) for v in range(lo, hi))
    n = hdist.allreduce_gradients([pose, obj], world, average=False)
    out[rank] = (n, (lo, hi), pose.grad.clone(), obj.grad.clone())
    dist.destroy_process_group()


def test_view_sharded_pose_gradient_sum_gloo_world2():
    world = 2
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_fit_worker, args=(world, _free_port(), out), nprocs=world, join=True)
    total = sum(range(1, 9))
    assert [out[r][1] for r in range(world)] == [(0, 4), (4, 8)]
    for r in range(world):
        n, _, gp, go = out[r]
        assert n == 54
        assert torch.equal(gp, torch.full((45,), float(total))) and torch.equal(go, torch.arange(9.0) * total)


def _loss_worker(rank, world, port, out, n_total):
    """Ray-sharded training loss with GLOBAL normalisers (honerf_b200.dist.loss_normalisers) and SUMMED gradients against
    the single-process loss of exp_runner.py:206-227 on the whole batch -- unequal shards, unequal mask counts, a rank
    whose parameters do not all receive a gradient, and (n_total < world) an empty shard."""
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle"))
    import analytic as A
    from honerf_b200 import dist as hdist
    g = torch.Generator().manual_seed(5)
    w1 = torch.nn.Parameter(torch.randn(3, 3, generator=g))
    w2 = torch.nn.Parameter(torch.randn(3, generator=g))
    w3 = torch.nn.Parameter(torch.randn(2, generator=g))          # only touched by rays with index >= 7
    feats = torch.randn(n_total, 3, generator=g)
    rgb = torch.rand(n_total, 3, generator=g)
    mask = (torch.rand(n_total, 1, generator=g) > 0.3).float()

    def render(lo, hi):
        x = feats[lo:hi]
        color = torch.sigmoid(x @ w1)
        wsum = torch.sigmoid(x @ w2)[:, None]
        eik = ((x * w2).sum(-1) ** 2)
        if hi > 7:
            sel = torch.arange(lo, hi) >= 7
            color = color + (sel[:, None] * w3.sum())
        return color, wsum, eik.mean() if hi > lo else x.sum() * 0.0
    # single process
    c, ws, ge = render(0, n_total)
    ref = A.render_loss_closed_form(c, ws, rgb, mask, ge, 0.0, 1.0, 0.7, 0.3)[0]
    ref_g = torch.autograd.grad(ref, [w1, w2, w3], allow_unused=True)
    # sharded
    lo, hi = hdist.shard_rays(n_total, rank, world)
    div, share = hdist.loss_normalisers(mask[lo:hi], n_total)
    params = [w1, w2, w3]
    if hi > lo:
        c, ws, ge = render(lo, hi)
        loss = A.render_loss_closed_form(c, ws, rgb[lo:hi], mask[lo:hi], ge, float(div), 1.0, 0.7 * share, 0.3 * share)[0]
        # the closed form divides its mask term by the LOCAL ray count: share * (sum / n_local) = sum / n_total
        loss.backward()
    else:
        loss = torch.zeros(())
    n_sent = hdist.allreduce_gradients(params, world, average=False)
    tot = loss.detach().clone().reshape(1)
    dist.all_reduce(tot)
    out[rank] = (n_sent, (lo, hi), float(tot), float(ref), [p.grad.clone() for p in params],
                 [g_ if g_ is not None else torch.zeros_like(p) for g_, p in zip(ref_g, params)])
    dist.destroy_process_group()


def test_sharded_loss_with_global_normalisers_equals_single_process_gloo_world2():
    for n_total in (11, 8, 1):
        world = 2
        mgr = mp.Manager()
        out = mgr.dict()
        mp.spawn(_loss_worker, args=(world, _free_port(), out, n_total), nprocs=world, join=True)
        shards = [out[r][1] for r in range(world)]
        assert shards[0][0] == 0 and shards[-1][1] == n_total and shards[0][1] == shards[1][0]
        assert abs((shards[0][1] - shards[0][0]) - (shards[1][1] - shards[1][0])) <= 1          # balanced
        for r in range(world):
            n_sent, _, tot, ref, grads, ref_g = out[r]
            assert n_sent == 14
            assert abs(tot - ref) <= 1e-5 * max(1.0, abs(ref)), (n_total, tot, ref)
            for a, b in zip(grads, ref_g):
                assert torch.allclose(a, b, rtol=1e-4, atol=1e-6), (n_total, a, b)


def _gather_worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from honerf_b200 import dist as hdist
    res = 7
    sizes = [hdist.shard_rays(res, r, world) for r in range(world)]
    lo, hi = sizes[rank]
    full = torch.arange(res * 3 * 2, dtype=torch.float32).reshape(res, 3, 2)
    got = hdist.gather_slabs(full[lo:hi].clone(), [b - a for a, b in sizes], dim=0)
    out[rank] = torch.equal(got, full)
    dist.destroy_process_group()


def test_gather_of_lattice_slabs_gloo_world2():
    """extract_geometry sharded in x slabs (BASELINE configs[1]): the ranks' slabs of `u` gathered into the full lattice."""
    world = 2
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_gather_worker, args=(world, _free_port(), out), nprocs=world, join=True)
    assert all(out[r] for r in range(world))
