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
        assert n == 22
        assert torch.allclose(g0, torch.full((5, 3), 1.5))
        assert torch.allclose(g1, torch.arange(7.0) * 1.5)
        assert g2 is None
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
        assert n == 9 and g1 is None
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
