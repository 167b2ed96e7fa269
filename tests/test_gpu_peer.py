"""The step's gradient exchange fused with Adam over peer memory (hn_peer_adam_flat, csrc/peer.cu) against
hn_adam_flat on the summed gradient.

* one process, world 1: the kernel's two phases and its Adam arithmetic, through the C ABI;
* 2, 4 and 8 PROCESSES on one GPU (IPC handles exchanged over gloo, every rank on cuda:0): the cross-process protocol --
  handle exchange, barrier flags, two-shot ownership, CUDA-graph replay -- on the single-GPU test tier;
* two GPUs over NCCL / NVLink when the box has them (skipped otherwise).
Every element is summed once, in rank order, by its owner and broadcast: the reduced gradient equals the rank-ordered
fp32 sum exactly and the parameters are BIT-identical on every rank; against hn_adam_flat on that sum they agree to fp32
rounding of the update (1e-6)."""
import ctypes
import os
import socket

import pytest
import torch

from gpu_util import DEV

pytestmark = pytest.mark.gpu


def test_peer_kernel_world1_matches_adam_flat():
    from honerf_b200._lib import check, lib
    n = 100_000
    g = torch.Generator().manual_seed(11)
    p0 = torch.randn(n, generator=g).to(DEV)
    grads = [torch.randn(n, generator=g).to(DEV) * 10.0 ** (k - 1) for k in range(3)]
    nbytes = lib.hn_peer_block_bytes(n, 1)
    ptr, handle = ctypes.c_void_p(), (ctypes.c_uint8 * 64)()
    check(lib.hn_peer_alloc(nbytes, ctypes.byref(ptr), handle), "hn_peer_alloc")
    try:
        blocks = (ctypes.c_void_p * 1)(ptr.value)
        epoch = torch.zeros(148, device=DEV, dtype=torch.int32)
        err = torch.zeros(1, device=DEV, dtype=torch.int32)
        step = torch.zeros(1, device=DEV)
        lr = torch.full((1,), 1e-2, device=DEV)
        pa, ma, va = p0.clone(), torch.zeros_like(p0), torch.zeros_like(p0)
        pb, mb, vb = p0.clone(), torch.zeros_like(p0), torch.zeros_like(p0)
        v = lambda t: ctypes.c_void_p(t.data_ptr())
        stream = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
        from honerf_b200.optim import _DeviceArray
        gbuf = torch.as_tensor(_DeviceArray(ptr.value, n), device=DEV)
        assert gbuf.data_ptr() == ptr.value and gbuf.numel() == n       # a view of the block, not a copy
        for k, gr in enumerate(grads):
            step += 1.0
            gbuf.copy_(gr)
            check(lib.hn_peer_adam_flat(v(pa), v(ma), v(va), n, blocks, 0, 1, v(epoch), v(err), 0, v(step), v(lr), 1e-2, 0.9, 0.999,
                                        1e-8, 1e-2, 0.5, stream), "hn_peer_adam_flat")
            check(lib.hn_adam_flat(v(pb), v(gr), v(mb), v(vb), n, v(step), v(lr), None, 1e-2, 0.9, 0.999, 1e-8, 1e-2, 0.5, stream),
                  "hn_adam_flat")
        out = torch.empty_like(p0)
        check(lib.hn_peer_adam_flat(v(out), None, None, n, blocks, 0, 1, v(epoch), v(err), 1, None, None, 0.0, 0.0, 0.0, 0.0, 0.0,
                                    0.25, stream), "hn_peer_adam_flat(mode 1)")
        torch.cuda.synchronize()
        assert int(err.item()) == 0 and int(epoch[0].item()) == 8        # two barriers per launch, four launches
        assert torch.equal(out, grads[-1] * 0.25)
        for a, b in ((pa, pb), (ma, mb), (va, vb)):
            assert float((a - b).abs().max() / b.abs().max()) < 1e-6
        del gbuf
    finally:
        torch.cuda.synchronize()
        lib.hn_peer_free(ptr)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, backend, one_gpu, out):
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dev = torch.device("cuda", 0 if one_gpu else rank)
    torch.cuda.set_device(dev)
    if backend == "nccl":
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    else:
        dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from honerf_b200.optim import FlatAdam
        g = torch.Generator().manual_seed(21)
        shapes = [(257, 39), (257,), (3,), (64, 3)]
        init = [torch.randn(*s, generator=g) for s in shapes]
        pa = [torch.nn.Parameter(t.clone().to(dev)) for t in init]      # peer exchange
        pb = [torch.nn.Parameter(t.clone().to(dev)) for t in init]      # hn_adam_flat on the summed gradient
        oa, ob = FlatAdam(pa, lr=1e-2), FlatAdam(pb, lr=1e-2)
        ok = oa.enable_peer_exchange()
        if not ok:
            out[rank] = ("setup-failed",)
            return
        n_steps = 6
        # gradients of every rank, known to every rank (seeded): parameter 2 has no gradient on rank 1
        grads = [[[torch.randn(*s, generator=g).to(dev) * (r + 1) for s in shapes] for r in range(world)] for _ in range(n_steps)]
        static = [torch.zeros(*s, device=dev) for s in shapes]

        def set_static(k):
            for i, t in enumerate(static):
                t.copy_(grads[k][rank][i])

        def peer_step():
            for i, p in enumerate(pa):
                p.grad = None if (i == 2 and rank == 1) else static[i]
            oa.step(oa.gather_grads(), grad_scale=1.0 / world, peer_exchange=True)

        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for k in range(2):                       # eager
                set_static(k)
                peer_step()
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph):
                peer_step()
        torch.cuda.current_stream().wait_stream(side)
        for k in range(2, n_steps):                  # replayed
            set_static(k)
            graph.replay()
        reduced = oa.peer_allreduce(grad_scale=1.0)
        torch.cuda.synchronize()
        for k in range(n_steps):
            for i, p in enumerate(pb):
                p.grad = sum(grads[k][r][i] for r in range(world) if not (i == 2 and r == 1))
            ob.step(grad_scale=1.0 / world)
        want_reduced = sum(torch.cat([(torch.zeros_like(grads[-1][r][i]) if (i == 2 and r == 1) else grads[-1][r][i]).reshape(-1)
                                      for i in range(len(shapes))]) for r in range(world))
        got = torch.cat([reduced[off:off + p.numel()] for p, off in zip(pa, oa.offsets)])
        errs = [float((a - b).abs().max() / b.abs().max()) for a, b in zip(pa, pb)]
        flat_after = oa.flat.detach().cpu().clone()          # the parameters after the exchanged steps (compared across ranks)
        err = oa.peer_error()
        # leaving the exchange: flat_grad keeps its values in private memory and the NCCL-form step works on it
        before = oa.flat_grad.clone()
        oa.close_peer_exchange()
        closed_ok = oa._peer is None and torch.equal(oa.flat_grad, before) and oa.flat_grad.data_ptr() != before.data_ptr()
        for i, p in enumerate(pa):
            p.grad = static[i]
        oa.step(grad_scale=1.0)
        torch.cuda.synchronize()
        out[rank] = ("ok" if closed_ok else "close-failed", err, errs, flat_after, float((got - want_reduced).abs().max()))
        dist.barrier()
    finally:
        torch.cuda.synchronize()
        dist.destroy_process_group()


def _run(world, backend, one_gpu):
    import torch.multiprocessing as mp
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), backend, one_gpu, out), nprocs=world, join=True)
    for r in range(world):
        assert out[r][0] == "ok", out[r]
        _, err, errs, flat, red_err = out[r]
        assert err == 0
        assert max(errs) < 1e-6, errs
        assert red_err == 0.0                              # one summation order (rank 0, 1, 2, ...) for every element
    for r in range(1, world):
        assert torch.equal(out[0][3], out[r][3])         # bit-identical parameters on every rank


@pytest.mark.timeout(600)
@pytest.mark.parametrize("world", [2, 4, 8])
def test_peer_exchange_processes_on_one_gpu(world):
    """world processes time-share cuda:0 (each barrier needs every context scheduled once: slow, but the same protocol as
    one process per GPU -- world 8 covers the second group of four ranks in the reduction and the chunk search)"""
    _run(world, "gloo", True)


@pytest.mark.timeout(300)
def test_peer_exchange_two_gpus_nccl():
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    _run(2, "nccl", False)
