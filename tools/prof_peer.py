"""Device time of the step's gradient exchange + Adam, NCCL form against the peer-memory kernel (csrc/peer.cu).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 tools/prof_peer.py

Every variant is captured as a CUDA graph of 50 launches (no CPU launch cost in the number) and replayed; the ranks are
lined up by a barrier before each timed replay, time = max over ranks / launches."""
import json
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    sys.stdout.flush()
    json_fd = os.dup(1)             # stdout carries the JSON line only (NCCL prints its banner there)
    os.dup2(2, 1)
    dev = torch.device("cuda", int(os.environ.get("LOCAL_RANK", "0")))
    torch.cuda.set_device(dev)
    dist.init_process_group("nccl", device_id=dev)
    from honerf_b200.optim import FlatAdam
    n = int(os.environ.get("PROF_PEER_N", "824064"))
    p = [torch.nn.Parameter(torch.randn(n, device=dev))]
    opt = FlatAdam(p, lr=1e-4)
    assert opt.enable_peer_exchange()
    opt.flat_grad.copy_(torch.randn(n, device=dev))
    p[0].grad = None
    runs = [[0, 0]]
    res = {}

    def timed(name, fn, launches=50, reps=5):
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            fn()
            torch.cuda.synchronize()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                for _ in range(launches):
                    fn()
        torch.cuda.current_stream().wait_stream(side)
        g.replay()
        best = 1e9
        for _ in range(reps):
            torch.cuda.synchronize()
            dist.barrier()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            g.replay()
            e1.record()
            torch.cuda.synchronize()
            t = torch.tensor([e0.elapsed_time(e1)], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            best = min(best, float(t) / launches * 1e3)
        res[name] = round(best, 2)

    def nccl_step():
        dist.all_reduce(opt.flat_grad, op=dist.ReduceOp.SUM)
        opt.step(runs, grad_scale=1.0 / world)

    timed("nccl_allreduce+hn_adam_flat_us", nccl_step)
    timed("hn_adam_flat_alone_us", lambda: opt.step(runs, grad_scale=1.0 / world))
    for ctas in (148, 64, 16):
        os.environ["HONERF_PEER_CTAS"] = str(ctas)
        timed("hn_peer_adam_flat_%d_ctas_us" % ctas, lambda: opt.step(runs, grad_scale=1.0 / world, peer_exchange=True))
    os.environ.pop("HONERF_PEER_CTAS")
    # one rank arrives ~200 us late (a spinning kernel of 400k cycles on rank 0 before every exchange)
    def late(fn):
        def run():
            if rank == 0:
                torch.cuda._sleep(400000)
            fn()
        return run
    timed("late_rank:nccl_allreduce+hn_adam_flat_us", late(nccl_step), launches=20)
    timed("late_rank:hn_peer_adam_flat_us", late(lambda: opt.step(runs, grad_scale=1.0 / world, peer_exchange=True)), launches=20)
    timed("late_rank:sleep_alone_us", late(lambda: None), launches=20)
    # the end-to-end pattern of bench.py: replay, then a blocking device -> host read, every step (host in the loop)
    import time
    small = torch.zeros(1, device=dev)
    host = torch.empty(1, pin_memory=True)

    def host_loop(name, fn, eager_kernels=0, iters=40):
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            fn()
            torch.cuda.synchronize()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                torch.cuda._sleep(100000)
                fn()
                small.add_(1.0)
        torch.cuda.current_stream().wait_stream(side)
        best = 1e9
        for _ in range(3):
            torch.cuda.synchronize()
            dist.barrier()
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            for _ in range(iters):
                if eager_kernels:
                    for _ in range(eager_kernels):
                        small.add_(1.0)
                    fn()
                else:
                    g.replay()
                    host.copy_(small, non_blocking=False)
            torch.cuda.synchronize()
            t = torch.tensor([(time.perf_counter() - t0) / iters * 1e6], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            best = min(best, float(t))
        res[name] = round(best, 1)

    peer_step = lambda: opt.step(runs, grad_scale=1.0 / world, peer_exchange=True)
    host_loop("host_in_loop:sleep50us+nccl+adam+d2h_us", nccl_step)
    host_loop("host_in_loop:sleep50us+peer+d2h_us", peer_step)
    host_loop("host_in_loop:sleep50us+adam_only+d2h_us", lambda: opt.step(runs, grad_scale=1.0 / world))
    host_loop("eager:60_small_kernels+nccl+adam_us", nccl_step, eager_kernels=60)
    host_loop("eager:60_small_kernels+peer_us", peer_step, eager_kernels=60)
    host_loop("eager:60_small_kernels+adam_only_us", lambda: opt.step(runs, grad_scale=1.0 / world), eager_kernels=60)
    res["peer_error"] = opt.peer_error()
    if rank == 0:
        os.write(json_fd, (json.dumps({"world": world, "n_floats": n, **res}) + "\n").encode())
    dist.barrier()
    torch.cuda.synchronize()
    os._exit(0)


if __name__ == "__main__":
    main()
