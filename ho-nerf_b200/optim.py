"""Adam over ONE flat parameter buffer (``hn_adam_flat``): the optimiser step of the train loop in one launch.

The reference builds a single ``torch.optim.Adam`` over every network's parameters (exp_runner.py:83-90); with the
weight-normalised MLPs that is ~90 small tensors per step.  ``FlatAdam`` re-homes the parameters as views of one
contiguous buffer (their values, shapes and the modules' ``state_dict`` are unchanged), gathers the step's gradients into a
second flat buffer and updates everything with one kernel.  Same update rule as ``torch.optim.Adam`` (bias correction
with a per-parameter step count, ``eps`` outside the square root, L2 ``weight_decay``).

Drop-in surface the reference's training loop uses (exp_runner.py): ``param_groups`` (``update_learning_rate`` writes
``param_groups[i]['lr']`` every iteration, the logger reads ``param_groups[0]['lr']``), ``zero_grad``, ``step``,
``state_dict`` / ``load_state_dict`` in ``torch.optim.Adam``'s own format (checkpoints keep working in both directions).
The step count and the learning rate live on the device, so the step can be captured in a CUDA graph and the schedule
still applies on replay: ``step()`` pushes ``param_groups[0]['lr']`` to the device scalar whenever it is called outside a
capture; a loop that only replays a graph calls ``sync_lr()`` before the replay.  ``flat_grad`` is also the payload of
the multi-GPU gradient all-reduce.
"""
import ctypes

import torch

from ._lib import check, lib


class _DeviceArray:
    """n fp32 elements at a raw device address, as the __cuda_array_interface__ torch.as_tensor maps without a copy"""

    def __init__(self, ptr, n):
        self.__cuda_array_interface__ = {"shape": (int(n),), "typestr": "<f4", "data": (int(ptr), False), "version": 2}


class FlatAdam:
    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0):
        self.params = [p for p in params]
        if not self.params:
            raise ValueError("FlatAdam: no parameters")
        if isinstance(self.params[0], dict):
            raise ValueError("FlatAdam: one parameter group only (the reference uses a single group)")
        dev = self.params[0].device
        if dev.type != "cuda":
            raise RuntimeError("FlatAdam: parameters must be CUDA tensors (there is no CPU path)")
        for p in self.params:
            if p.dtype != torch.float32 or p.device != dev:
                raise ValueError("FlatAdam: parameters must be fp32 tensors on one device")
        # torch.optim-style view of the hyper-parameters; 'lr' may be rewritten by the caller at any time
        self.param_groups = [{"params": self.params, "lr": float(lr), "betas": (float(betas[0]), float(betas[1])),
                              "eps": float(eps), "weight_decay": float(weight_decay), "amsgrad": False, "maximize": False,
                              "foreach": None, "capturable": True, "differentiable": False, "fused": None}]
        # every parameter starts on a 128-byte boundary (the field kernels read weights and biases with vector loads);
        # the padding stays zero under Adam
        self.offsets, n = [], 0
        for p in self.params:
            self.offsets.append(n)
            n += (p.numel() + 31) // 32 * 32
        self.n = n
        self.flat = torch.zeros(n, device=dev, dtype=torch.float32)
        with torch.no_grad():
            for p, off in zip(self.params, self.offsets):
                view = self.flat[off:off + p.numel()].view(p.shape)
                view.copy_(p.data)
                p.data = view                       # the parameter now lives in the flat buffer
        self.exp_avg = torch.zeros_like(self.flat)
        self.exp_avg_sq = torch.zeros_like(self.flat)
        self.flat_grad = torch.zeros_like(self.flat)
        self._grad_views = [self.flat_grad[off:off + p.numel()].view(p.shape) for p, off in zip(self.params, self.offsets)]
        self.step_t = torch.zeros(1, device=dev, dtype=torch.float32)
        self.lr_t = torch.full((1,), float(lr), device=dev, dtype=torch.float32)
        self._lr_on_device = float(lr)
        # steps a parameter sat out (no gradient), per 32-element block of the flat buffer; torch counts steps per parameter
        self.skipped = torch.zeros(n // 32, device=dev, dtype=torch.float32)
        self._block_slices = [slice(off // 32, (off + (p.numel() + 31) // 32 * 32) // 32) for p, off in zip(self.params, self.offsets)]
        self._any_skipped = False
        self._peer = None                   # set by enable_peer_exchange()

    # -- torch.optim surface ---------------------------------------------------------------------------------------
    @property
    def lr(self):
        return self.param_groups[0]["lr"]

    @lr.setter
    def lr(self, value):
        self.param_groups[0]["lr"] = float(value)

    def zero_grad(self, set_to_none=True):
        for p in self.params:
            if set_to_none:
                p.grad = None
            elif p.grad is not None:
                p.grad.zero_()

    def sync_lr(self):
        """Push param_groups[0]['lr'] to the device scalar the kernel reads (a host-to-device fill: call it OUTSIDE a CUDA
        graph capture; step() does it itself whenever it runs eagerly)."""
        lr = float(self.param_groups[0]["lr"])
        if lr != self._lr_on_device:
            self.lr_t.fill_(lr)
            self._lr_on_device = lr

    def state_dict(self):
        """torch.optim.Adam's format: loadable by torch.optim.Adam(params).load_state_dict and by FlatAdam."""
        step = float(self.step_t.item())
        skipped = self.skipped.cpu()
        state = {}
        for i, (p, off, bs) in enumerate(zip(self.params, self.offsets, self._block_slices)):
            t = step - float(skipped[bs.start])
            if t <= 0:
                continue                              # torch creates a parameter's state at its first step
            state[i] = {"step": torch.tensor(t, dtype=torch.float32),
                        "exp_avg": self.exp_avg[off:off + p.numel()].view(p.shape).clone(),
                        "exp_avg_sq": self.exp_avg_sq[off:off + p.numel()].view(p.shape).clone()}
        group = {k: v for k, v in self.param_groups[0].items() if k != "params"}
        group["params"] = list(range(len(self.params)))
        return {"state": state, "param_groups": [group]}

    def load_state_dict(self, sd):
        groups = sd["param_groups"]
        if len(groups) != 1 or len(groups[0]["params"]) != len(self.params):
            raise ValueError("FlatAdam.load_state_dict: expected one parameter group with %d parameters" % len(self.params))
        for k in ("lr", "betas", "eps", "weight_decay"):
            if k in groups[0]:
                self.param_groups[0][k] = tuple(groups[0][k]) if k == "betas" else float(groups[0][k])
        steps = [float(sd["state"][i]["step"]) if i in sd["state"] else 0.0 for i in range(len(self.params))]
        top = max(steps) if steps else 0.0
        self.step_t.fill_(top)
        self.exp_avg.zero_(); self.exp_avg_sq.zero_(); self.skipped.zero_()
        self._any_skipped = False
        for i, (p, off, bs) in enumerate(zip(self.params, self.offsets, self._block_slices)):
            if steps[i] != top:
                self.skipped[bs] = top - steps[i]
                self._any_skipped = True
            if i in sd["state"]:
                st = sd["state"][i]
                self.exp_avg[off:off + p.numel()].view(p.shape).copy_(st["exp_avg"])
                self.exp_avg_sq[off:off + p.numel()].view(p.shape).copy_(st["exp_avg_sq"])
        self._lr_on_device = None
        self.sync_lr()

    # -- the step ----------------------------------------------------------------------------------------------------
    def _runs(self):
        """maximal runs of consecutive parameters that have a gradient (torch's Adam skips the others entirely)"""
        runs, cur = [], None
        for i, p in enumerate(self.params):
            if p.grad is not None:
                if cur is None:
                    cur = [i, i]
                cur[1] = i
            elif cur is not None:
                runs.append(cur)
                cur = None
        if cur is not None:
            runs.append(cur)
        return runs

    def gather_grads(self):
        """p.grad of every parameter -> its slot of flat_grad (one multi-tensor copy); returns the runs"""
        have = [i for i, p in enumerate(self.params) if p.grad is not None]
        if have:
            torch._foreach_copy_([self._grad_views[i] for i in have], [self.params[i].grad for i in have])
        return self._runs()

    # -- multi-GPU: gradient exchange fused with the update (csrc/peer.cu) -----------------------------------------------
    def enable_peer_exchange(self):
        """COLLECTIVE (every rank of the default process group calls it, after construction and before the first step):
        moves ``flat_grad`` into a peer-mapped block (hn_peer_alloc), exchanges the IPC handles and opens the other ranks'
        blocks.  Afterwards ``step(..., peer_exchange=True)`` sums the ranks' gradients over NVLink peer memory and applies
        Adam in ONE launch (hn_peer_adam_flat) instead of ncclAllReduce(flat_grad) + hn_adam_flat.  Returns True when every
        rank succeeded; on False nothing changed and the caller keeps the NCCL path (dist.allreduce_flat + step)."""
        import torch.distributed as dist
        if self._peer is not None:
            return True
        if not (dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1):
            return False
        world, rank, dev = dist.get_world_size(), dist.get_rank(), self.flat.device
        nbytes = lib.hn_peer_block_bytes(self.n, world)
        ptr, handle = ctypes.c_void_p(), (ctypes.c_uint8 * 64)()
        ok = nbytes > 0 and lib.hn_peer_alloc(nbytes, ctypes.byref(ptr), handle) == 0
        handles = [None] * world
        dist.all_gather_object(handles, bytes(handle) if ok else None)
        ok = ok and all(h is not None for h in handles)
        blocks = [None] * world
        if ok:
            blocks[rank] = ptr.value
            for r in range(world):
                if r == rank:
                    continue
                q = ctypes.c_void_p()
                hb = (ctypes.c_uint8 * 64).from_buffer_copy(handles[r])
                if lib.hn_peer_open(hb, ctypes.byref(q)) != 0:
                    ok = False
                    break
                blocks[r] = q.value
        grad = None
        if ok:
            try:                                    # the block's first n floats as a tensor: p.grad is gathered straight into it
                grad = torch.as_tensor(_DeviceArray(ptr.value, self.n), device=dev)
                ok = grad.data_ptr() == ptr.value and grad.numel() == self.n and grad.dtype == torch.float32
            except Exception:                       # noqa: BLE001
                ok = False
        flag = torch.tensor([1 if ok else 0], device=dev if dist.get_backend() == "nccl" else "cpu")
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        if int(flag) != 1:
            for r, b in enumerate(blocks):
                if b is not None and r != rank:
                    lib.hn_peer_close(ctypes.c_void_p(b))
            if ptr.value:
                lib.hn_peer_free(ptr)
            return False
        grad.copy_(self.flat_grad)
        self.flat_grad = grad
        self._grad_views = [grad[off:off + p.numel()].view(p.shape) for p, off in zip(self.params, self.offsets)]
        self._peer = {"blocks": (ctypes.c_void_p * world)(*blocks), "rank": rank, "world": world, "own": ptr,
                      "epoch": torch.zeros(148, device=dev, dtype=torch.int32),
                      "err": torch.zeros(1, device=dev, dtype=torch.int32)}
        torch.cuda.synchronize(dev)
        dist.barrier()
        return True

    def close_peer_exchange(self):
        """COLLECTIVE: back to a private flat_grad (same values), unmap the other ranks' blocks, and -- after a barrier, so
        that no rank can still be reading it -- free the own one.  A captured graph that contains the peer step must not be
        replayed afterwards."""
        import torch.distributed as dist
        if self._peer is None:
            return
        pe, dev = self._peer, self.flat.device
        torch.cuda.synchronize(dev)
        grad = self.flat_grad.clone()
        self.flat_grad = grad
        self._grad_views = [grad[off:off + p.numel()].view(p.shape) for p, off in zip(self.params, self.offsets)]
        self._peer = None
        for r in range(pe["world"]):
            if r != pe["rank"] and pe["blocks"][r]:
                lib.hn_peer_close(ctypes.c_void_p(pe["blocks"][r]))
        dist.barrier()
        lib.hn_peer_free(pe["own"])

    def peer_error(self):
        """0, or 1 + the rank a barrier of hn_peer_adam_flat gave up waiting for (device -> host read: not inside a capture)"""
        return 0 if self._peer is None else int(self._peer["err"].item())

    def _peer_launch(self, mode, out, grad_scale):
        g, pe = self.param_groups[0], self._peer
        stream = ctypes.c_void_p(torch.cuda.current_stream(self.flat.device).cuda_stream)
        v = lambda t: ctypes.c_void_p(t.data_ptr())
        check(lib.hn_peer_adam_flat(v(out), v(self.exp_avg), v(self.exp_avg_sq), self.n, pe["blocks"], pe["rank"], pe["world"],
                                    v(pe["epoch"]), v(pe["err"]), mode, v(self.step_t), v(self.lr_t), float(g["lr"]),
                                    float(g["betas"][0]), float(g["betas"][1]), float(g["eps"]), float(g["weight_decay"]),
                                    float(grad_scale), stream), "hn_peer_adam_flat")

    def peer_allreduce(self, out=None, grad_scale=1.0):
        """COLLECTIVE: out [n] = grad_scale * (sum over ranks of flat_grad), by the same exchange the fused step uses (mode 1
        of hn_peer_adam_flat); flat_grad itself is left untouched.  Used to check the exchange against ncclAllReduce."""
        if self._peer is None:
            raise RuntimeError("FlatAdam.peer_allreduce: call enable_peer_exchange() first")
        out = torch.empty_like(self.flat) if out is None else out
        self._peer_launch(1, out, grad_scale)
        return out

    def step(self, runs=None, grad_scale=1.0, peer_exchange=False):
        """One Adam update.  ``runs``: the result of an earlier ``gather_grads`` (e.g. before an all-reduce of
        ``flat_grad``); gathered here when omitted.  ``peer_exchange=True`` (after enable_peer_exchange(); COLLECTIVE): the
        gradients are summed over the ranks inside the update kernel -- the caller does NOT all-reduce flat_grad."""
        if runs is None:
            runs = self.gather_grads()
        dev = self.flat.device
        if not torch.cuda.is_current_stream_capturing():
            self.sync_lr()
        if peer_exchange:
            # every rank launches the same kernel over the whole buffer: a parameter without a gradient on this rank may
            # have one elsewhere, so its slot of flat_grad is sent as zeros
            covered = set()
            for a, b in runs:
                covered.update(range(a, b + 1))
            missing = [self._grad_views[i] for i in range(len(self.params)) if i not in covered]
            if missing:
                torch._foreach_zero_(missing)
            if self._peer is None:
                raise RuntimeError("FlatAdam.step(peer_exchange=True): call enable_peer_exchange() first")
            if self._any_skipped:
                raise RuntimeError("FlatAdam.step(peer_exchange=True): per-parameter step counts are not supported on this path")
            self.step_t += 1.0
            self._peer_launch(0, self.flat, grad_scale)
            torch.autograd.graph.increment_version(self.params)
            return
        self.step_t += 1.0
        # parameters outside every run sit this step out: their own step count stays behind (torch: per-parameter `step`)
        covered = set()
        for a, b in runs:
            covered.update(range(a, b + 1))
        missing = [i for i in range(len(self.params)) if i not in covered]
        if missing:
            for i in missing:
                self.skipped[self._block_slices[i]] += 1.0
            self._any_skipped = True
        g = self.param_groups[0]
        stream = ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
        f = lambda t, off: ctypes.c_void_p(t.data_ptr() + 4 * off)
        for a, b in runs:
            lo, hi = self.offsets[a], self.offsets[b] + self.params[b].numel()
            sk = f(self.skipped, lo // 32) if self._any_skipped else None
            check(lib.hn_adam_flat(f(self.flat, lo), f(self.flat_grad, lo), f(self.exp_avg, lo), f(self.exp_avg_sq, lo), hi - lo,
                                   ctypes.c_void_p(self.step_t.data_ptr()), ctypes.c_void_p(self.lr_t.data_ptr()), sk,
                                   float(g["lr"]), float(g["betas"][0]), float(g["betas"][1]), float(g["eps"]),
                                   float(g["weight_decay"]), float(grad_scale), stream), "hn_adam_flat")
        # the kernel wrote through raw pointers: tell autograd (and the packed-weight caches keyed on _version)
        torch.autograd.graph.increment_version(self.params)
