"""Adam over ONE flat parameter buffer (``hn_adam_flat``): the optimiser step of the train loop in one launch.

The reference builds a single ``torch.optim.Adam`` over every network's parameters (exp_runner.py:83-90); with the
weight-normalised MLPs that is ~90 small tensors per step.  ``FlatAdam`` re-homes the parameters as views of one
contiguous buffer (their values, shapes and ``state_dict`` are unchanged), gathers the step's gradients into a
second flat buffer and updates everything with one kernel.  Same update rule as ``torch.optim.Adam`` (bias
correction, ``eps`` outside the square root, L2 ``weight_decay``); the step count lives on the device so the step
can be captured in a CUDA graph.  ``flat_grad`` is also the payload of the multi-GPU gradient all-reduce.
"""
import ctypes

import torch

from ._lib import check, lib


class FlatAdam:
    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0):
        self.params = [p for p in params]
        if not self.params:
            raise ValueError("FlatAdam: no parameters")
        dev = self.params[0].device
        if dev.type != "cuda":
            raise RuntimeError("FlatAdam: parameters must be CUDA tensors (there is no CPU path)")
        for p in self.params:
            if p.dtype != torch.float32 or p.device != dev:
                raise ValueError("FlatAdam: parameters must be fp32 tensors on one device")
        self.lr, self.betas, self.eps, self.weight_decay = float(lr), (float(betas[0]), float(betas[1])), float(eps), float(weight_decay)
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

    def zero_grad(self, set_to_none=True):
        for p in self.params:
            if set_to_none:
                p.grad = None
            elif p.grad is not None:
                p.grad.zero_()

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

    def step(self, runs=None, grad_scale=1.0):
        """One Adam update.  ``runs``: the result of an earlier ``gather_grads`` (e.g. before an all-reduce of
        ``flat_grad``); gathered here when omitted."""
        if runs is None:
            runs = self.gather_grads()
        self.step_t += 1.0
        stream = ctypes.c_void_p(torch.cuda.current_stream(self.flat.device).cuda_stream)
        f = lambda t, off: ctypes.c_void_p(t.data_ptr() + 4 * off)
        for a, b in runs:
            lo, hi = self.offsets[a], self.offsets[b] + self.params[b].numel()
            check(lib.hn_adam_flat(f(self.flat, lo), f(self.flat_grad, lo), f(self.exp_avg, lo), f(self.exp_avg_sq, lo), hi - lo,
                                   ctypes.c_void_p(self.step_t.data_ptr()), self.lr, self.betas[0], self.betas[1], self.eps,
                                   self.weight_decay, float(grad_scale), stream), "hn_adam_flat")
        # the kernel wrote through raw pointers: tell autograd (and the packed-weight caches keyed on _version)
        torch.autograd.graph.increment_version(self.params)
