"""Micro-benchmark of the fused tile-chain kernels (run on a B200): device time per call with CUDA events."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import torch  # noqa: E402

import honerf_b200 as H  # noqa: E402
from gpu_util import obj_modules  # noqa: E402

F_O = 1049088


def timeit(fn, iters=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / iters


def main():
    sdf, col, dev, _, _ = obj_modules(requires_grad=False)
    for n in (65536, 1 << 20):
        x = (0.45 * torch.randn(n, 3)).cuda()
        for prec in ("tc_bf16x3", "tc_tf32x3", "simt_fp32"):
            p = H.ops._PRECISIONS[prec]
            ms = timeit(lambda: H.ops.sdf_obj_sdf_only(sdf.packed(), x, 1.0, precision=p))
            print("sdf_only n=%d %-10s %.3f ms  %.1f algorithmic TFLOP/s" % (n, prec, ms, n * (F_O - 2 * 257 * 256 + 512) / ms / 1e9))


def prof():
    """cycle split of the chain kernel's MMA warp / epilogue (1M points)"""
    import ctypes
    sdf, col, dev, _, _ = obj_modules(requires_grad=False)
    n = 1 << 20
    x = (0.45 * torch.randn(n, 3)).cuda()
    buf = torch.zeros(148 * 4, dtype=torch.int64, device="cuda")
    p = H.ops._PRECISIONS["tc_bf16x3"]
    H.ops.sdf_obj_sdf_only(sdf.packed(), x, 1.0, precision=p)
    H._lib.lib.hn_chain_set_prof(ctypes.c_void_p(buf.data_ptr()))
    H.ops.sdf_obj_sdf_only(sdf.packed(), x, 1.0, precision=p)
    torch.cuda.synchronize()
    H._lib.lib.hn_chain_set_prof(None)
    b = buf.reshape(148, 4).double().mean(0).tolist()
    print("MMA warp: wait A %.0f  wait weights %.0f  total %.0f cycles;  epilogue waits for acc %.0f  (per CTA, 55 tiles x 8 layers)"
          % tuple(b))
    steps = (n / 128 / 148) * 8
    print("per layer-tile: total %.0f  waitA %.0f  waitW %.0f  mma-issue+rest %.0f ; epi wait acc %.0f"
          % (b[2] / steps, b[0] / steps, b[1] / steps, (b[2] - b[0] - b[1]) / steps, b[3] / steps))


def fwd_bench():
    sdf, col, dev, _, _ = obj_modules(requires_grad=False)
    n = 65536
    x = (0.45 * torch.randn(n, 3)).cuda()
    for prec in ("tc_bf16x3", "tc_tf32x3"):
        p = H.ops._PRECISIONS[prec]
        ms = timeit(lambda: H.ops.sdf_obj(sdf.packed(), x, 1.0, precision=p))
        print("sdf fwd (value+feat+normal) n=%d %-10s %.3f ms  %.1f algorithmic TFLOP/s" % (n, prec, ms, n * 2 * F_O / ms / 1e9))


def fwdbwd_bench():
    sdf, col, dev, _, _ = obj_modules(requires_grad=True)
    n = 65536
    x = (0.45 * torch.randn(n, 3)).cuda().requires_grad_(True)
    gs, gf, gn = torch.randn(n, 1).cuda(), torch.randn(n, 256).cuda(), torch.randn(n, 3).cuda()
    for prec in ("tc_bf16x3", "tc_tf32x3"):
        p = H.ops._PRECISIONS[prec]

        def step():
            s, f, nn = H.ops.sdf_obj(sdf.packed(), x, 1.0, precision=p)
            torch.autograd.backward([s, f, nn], [gs, gf, gn])
        ms = timeit(step, iters=5)
        print("sdf fwd+bwd (2nd order, dW) n=%d %-10s %.3f ms  %.1f algorithmic TFLOP/s" % (n, prec, ms, n * 6 * F_O / ms / 1e9))


if __name__ == "__main__" and len(sys.argv) == 1:
    prof()
    fwd_bench()
    fwdbwd_bench()
    main()


def prof_fwd_bwd():
    import ctypes
    sdf, col, dev, _, _ = obj_modules(requires_grad=True)
    n = 65536
    x = (0.45 * torch.randn(n, 3)).cuda().requires_grad_(True)
    gs, gf, gn = torch.randn(n, 1).cuda(), torch.randn(n, 256).cuda(), torch.randn(n, 3).cuda()
    p = H.ops._PRECISIONS["tc_bf16x3"]
    buf = torch.zeros(148 * 4 + 148 * 32, dtype=torch.int64, device="cuda")
    for which in ("fwd", "bwd"):
        s, f, nn = H.ops.sdf_obj(sdf.packed(), x, 1.0, precision=p)
        torch.autograd.backward([s, f, nn], [gs, gf, gn])
        torch.cuda.synchronize()
        buf.zero_()
        if which == "fwd":
            H._lib.lib.hn_chain_set_prof(ctypes.c_void_p(buf.data_ptr()))
            s, f, nn = H.ops.sdf_obj(sdf.packed(), x, 1.0, precision=p)
            torch.cuda.synchronize()
            H._lib.lib.hn_chain_set_prof(None)
            torch.autograd.backward([s, f, nn], [gs, gf, gn])
        else:
            s, f, nn = H.ops.sdf_obj(sdf.packed(), x, 1.0, precision=p)
            torch.cuda.synchronize()
            H._lib.lib.hn_chain_set_prof(ctypes.c_void_p(buf.data_ptr()))
            torch.autograd.backward([s, f, nn], [gs, gf, gn])
            torch.cuda.synchronize()
            H._lib.lib.hn_chain_set_prof(None)
        b = buf[:148 * 4].reshape(148, 4).double().cpu()
        per_step = buf[148 * 4:].reshape(148, 32).double().cpu()
        tl = torch.tensor([4.0 if i < 68 else 3.0 for i in range(148)], dtype=torch.float64)
        print(which, "wait for the epilogue BEFORE step s (cycles per tile):", [int(x) for x in (per_step / tl[:, None]).mean(0)[:18].tolist()])
        first = per_step[:, 31]
        later = (per_step[:, 0] - first) / (tl - 1)
        print(which, "wait before step 0: first tile of a CTA %.0f (min %.0f max %.0f), later tiles %.0f cycles" % (first.mean(), first.min(), first.max(), later.mean()))
        tiles = torch.tensor([4.0 if i < 68 else 3.0 for i in range(148)], dtype=torch.float64)
        steps = tiles * 17
        print("%s chain kernel, per layer-tile cycles (mean over CTAs): total %.0f  MMA-warp waits for epilogue %.0f  for weights %.0f  issue+rest %.0f; max CTA total %.0f cycles"
              % (which, (b[:, 2] / steps).mean(), (b[:, 0] / steps).mean(), (b[:, 1] / steps).mean(),
                 ((b[:, 2] - b[:, 0] - b[:, 1]) / steps).mean(), b[:, 2].max()))


if __name__ == "__main__" and len(sys.argv) > 1 and sys.argv[1] == "prof2":
    prof_fwd_bwd()


if __name__ == "__main__" and len(sys.argv) > 1 and sys.argv[1] == "stagger":
    for fc, bc in ((0, 0), (5000, 8000), (3000, 5000), (8000, 12000), (12000, 18000)):
        H._lib.lib.hn_chain_set_stagger(fc, bc)
        print("stagger", fc, bc)
        fwd_bench()
        fwdbwd_bench()
