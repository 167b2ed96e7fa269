import sys, ctypes
sys.path[:0] = ['/root/repo']
import torch
from honerf_b200 import _lib
P = lambda t: ctypes.c_void_p(t.data_ptr()) if t is not None else None
st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
def timeit(layout, passes, M, N, K, reps=20):
    if layout == 0: A = torch.randn(M, K, device='cuda'); B = torch.randn(N, K, device='cuda'); lda, ldb = K, K
    elif layout == 1: A = torch.randn(M, K, device='cuda'); B = torch.randn(K, N, device='cuda'); lda, ldb = K, N
    else: A = torch.randn(K, M, device='cuda'); B = torch.randn(K, N, device='cuda'); lda, ldb = M, N
    C = torch.zeros(M, N, device='cuda'); bias = torch.zeros(N, device='cuda') if layout == 0 else None
    f = lambda: _lib.check(_lib.lib.hn_gemm_test(layout, passes, M, N, K, P(A), lda, P(B), ldb, P(bias), P(C), N, st), 'g')
    for _ in range(3): f()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): f()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    print('layout %d passes %d M=%d N=%d K=%d: %.1f us  %.1f TFLOP/s' % (layout, passes, M, N, K, ms * 1e3, 2.0 * M * N * K / ms / 1e9))
if len(sys.argv) > 1:
    timeit(int(sys.argv[1]), int(sys.argv[2]), 65536, 256, 256, reps=2)
    sys.exit(0)
for layout, passes in [(0, 0), (0, 1), (0, 3), (1, 0), (1, 1)]:
    timeit(layout, passes, 65536, 256, 256)
for passes in (0, 1):
    timeit(2, passes, 256, 256, 65536)
timeit(0, 1, 65536, 256, 64); timeit(0, 1, 8192, 256, 256); timeit(0, 1, 1 << 20, 256, 256, reps=5)
