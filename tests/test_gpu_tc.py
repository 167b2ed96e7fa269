"""tcgen05 bring-up: the self-test GEMM (fp16 / bf16 operands, fp32 accumulation in TMEM) against
torch.matmul of the same rounded operands.  Tolerance: 2e-3 of the row scale (fp32 accumulation order
differs)."""
import ctypes

import pytest
import torch

pytestmark = pytest.mark.gpu


def _run(M, N, K, bf16):
    from honerf_b200 import _lib
    g = torch.Generator().manual_seed(M * 7 + N * 3 + K)
    dt = torch.bfloat16 if bf16 else torch.float16
    A = torch.randn(M, K, generator=g).to(dt).cuda()
    B = torch.randn(N, K, generator=g).to(dt).cuda()
    C = torch.full((M, N), float("nan"), device="cuda")
    _lib.check(_lib.lib.hn_tc_gemm_test(ctypes.c_void_p(A.data_ptr()), ctypes.c_void_p(B.data_ptr()), M, N, K,
                                        int(bf16), ctypes.c_void_p(C.data_ptr()),
                                        ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)), "hn_tc_gemm_test")
    torch.cuda.synchronize()
    ref = A.float() @ B.float().T
    err = (C - ref).abs().max().item()
    scale = ref.abs().max().item()
    return err, scale


@pytest.mark.parametrize("M,N,K,bf16", [(128, 256, 64, False), (128, 64, 128, False), (300, 256, 256, False),
                                         (1000, 208, 256, False), (128, 16, 64, False), (257, 256, 256, True)])
def test_tc_gemm_selftest(M, N, K, bf16):
    err, scale = _run(M, N, K, bf16)
    assert err < 2e-3 * scale, (err, scale)
