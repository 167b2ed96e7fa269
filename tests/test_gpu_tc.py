"""tcgen05 bring-up: the self-test GEMM (fp16 / bf16 operands, fp32 accumulation in TMEM) against
torch.matmul of the same rounded operands.  Tolerance: 2e-3 of the row scale (fp32 accumulation order
differs)."""
import ctypes

import pytest
import torch

pytestmark = pytest.mark.gpu


_ST = None


def _selftest():
    """libhonerf_b200_selftest.so (csrc/selftest/): the bring-up GEMMs are not part of the product library."""
    global _ST
    if _ST is None:
        from honerf_b200 import _lib
        _ST = _lib.load_selftest()
    return _ST


def _check(status, what):
    if status != 0:
        raise RuntimeError("%s failed (%d): %s" % (what, status, (_selftest().hn_last_error() or b"?").decode()))


def _run(M, N, K, bf16, ts=False):
    g = torch.Generator().manual_seed(M * 7 + N * 3 + K)
    dt = torch.bfloat16 if bf16 else torch.float16
    A = torch.randn(M, K, generator=g).to(dt).cuda()
    B = torch.randn(N, K, generator=g).to(dt).cuda()
    C = torch.full((M, N), float("nan"), device="cuda")
    fn = _selftest().hn_tc_gemm_ts_test if ts else _selftest().hn_tc_gemm_test
    _check(fn(ctypes.c_void_p(A.data_ptr()), ctypes.c_void_p(B.data_ptr()), M, N, K, int(bf16),
                  ctypes.c_void_p(C.data_ptr()), ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)), "hn_tc_gemm_test")
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


@pytest.mark.parametrize("M,N,K,bf16", [(128, 256, 64, False), (300, 128, 256, True), (129, 256, 256, True)])
def test_tc_gemm_a_operand_in_tensor_memory(M, N, K, bf16):
    """The `ts` MMA form: A written to TMEM with tcgen05.st (two 16-bit values per 32-bit column, lane = row)."""
    err, scale = _run(M, N, K, bf16, ts=True)
    assert err < 2e-3 * scale, (err, scale)


def _gemm(layout, passes, M, N, K, lda_pad=0, with_bias=False, seed=0):
    from honerf_b200 import _lib
    g = torch.Generator().manual_seed(seed + M + 3 * N + 7 * K)
    P = lambda t: ctypes.c_void_p(t.data_ptr()) if t is not None else None
    st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    r4 = lambda x: (x + 3) // 4 * 4
    if layout == 0:      # C = A[M,K] @ B[N,K]^T
        lda, ldb = r4(K) + lda_pad, r4(K)
        A = torch.randn(M, lda, generator=g).cuda(); B = torch.randn(N, ldb, generator=g).cuda()
        ref = A[:, :K].double() @ B[:, :K].double().T
    elif layout == 1:    # C = A[M,K] @ B[K,N]
        lda, ldb = r4(K) + lda_pad, r4(N)
        A = torch.randn(M, lda, generator=g).cuda(); B = torch.randn(K, ldb, generator=g).cuda()
        ref = A[:, :K].double() @ B[:, :N].double()
    else:                # C += A[K,M]^T @ B[K,N]
        lda, ldb = r4(M) + lda_pad, r4(N)
        A = torch.randn(K, lda, generator=g).cuda(); B = torch.randn(K, ldb, generator=g).cuda()
        ref = A[:, :M].double().T @ B[:, :N].double()
    bias = torch.randn(N, generator=g).cuda() if with_bias else None
    if bias is not None:
        ref = ref + bias.double()
    ldc = r4(N)
    C = torch.zeros(M, ldc, device="cuda")
    _check(_selftest().hn_gemm_test(layout, passes, M, N, K, P(A), lda, P(B), ldb, P(bias), P(C), ldc, st),
               "hn_gemm_test")
    torch.cuda.synchronize()
    err = (C[:, :N].double() - ref).abs().max().item()
    return err, ref.abs().max().item()


@pytest.mark.parametrize("layout,M,N,K", [(0, 128, 256, 256), (0, 1000, 193, 256), (0, 333, 256, 63), (0, 257, 3, 256),
                                           (0, 640, 256, 373), (1, 512, 256, 257), (1, 300, 63, 256),
                                           (1, 129, 373, 256), (1, 200, 256, 193),
                                           (2, 256, 256, 4099), (2, 193, 256, 1000), (2, 257, 256, 777),
                                           (2, 256, 373, 640), (2, 3, 256, 500)])
@pytest.mark.parametrize("passes", [0, 1, 3])
def test_production_gemms(layout, M, N, K, passes):
    """fp32 SIMT <= 2e-6, single-pass TF32 <= 2e-3, split TF32 <= 2e-5 of the result scale (fp64 reference)."""
    err, scale = _gemm(layout, passes, M, N, K, with_bias=(layout == 0))
    tol = {0: 2e-6, 1: 2e-3, 3: 2e-5}[passes]
    assert err < tol * scale, (err, scale)
