import sys, ctypes
sys.path[:0] = ['/root/repo']
import torch
from honerf_b200 import _lib
P = lambda t: ctypes.c_void_p(t.data_ptr())
st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
def run(A, B, M, N, K):
    C = torch.zeros(M, N, device='cuda')
    _lib.check(_lib.lib.hn_gemm_test(2, 1, M, N, K, P(A), A.shape[1], P(B), B.shape[1], None, P(C), N, st), 'x')
    torch.cuda.synchronize()
    return C.cpu()
M, N, K = 128, 64, 32
A = torch.ones(K, M, device='cuda'); B = torch.ones(K, N, device='cuda')
C = run(A, B, M, N, K); print('ones: min/max', C.min().item(), C.max().item())
A = torch.zeros(K, M, device='cuda'); A[0] = torch.arange(M).float()
C = run(A, B, M, N, K); print('A[0,m]=m: C[:,0] first 10', C[:10, 0].tolist(), 'C[40:44,0]', C[40:44, 0].tolist(), 'C[5,:6]', C[5, :6].tolist())
A = torch.ones(K, M, device='cuda'); B = torch.zeros(K, N, device='cuda'); B[0] = torch.arange(N).float()
C = run(A, B, M, N, K); print('B[0,n]=n: C[0,:10]', C[0, :10].tolist(), 'C[0,30:36]', C[0, 30:36].tolist())
for k0 in (0, 1, 7, 8, 9, 31):
    A = torch.arange(K, device='cuda').float()[:, None].expand(K, M).contiguous() + 1
    B = torch.zeros(K, N, device='cuda'); B[k0] = 1
    C = run(A, B, M, N, K); print('k0', k0, 'C[0,0]', C[0, 0].item(), 'uniq', C.unique().tolist()[:6])
