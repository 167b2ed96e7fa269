#!/bin/bash
# round 2, call B: first run of the HN_TC_MIXED16 kernels (wrapped in timeouts: a hang must not take the box down)
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
echo "== dw16 (descriptor order as designed)"; timeout 120 python -m pytest tests/test_gpu_mixed16.py -q -s -k dw16 2>&1 | tail -15 | tee gpurun_out/r2b_dw16.log
echo "== dw16 with LBO/SBO swapped"; HONERF_DW16_SWAP=1 timeout 120 python - <<'PY' 2>&1 | tail -5 | tee gpurun_out/r2b_dw16_swap.log
import sys; sys.path[:0]=['.','oracle','tests']
import torch, ctypes
from honerf_b200 import _lib
_lib.lib.hn_dw16_set_debug(1)
n,out,nin=1000,256,256
g=torch.Generator().manual_seed(1)
P=torch.randn(n,out,generator=g).cuda(); Q=torch.randn(n,nin,generator=g).cuda()
ref=P.bfloat16().double().T@Q.bfloat16().double()
C=torch.zeros(out,nin,device='cuda'); db=torch.zeros(out,device='cuda'); part=torch.empty(16*65536,device='cuda')
tiles=torch.empty(4*1024*512,device='cuda',dtype=torch.uint8)
p=lambda t: ctypes.c_void_p(t.data_ptr())
_lib.check(_lib.lib.hn_dw16_test(p(P),out,p(Q),nin,None,None,n,p(C),nin,p(db),p(tiles),tiles.numel(),p(part),part.numel(),ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)),"t")
torch.cuda.synchronize()
print("swapped: rel err", float((C.double().cpu()-ref.cpu()).abs().max()/ref.abs().max()))
PY
echo "== mixed16 operator"; timeout 300 python -m pytest tests/test_gpu_mixed16.py -q -s -k "not dw16" 2>&1 | tail -40 | tee gpurun_out/r2b_m16.log
echo "== bench mixed16"; timeout 300 python bench.py --precision tc_mixed16 --no-cpu-baseline --large-rays 0 --fit-rays 0 --grid-res 0 > gpurun_out/r2b_bench_m16.json 2> gpurun_out/r2b_bench_m16.err; tail -c 1500 gpurun_out/r2b_bench_m16.json; tail -3 gpurun_out/r2b_bench_m16.err
