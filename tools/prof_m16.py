"""MMA-issuer wait breakdown of the HN_TC_MIXED16 sweep kernels (cycle counters written to the host-mapped debug buffer)."""
import ctypes, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import torch
import honerf_b200 as H
from honerf_b200 import _lib
from gpu_util import obj_modules
dbg = torch.zeros(4 * 4 * 148 * 8, dtype=torch.int32).pin_memory()
_lib.lib.hn_chain16_set_debug(ctypes.c_void_p(dbg.data_ptr()))
sdf, col, dev, _, _ = obj_modules()
n = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
x = (0.45 * torch.randn(n, 3)).cuda().requires_grad_(True)
p = H.ops._PRECISIONS["tc_mixed16"]
for _ in range(2):
    s, f, nn = H.ops.sdf_obj(sdf.packed(), x, 1.0, precision=p)
    (s.sum() + f.sum() + (nn * nn).sum()).backward()
torch.cuda.synchronize()
d = dbg.view(4, 4, 148, 8).numpy().astype("uint32")
for k, name in enumerate(["trunk16", "nsweep16", "bwd16", "dw16"]):
    for inst in range(4):
        tot = d[k, inst, :, 6].astype("float64") * 16
        if tot.max() == 0:
            continue
        ta, tw = d[k, inst, :, 4].astype("float64") * 16, d[k, inst, :, 5].astype("float64") * 16
        print("%s inst %d: total %.0f kcyc; MMA issuer waits: A operand %.1f %%, weights %.1f %%, issuing %.1f %%" % (
            name, inst, tot.mean() / 1e3, 100 * ta.mean() / tot.mean(), 100 * tw.mean() / tot.mean(),
            100 * (1 - (ta.mean() + tw.mean()) / tot.mean())))
