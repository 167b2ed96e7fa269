"""MMA-issuer wait breakdown of the hand-field sweep kernels (cycle counters in the host-mapped debug buffer)."""
import ctypes, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import torch
import honerf_b200 as H
import synth
from honerf_b200 import _lib
from gpu_util import hand_modules
dbg = torch.zeros(4 * 4 * 148 * 8, dtype=torch.int32).pin_memory()
_lib.lib.hn_chain16_set_debug(ctypes.c_void_p(dbg.data_ptr()))
sdf, col, dev, _, _ = hand_modules(requires_grad=False)
bt0, T, J = synth.hand_pose()
n = int(sys.argv[1]) if len(sys.argv) > 1 else 98304
g = torch.Generator().manual_seed(1)
x = (J[torch.randint(0, 21, (n,), generator=g)] + 0.03 * torch.randn(n, 3, generator=g)).cuda().requires_grad_(True)
bt = bt0.cuda().requires_grad_(True)
T = T.cuda()
for _ in range(2):
    s, f, nn, xyz = sdf.fused(x, bt, T)
    (s.sum() + f.sum() + (nn * nn).sum()).backward()
torch.cuda.synchronize()
d = dbg.view(4, 4, 148, 8).numpy().astype("uint32")
for k, name in enumerate(["trunk16", "nsweep16", "bwd16", "dw16"]):
    for inst in range(4):
        tot = d[k, inst, :, 6].astype("float64") * 16
        if tot.max() == 0:
            continue
        ta, tw = d[k, inst, :, 4].astype("float64") * 16, d[k, inst, :, 5].astype("float64") * 16
        print("%s inst %d: total %.0f kcyc (max %.0f); MMA issuer waits: A operand %.1f %%, weights %.1f %%, issuing %.1f %%" % (
            name, inst, tot.mean() / 1e3, tot.max() / 1e3, 100 * ta.mean() / tot.mean(), 100 * tw.mean() / tot.mean(),
            100 * (1 - (ta.mean() + tw.mean()) / tot.mean())))
