"""Short workload for the ncu launch list of the hand field: SDFNetwork (HALO) value + normal forward and the second-order
backward to the points / bone transforms (weights frozen, as in pose fitting), plus one sdf-only call, on 98 304 points
(512 rays x 192 samples)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import torch  # noqa: E402

import honerf_b200 as H  # noqa: E402
import synth  # noqa: E402
from gpu_util import hand_modules  # noqa: E402

grad_w = os.environ.get("PROF_HAND_WEIGHTS", "0") == "1"
sdf, col, dev, _, _ = hand_modules(requires_grad=grad_w)
for q in col.parameters():
    q.requires_grad_(False)
bt0, T, J = synth.hand_pose()
n = int(os.environ.get("PROF_HAND_POINTS", 98304))
g = torch.Generator().manual_seed(1)
x = (J[torch.randint(0, 21, (n,), generator=g)] + 0.03 * torch.randn(n, 3, generator=g)).cuda().requires_grad_(True)
d = torch.nn.functional.normalize(torch.randn(n, 3, generator=g), dim=-1).cuda()
bt = bt0.cuda().requires_grad_(True)
T = T.cuda()
e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
for it in range(3):
    e0.record()
    s, f, nn, xyz = sdf.fused(x, bt, T) if hasattr(sdf, "fused") else None
    rgb = col(d, xyz, f, None, nn) if os.environ.get("PROF_HAND_COLOR", "1") == "1" else None
    e1.record()
    loss = s.sum() + (nn * nn).sum() + (rgb.sum() if rgb is not None else 0.0)
    loss.backward()
    e2.record()
    torch.cuda.synchronize()
    print("iter %d: fwd %.3f ms, bwd %.3f ms" % (it, e0.elapsed_time(e1), e1.elapsed_time(e2)))
with torch.no_grad():
    e0.record()
    sdf.sdf(x, bt, T)
    e1.record()
torch.cuda.synchronize()
print("sdf only %.3f ms" % e0.elapsed_time(e1))
