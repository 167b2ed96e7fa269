"""Short workload for `ncu --set full`: per iteration one SDF-only query (57 344 points), one object-SDF forward + second-order
backward (65 536 points) and the colour forward / backward through the chain kernels: 8 chain-kernel launches."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import torch  # noqa: E402

import honerf_b200 as H  # noqa: E402
from gpu_util import obj_modules  # noqa: E402

sdf, col, dev, _, _ = obj_modules(requires_grad=True)
n = 65536
x = (0.45 * torch.randn(n, 3)).cuda().requires_grad_(True)
d = torch.nn.functional.normalize(torch.randn(n, 3), dim=-1).cuda()
p = H.ops._PRECISIONS[os.environ.get("PROF_PRECISION", "tc_bf16x3")]
xs = x.detach()[:57344]          # the largest SDF-only query of the 512-ray step (112 samples per ray)
for _ in range(3):
    with torch.no_grad():
        sdf.sdf(xs)
    s, f, nn = H.ops.sdf_obj(sdf.packed(), x, 1.0, precision=p)
    rgb = H.ops.color_obj(col.packed(), x, d, f, nn, precision=p)
    (rgb.sum() + s.sum() + (nn * nn).sum()).backward()
torch.cuda.synchronize()
