"""Colour-net gradients: HN_TC_MIXED16 (16-bit hi/lo stash tiles + dw16_kernel) against HN_TC_BF16X3 (fp32 stash + dw_kernel) on
the same inputs, per tensor."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import torch
import honerf_b200 as H
from gpu_util import obj_modules
n = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
sdf, col, dev, _, _ = obj_modules(requires_grad=True)
g = torch.Generator().manual_seed(0)
x = (0.45 * torch.randn(n, 3, generator=g)).cuda()
d = torch.nn.functional.normalize(torch.randn(n, 3, generator=g), dim=-1).cuda()
feat = torch.randn(n, 256, generator=g).cuda() * 0.3
nrm = torch.nn.functional.normalize(torch.randn(n, 3, generator=g), dim=-1).cuda()
go = (torch.randn(n, 3, generator=g) * 1e-3).cuda()
res = {}
for name in ("tc_bf16x3", "tc_mixed16"):
    for q in col.parameters():
        q.grad = None
    xs = [t.clone().requires_grad_(True) for t in (x, d, feat, nrm)]
    rgb = H.ops.color_obj(col.packed(), xs[0], xs[1], xs[2], xs[3], precision=H.ops._PRECISIONS[name])
    (rgb * go).sum().backward()
    res[name] = ({k: q.grad.clone() for k, q in col.named_parameters()}, [t.grad.clone() for t in xs], rgb.detach())
a, b = res["tc_mixed16"], res["tc_bf16x3"]
print("rgb", float((a[2] - b[2]).abs().max()))
for i, nm in enumerate(["d_pts", "d_dirs", "d_feat", "d_normal"]):
    print(nm, float((a[1][i] - b[1][i]).norm() / b[1][i].norm()))
for k in a[0]:
    print(k, "%.2e" % float((a[0][k] - b[0][k]).norm() / b[0][k].norm()))
