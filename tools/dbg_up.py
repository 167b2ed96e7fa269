import sys; sys.path[:0]=['/root/repo','/root/repo/oracle','/root/repo/tests']
import torch
from golden_util import load_golden
from honerf_b200 import ops
g = load_golden("sampling")
z, s = g["z0"].cuda(), g["sdf0"].cuda()
for i in range(4):
    new_z = ops.up_sample(z, s, 16, 64 * 2 ** i).cpu()
    ref = g["new_z%d" % i]
    d = (new_z-ref).abs()
    print(i, 'max', d.max().item(), 'exact frac', (new_z==ref).float().mean().item(), 'n>2e-6', (d>2e-6).sum().item(), 'argmax', divmod(d.argmax().item(), 16))
    b = d.argmax().item()//16
    print('  row', b, new_z[b].tolist()[:16]); print('  ref', ref[b].tolist()[:16])
    z = g["z%d" % (i + 1)].cuda()
    if i < 3: s = g["sdf%d" % (i + 1)].cuda()
