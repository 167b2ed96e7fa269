"""Stand-alone compositor workload (bench.py's micro-benchmark: 2^18 rays x 128 samples) for ncu / timing."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import torch
import bench
import honerf_b200 as H
res = bench.compositor_roofline(H, torch.device("cuda:0"))
print({k: (round(v["ms"], 4), round(v["frac"], 3)) for k, v in res.items() if isinstance(v, dict)})
