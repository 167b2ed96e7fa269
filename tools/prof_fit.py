"""Short workload for an ncu launch list: a few two-field fitting steps (bench.fitting_extra)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import torch  # noqa: E402

import bench  # noqa: E402
import honerf_b200 as H  # noqa: E402

print(bench.fitting_extra(H, torch.device("cuda", 0), int(sys.argv[1]) if len(sys.argv) > 1 else 512, "tc_bf16x3"))
