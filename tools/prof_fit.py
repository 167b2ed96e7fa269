"""Launch-list workload: the two-field pose-fitting step of bench.py (512 rays x 192 samples x 2 fields, frozen nets)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import torch  # noqa: E402

import bench  # noqa: E402
import honerf_b200 as H  # noqa: E402

print(bench.fitting_extra(H, torch.device("cuda:0"), int(os.environ.get("PROF_FIT_RAYS", 512)), "tc_mixed16"))
