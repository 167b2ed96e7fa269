"""Stand-alone HBM roofline of the compositor kernels (2^18 rays x 128 samples), as bench.py reports it."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import torch  # noqa: E402

import bench  # noqa: E402
import honerf_b200 as H  # noqa: E402

print(json.dumps(bench.compositor_roofline(H, torch.device("cuda", 0))))
