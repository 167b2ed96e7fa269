"""Condense an `ncu --set full` report into the small JSON kept under profiles/ (and read by bench.py for the
`traffic` field of its roofline object).

    ncu -i gpurun_out/X.ncu-rep --page raw --csv > /tmp/x.csv
    python tools/ncu_metrics.py /tmp/x.csv profiles/rNN_chain_ncu_metrics.json "<the ncu command line>"
"""
import csv
import json
import re
import sys

KEEP = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "launch__registers_per_thread",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "launch__grid_size", "launch__block_size",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio"]
SCALE = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "us": 1e-3, "ms": 1.0, "ns": 1e-6, "s": 1e3}
FAMILY = {"sdf_only": "chain::sdf_only_kernel", "sdf_fwd": "chain::sdf_fwd_kernel", "sdf_bwd": "chain::sdf_bwd_kernel",
          "dw_kernel": "chain::dw_kernel", "color_fwd": "chain::color_fwd_kernel", "color_bwd": "chain::color_bwd_kernel",
          # HN_TC_MIXED16: the forward is two launches (their traffic adds up), the backward one, the weight gradients one
          "trunk16": "chain::sdf_fwd_kernel", "nsweep16": "chain::sdf_fwd_kernel", "bwd16": "chain::sdf_bwd_kernel",
          "dw16": "chain::dw_kernel"}
ADDITIVE = ("trunk16", "nsweep16")


def build_id(root=None):
    """sha1 over the CUDA sources of the object-field chain kernels (the families the bench's roofline quotes traffic for) and
    the headers they include: bench.py refuses a traffic figure whose build differs from the library it is timing."""
    import hashlib
    import os
    root = root or os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "ho-nerf_b200", "csrc")
    h = hashlib.sha1()
    for f in sorted(os.listdir(root)):
        if f.endswith((".cu", ".cuh")) and ((f.startswith("chain") and "hand" not in f) or f in ("tc_common.cuh", "common.cuh")):
            h.update(f.encode())
            h.update(open(os.path.join(root, f), "rb").read())
    return h.hexdigest()[:16]


def main(src, dst, command):
    rows = list(csv.reader(open(src)))
    hdr, units, body = rows[0], rows[1], rows[2:]
    col = {h: i for i, h in enumerate(hdr)}
    kernels, per_family = [], {}
    for r in body:
        name = re.sub(r"\(.*", "", r[col["Kernel Name"]]).replace("hn::", "").replace("void ", "")
        k = {"kernel": name}
        for m in KEEP:
            if m in col and r[col[m]] != "":
                k[m] = float(r[col[m]].replace(",", "")) * SCALE.get(units[col[m]], 1.0)
        kernels.append(k)
        for key, fam in FAMILY.items():
            if key in name:
                per_family.setdefault((fam, key), []).append(k.get("dram__bytes_read.sum", 0.0) + k.get("dram__bytes_write.sum", 0.0))
    fam_bytes = {}
    for (fam, key), v in per_family.items():
        # largest launch of each kernel (the step's dominant configuration); a family made of two kernels adds them up
        fam_bytes[fam] = fam_bytes.get(fam, 0.0) + max(v) if key in ADDITIVE else max(fam_bytes.get(fam, 0.0), max(v))
    out = {"source": command, "build_id": build_id(),
           "dram_bytes_per_launch": fam_bytes,
           "kernels": kernels}
    json.dump(out, open(dst, "w"), indent=1)
    print("wrote", dst, "with", len(kernels), "launches")


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2], sys.argv[3] if len(sys.argv) > 3 else "")
