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
          "dw_kernel": "chain::dw_kernel", "color_fwd": "chain::color_fwd_kernel", "color_bwd": "chain::color_bwd_kernel"}


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
                per_family.setdefault(fam, []).append(k.get("dram__bytes_read.sum", 0.0) + k.get("dram__bytes_write.sum", 0.0))
    out = {"source": command,
           # largest launch of each family (the step's dominant configuration), bytes per launch
           "dram_bytes_per_launch": {fam: max(v) for fam, v in per_family.items()},
           "kernels": kernels}
    json.dump(out, open(dst, "w"), indent=1)
    print("wrote", dst, "with", len(kernels), "launches")


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2], sys.argv[3] if len(sys.argv) > 3 else "")
