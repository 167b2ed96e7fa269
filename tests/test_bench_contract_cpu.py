"""The driver-facing contract of bench.py, checked without a GPU: the reference arm's JSON line (run here on a tiny
sample), and the committed artefacts under profiles/ (bench line schema, roofline arithmetic, whole-step launch list)."""
import csv
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BASE_KEYS = {"metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
             "vs_baseline", "dtype", "data", "config", "e2e", "gpu_launches"}


def test_reference_arm_prints_one_json_line():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1",
                        "--rays", "32"], capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert BASE_KEYS <= set(d) and d["impl"] == "reference"
    assert d["unit"] == "rays/s" and d["higher_is_better"] is True and d["value"] > 0 and d["steps"] == 1
    assert d["vs_baseline"] is None and d["gpu_launches"] == 0
    # "reference" where the reference tree is reachable (this build container: /root/reference through oracle/ref_loader.py),
    # "port" on the GPU box; the arm honours the warm-up it is asked for
    have_ref = os.path.isfile("/root/reference/utils/renderer.py") or os.path.isfile(os.path.join(ROOT, "baseline", "_ref", "utils", "renderer.py"))
    assert d["cpu_baseline"]["kind"] == ("reference" if have_ref else "port")
    assert d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"] and d["warmup"] == 1
    assert d["forward_only"]["value"] > 0
    assert d["e2e"] == {"value": d["value"], "unit": "rays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in d["config"] and "model" not in d["config"]


def test_committed_bench_line_schema_and_roofline_arithmetic():
    d = json.load(open(os.path.join(ROOT, "profiles", "r02_bench.json")))
    assert BASE_KEYS | {"clocks", "roofline", "cpu_baseline"} <= set(d)
    assert d["n_gpus"] == 1 and d["unit"] == "rays/s" and d["scaling"] == "weak" and d["data"] == "synthetic"
    assert d["gpu_launches"] > 0 and d["warmup"] >= 3
    assert abs(d["value"] - 512 / (d["ms_per_step"] * 1e-3)) < 1e-6 * d["value"]
    e = d["e2e"]
    assert e["h2d_bytes_per_step"] > 0 and e["d2h_bytes_per_step"] > 0 and 0 < e["value"] < d["value"] * 1.02
    c = d["clocks"]
    assert c["samples"] >= 1 and c["sm_mhz"] <= c["sm_max_mhz"]
    assert not {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"} & set(c["reasons"])
    r = d["roofline"]
    assert r["bound"] in ("hbm", "tensor") and r["unit"] == "TFLOP/s"
    assert abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9
    # achieved = algorithmic FLOPs per launch / measured launch duration (2 F_o per point of the 512 x 128 batch)
    assert r["algorithmic_flops_per_launch"] == 512 * 128 * 2 * 1_049_088
    assert abs(r["achieved"] - r["algorithmic_flops_per_launch"] / (r["ms_per_launch"] * 1e-3) / 1e12) < 1e-6 * r["achieved"]
    # measured DRAM traffic of the dominant kernel >= its algorithmic bytes, roughly (HN_TC_MIXED16 backward: 16-bit tiles,
    # EM + D16 read, X16 written and read back, U16 + DZ16 written: 48 x 512 B per point), and stamped with the build it was
    # captured on
    assert r["traffic"] is None or r["traffic"] >= 0.9 * 512 * 128 * 48 * 512
    assert r["traffic"] is None or "build" in r["traffic_source"]
    b = d["cpu_baseline"]
    assert b["kind"] in ("port", "reference") and b["cores"] >= 1 and b["value"] > 0 and b["sample"]
    fams = {f["kernel"]: f for f in r["families"]}
    assert r["kernel"] in fams and fams[r["kernel"]]["ms_per_step"] == max(f["ms_per_step"] for f in fams.values())


def test_committed_launch_list_is_whole_steps():
    """profiles/r02_step_launches.csv = whole train steps (each ends with the optimiser launch), so kernel shares are
    those of complete steps; the dominant kernel of the bench line is the dominant kernel of the capture too."""
    rows = list(csv.DictReader(open(os.path.join(ROOT, "profiles", "r02_step_launches.csv"))))
    names = [r["Kernel Name"] for r in rows]
    ends = [i for i, n in enumerate(names) if "adam_flat_kernel" in n]
    assert len(ends) == 3 and ends[-1] == len(rows) - 1
    per_step = [b - a for a, b in zip([-1] + ends[:-1], ends)]
    assert len(set(per_step)) == 1                       # the same launch sequence every step
    tot = {}
    for r in rows:
        v = float(r["Metric Value"].replace(",", ""))
        v = v / 1000 if r["Metric Unit"] == "ns" else v
        key = r["Kernel Name"].split("(")[0]
        tot[key] = tot.get(key, 0.0) + v
    top = max(tot, key=tot.get)
    d = json.load(open(os.path.join(ROOT, "profiles", "r02_bench.json")))
    family = {"chain::bwd16_kernel": "chain::sdf_bwd_kernel", "chain::trunk16_kernel": "chain::sdf_fwd_kernel"}     # HN_TC_MIXED16 names
    assert family.get(top, top) == d["roofline"]["kernel"]
    # the kernel's SHARE of the step agrees between the serialised, cold-cache capture and the live event timing of bench.py
    share_ncu = tot[top] / sum(tot.values())
    fams = {f["kernel"]: f["ms_per_step"] for f in d["roofline"]["families"]}
    share_live = fams[d["roofline"]["kernel"]] / d["ms_per_step"]
    assert abs(share_ncu - share_live) < 0.06, (share_ncu, share_live)
