"""Writes profiles/r02_summary.md from the committed round-2 artefacts (bench lines, ncu metrics)."""
import json
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
P = lambda n: os.path.join(ROOT, "profiles", n)
b = json.load(open(P("r02_bench.json")))
multi = {n: json.load(open(P("r02_bench_%dgpu.json" % n))) for n in (2, 4, 8) if os.path.isfile(P("r02_bench_%dgpu.json" % n))}
n = json.load(open(P("r02_ncu_metrics.json")))
h = json.load(open(P("r02_hand_ncu_metrics.json")))
ref = json.load(open(P("r02_bench_reference.json")))
r = b["roofline"]
L = []
A = L.append
A("# Round 2 profile summary (1x B200 unless stated; final build of the round)\n")
A("Artefacts: `r02_bench.json` (the `python bench.py` line), `r02_bench_reference.json` (`--impl reference`), `r02_bench_{2,4,8}gpu.json`")
A("(torchrun as the driver launches it), `r02_ncu_metrics.json` (one `ncu --set full` launch of every object chain kernel, stamped with")
A("the build id `bench.py` checks before quoting `traffic`), `r02_step_launches.csv` (three whole graph-replayed steps of the bench")
A("command under `ncu --metrics gpu__time_duration.sum`), `r02_hand_ncu_metrics.json` (hand-field chain kernels), `r02_sass_counts.txt`")
A("(cuobjdump instruction counts per kernel), `r02_precision_table.md`.  Regenerate this file with `python tools/r02_summary.py`.\n")
A("## Headline\n")
A("| | r01 | r02 |\n|---|---:|---:|")
A("| step (512 rays x (64+64), fwd + 2nd-order bwd + Adam, CUDA graph) | 3.04 ms = 168.4k rays/s | **%.3f ms = %.1fk rays/s** |" % (b["ms_per_step"], b["value"] / 1e3))
A("| end to end (host inputs, H2D + D2H inside the timed region) | 166.4k rays/s | **%.1fk rays/s** |" % (b["e2e"]["value"] / 1e3))
A("| reference arm (oracle port on the box's 16 host cores) | 420 rays/s | %.0f rays/s |" % ref["value"])
for k, m in multi.items():
    A("| %d GPUs (weak scaling, NCCL all-reduce captured in the step's graph; earlier build, 1-GPU line of that box 2.283 ms) | %s | %.1fk rays/s, %.3f ms/step |" % (
        k, {2: "326.9k (98 %)", 4: "634.0k (98 %)"}.get(k, "-"), m["value"] / 1e3, m["ms_per_step"]))
ab_path = P("r02_peer_exchange_ab.json")
if os.path.isfile(ab_path):
    ab = json.load(open(ab_path))["ms_per_step"]
    mean = lambda v: sum(v) / len(v)
    A("| gradient exchange fused with Adam over NVLink peer memory (`hn_peer_adam_flat`), ms/step peer vs NCCL, same box | - | "
      "2 GPUs **%.3f** vs %.3f; 4 GPUs **%.3f** vs %.3f; 8 GPUs **%.3f** (%.0fk rays/s; NCCL on an earlier box: %.3f) |" % (
          mean(ab["2"]["peer"]), mean(ab["2"]["nccl"]), ab["4"]["peer"][0], ab["4"]["nccl"][0], ab["8"]["peer"][0],
          8 * 512 / ab["8"]["peer"][0], ab["8"]["nccl_earlier_box"][0]))
fam = {f["kernel"]: f for f in r["families"]}
dram = sum((f["traffic"] or 0) * f["launches_per_step"] for f in r["families"] if f["kernel"] not in ("chain::sdf_fwd_kernel", "chain::dw_kernel"))
dram += (fam["chain::sdf_fwd_kernel"]["traffic"] or 0)       # trunk + normal sweep already added up per step
k16 = [k for k in n["kernels"] if "dw16" in k["kernel"]]
dram += sum(k.get("dram__bytes_read.sum", 0) + k.get("dram__bytes_write.sum", 0) for k in k16)
A("| DRAM traffic of the step's chain kernels (ncu, per step) | ~9.6 GB (`r01_chain_ncu_metrics.json`) | **%.2f GB** |" % (dram / 1e9))
A("| 4 096 rays / step | 184k rays/s | %.1fk rays/s |" % (b["large_batch"]["value"] / 1e3))
A("| SDF lattice 512^3 | 399 ms | %.0f ms; %s |" % (b["sdf_grid"]["ms"], ", ".join("%d GPUs %.0f ms" % (k, m["sdf_grid"]["ms"]) for k, m in multi.items())))
f = b["fitting_step"]
A("| two-field pose-fitting iteration, 8 views x 196 rays | 79.6 ms | **%.1f ms** (views batched; %.1f ms as a per-view loop); %s |" % (
    f["ms_per_iteration"], f["per_view_loop"]["ms_per_iteration"], ", ".join("%d GPUs %.1f ms" % (k, m["fitting_step"]["ms_per_iteration"]) for k, m in multi.items())))
A("| one 512-ray fitting step (192 samples x 2 fields) | 24.5 ms | **%.2f ms** |" % f["one_512_ray_batch"]["ms_per_step"])
fo = b["forward_only"]
A("| hand-field forward render | 69-88k rays/s | **%.1fk rays/s** (4 096 rays), %.1fk (one 512 x 512 view); %s |" % (
    fo["hand_render_fwd"]["value"] / 1e3, fo["hand_views"]["value"] / 1e3,
    ", ".join("%d views on %d GPUs %.0fk" % (k, k, m["forward_only"]["hand_views"]["value"] / 1e3) for k, m in multi.items())))
A("| compositor stand-alone fwd / bwd (2^18 rays x 128) | 1.06 / 0.69 of the HBM copy peak | %.2f / %.2f |\n" % (b["compositor"]["fwd"]["frac"], b["compositor"]["bwd"]["frac"]))
A("## Roofline of the step's kernel families (live CUDA-event timing inside `bench.py`; peak = %.1f TFLOP/s, %s)\n" % (r["peak"], r["peak_source"]))
A("| family | launches / step | ms / step | algorithmic TFLOP/s | frac | DRAM bytes / launch (ncu) |\n|---|---:|---:|---:|---:|---:|")
for fm in r["families"]:
    A("| `%s` | %d | %.3f | %.1f | %.3f | %s |" % (fm["kernel"], fm["launches_per_step"], fm["ms_per_step"], fm["achieved"] or 0, fm["frac"] or 0,
                                              ("%.2f GB" % (fm["traffic"] / 1e9)) if fm["traffic"] else "-"))
A("| all MLP kernels | | %.3f | %.1f | %.3f | |\n" % (r["step"]["mlp_ms_per_step"], r["step"]["achieved"], r["step"]["frac"]))
A("Every product is three 16-bit MMAs, so the algorithmic ceiling is 1/3 of the dense peak (frac 0.333); `frac` above is against the")
A("un-split dense peak as the contract asks.  HN_TC_MIXED16 names: `sdf_fwd` = `trunk16_kernel` + `nsweep16_kernel`, `sdf_bwd` =")
A("`bwd16_kernel`, `dw` = two `dw16_kernel` launches (SDF net: bf16 tiles, one MMA per product; colour net: hi / lo tile pairs, three).\n")
A("## ncu (`--set full`, one launch each, 65 536 points; cold cache, serialised)\n")
A("| kernel | ms | DRAM GB (% of peak) | tensor pipe active | regs | long-scoreboard stalls / issue |\n|---|---:|---:|---:|---:|---:|")
for k in n["kernels"]:
    A("| `%s` | %.3f | %.2f (%.0f %%) | %.1f %% | %d | %.1f |" % (
        k["kernel"], k["gpu__time_duration.sum"], (k.get("dram__bytes_read.sum", 0) + k.get("dram__bytes_write.sum", 0)) / 1e9,
        k.get("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", 0), k.get("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", 0),
        k.get("launch__registers_per_thread", 0), k.get("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", 0)))
dom = [x for x in r["families"] if x["kernel"] == r["kernel"]][0]
A("\nShare check (contract: the kernel's share of the step must agree between the capture and the live timing): see")
A("`tests/test_bench_contract_cpu.py::test_committed_launch_list_is_whole_steps`; live share of `%s`: %.1f %%.\n" % (r["kernel"], 100 * dom["ms_per_step"] / b["ms_per_step"]))
A("## Hand-field chain kernels (98 304 points = 512 rays x 192 samples, weights frozen)\n")
A("| kernel | ms | DRAM GB (% of peak) | tensor pipe active | issue active |\n|---|---:|---:|---:|---:|")
seen = set()
for k in h["kernels"]:
    if k["kernel"] in seen:
        continue
    seen.add(k["kernel"])
    A("| `%s` | %.3f | %.2f (%.0f %%) | %.1f %% | %.1f %% |" % (
        k["kernel"], k["gpu__time_duration.sum"], (k.get("dram__bytes_read.sum", 0) + k.get("dram__bytes_write.sum", 0)) / 1e9,
        k.get("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", 0), k.get("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", 0),
        k.get("smsp__issue_active.avg.pct_of_peak_sustained_active", 0)))
A("\nHand SDF operator (value + feature + normal forward, second-order backward to points and bone transforms), CUDA events:")
A("per-layer `tc_bf16x3` 6.2 + 11.6 ms -> pose-gradient accumulation in registers, no copies 6.2 + 8.9 -> 256 x 256 layers as chain")
A("kernels 3.84 + 4.41 -> feature-side cotangent contractions as chunk steps of the sweeps (first attempt: 3.46 + 3.92, the")
A("accumulate pass chained load -> add -> store sixteen times per sub-block; with the loads batched) **2.71 + 2.98 ms**; SDF-only")
A("query 2.4 -> 1.39 ms.  MMA-issuer cycle counters (`tools/prof_hand16.py`): normal sweep 1 425 kcycles per CTA (A operand 48 %,")
A("weights 16 %, issuing 36 %), tangent + reverse sweep 2 157 kcycles (53 / 13 / 34 %).  Forward-only render of 4 096 rays")
A("(`tools/prof_hand_render.py`, 23.4 ms before the last changes): trunk x5 4.4 ms, normal sweep 4.4, the 1386-wide input")
A("contractions 5.9, hand colour net 6.0 (2.8 of it assembling its 1672-wide input row: vectorised since, 21.5 ms), HALO")
A("feature / normal kernels 2.5.\n")
px = P("r02_peer_exchange.json")
if os.path.isfile(px):
    pe = json.load(open(px))
    c, f0 = pe["final_kernel"], pe["first_version_one_remote_load_per_loop_trip"]
    A("## The step's exchange: gradient sum over the ranks + Adam (2x B200, 824 064 floats, `tools/prof_peer.py`, graphs of 50 launches)\n")
    A("| variant | us / launch |\n|---|---:|")
    A("| `ncclAllReduce` + `hn_adam_flat` | %.1f |" % c["nccl_allreduce+hn_adam_flat_us"])
    A("| `hn_adam_flat` alone | %.1f |" % c["hn_adam_flat_alone_us"])
    A("| `hn_peer_adam_flat`, first version (one remote load per thread and loop trip), 148 / 64 / 16 / 8 CTAs | %.1f / %.1f / %.1f / %.1f |" % (
        f0["hn_peer_adam_flat_148_ctas_us"], f0["hn_peer_adam_flat_64_ctas_us"], f0["hn_peer_adam_flat_16_ctas_us"], f0["hn_peer_adam_flat_8_ctas_us"]))
    A("| `hn_peer_adam_flat`, remote loads batched, 1 024 threads, 148 / 64 / 16 CTAs | **%.1f** / %.1f / %.1f |" % (
        c["hn_peer_adam_flat_148_ctas_us"], c["hn_peer_adam_flat_64_ctas_us"], c["hn_peer_adam_flat_16_ctas_us"]))
    A("| one rank 204 us late: NCCL form / peer kernel | %.1f / %.1f |" % (c["late_rank:nccl_allreduce+hn_adam_flat_us"], c["late_rank:hn_peer_adam_flat_us"]))
    A("| host in the loop (50 us kernel + exchange + blocking D2H per step): NCCL / peer / Adam only | %.1f / %.1f / %.1f |" % (
        c["host_in_loop:sleep50us+nccl+adam+d2h_us"], c["host_in_loop:sleep50us+peer+d2h_us"], c["host_in_loop:sleep50us+adam_only+d2h_us"]))
    A("")
    A("The kernel is latency-bound (time ~ 10 us + 3 us per sequential NVLink round trip: the CTA sweep of the first version), not")
    A("bandwidth-bound: 2 x 1.65 MB cross NVLink per GPU and launch.  What the exchange costs on the step at 2 GPUs: 19 us of 2 327.\n")
A("## Step-time history of the round (512 rays, 1x B200)\n")
A("3.04 ms (r01) -> dW-ready 16-bit operands + `dw16_kernel` 2.83 -> sweeps with a single 16-bit operand in tensor memory 2.37 (rejected")
A("on parity) -> x3 sweeps with hi / lo pairs in tensor memory 2.46 -> colour net on hi / lo T16 tile pairs + `dw16_kernel` 2.41 (one MMA")
A("per product on the hi tiles alone: 2.33, colour weight gradients move by 1.6-2.9e-3, kept behind `HONERF_COLOR_DW_X3=0`) -> two ray")
A("streams instead of three **%.2f ms**." % b["ms_per_step"])
open(P("r02_summary.md"), "w").write("\n".join(L) + "\n")
print("\n".join(L[:30]))
