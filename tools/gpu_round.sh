#!/bin/bash
# One GPU-box pass: full GPU suite, smoke, bench line, ncu launch list of whole train steps.  Every stage writes under
# gpurun_out/ as soon as it finishes, so a cut-off call still leaves results.
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
echo "== full gpu suite"; timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -25 | tee gpurun_out/tests_gpu.log
echo "== smoke"; timeout 200 python __graft_entry__.py smoke 2>&1 | tail -3 | tee gpurun_out/smoke.log
echo "== bench"; timeout 420 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; tail -c 300 gpurun_out/bench.json; tail -3 gpurun_out/bench.err
echo "== ncu launch list (eager launches of the same train step; cut to 3 whole steps)"
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/launches_full.csv python bench.py --no-graph --no-roofline --steps 3 --warmup 3 --no-cpu-baseline --large-rays 0 --fit-rays 0 --grid-res 0 > gpurun_out/ncu_bench.log 2>&1; wc -l gpurun_out/launches_full.csv
python tools/launch_summary.py gpurun_out/launches_full.csv 30 --steps 3 --out gpurun_out/launches.csv | tee gpurun_out/launch_summary.md
