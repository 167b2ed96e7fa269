#!/bin/bash
# ncu launch list of whole train steps (same command line as the bench's timed region, eager launches) + bench line.
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
echo "== bench"; timeout 420 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; tail -c 300 gpurun_out/bench.json; tail -2 gpurun_out/bench.err
echo "== ncu launch list"; timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/launches_full.csv python bench.py --no-graph --no-roofline --steps 3 --warmup 3 --no-cpu-baseline --large-rays 0 --fit-rays 0 --grid-res 0 > gpurun_out/ncu_bench.log 2>&1; wc -l gpurun_out/launches_full.csv
python tools/launch_summary.py gpurun_out/launches_full.csv 30 --steps 3 --out gpurun_out/launches.csv | tee gpurun_out/launch_summary.md
