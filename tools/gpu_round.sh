#!/bin/bash
# One GPU-box pass: new-row parity tests, bench line, full GPU suite, loss A/B, ncu launch list.  Every stage writes
# under gpurun_out/ as soon as it finishes, so a cut-off call still leaves results.
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
echo "== 8f tests" ; timeout 300 python -m pytest tests/test_gpu_8f.py -q -x 2>&1 | tail -25 | tee gpurun_out/tests_8f.log
echo "== bench (fused loss)"; timeout 420 python bench.py > gpurun_out/bench_fused.json 2> gpurun_out/bench_fused.err; tail -c 600 gpurun_out/bench_fused.json; tail -3 gpurun_out/bench_fused.err
echo "== full gpu suite"; timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 | tee gpurun_out/tests_gpu.log
echo "== bench (torch loss, main measurement only)"; timeout 200 python bench.py --loss torch --no-cpu-baseline --large-rays 0 --fit-rays 0 --grid-res 0 > gpurun_out/bench_torchloss.json 2> gpurun_out/bench_torchloss.err; tail -c 300 gpurun_out/bench_torchloss.json
echo "== smoke"; timeout 200 python __graft_entry__.py smoke 2>&1 | tail -3 | tee gpurun_out/smoke.log
echo "== ncu launch list"; timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 300 -c 300 --csv --log-file gpurun_out/launches.csv python bench.py --no-graph --steps 2 --warmup 3 --no-cpu-baseline --large-rays 0 --fit-rays 0 --grid-res 0 > gpurun_out/ncu_bench.log 2>&1; wc -l gpurun_out/launches.csv
