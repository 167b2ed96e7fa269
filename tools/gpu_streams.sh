#!/bin/bash
# A/B of the per-shard loss (no join between forward and backward) on the GPU box + its parity test.
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
timeout 300 python -m pytest tests/test_gpu_product_default.py tests/test_gpu_8f.py -q -x 2>&1 | tail -15 | tee gpurun_out/tests_streams.log
for cfg in "--shard-loss 0" "--shard-loss 1" "--shard-loss 1 --ray-streams 2" "--shard-loss 1 --ray-streams 4" "--shard-loss 1 --ray-streams 5" "--shard-loss 1 --ray-streams 6"; do
  timeout 200 python bench.py $cfg --no-cpu-baseline --no-roofline --large-rays 0 --fit-rays 0 --grid-res 0 > gpurun_out/bench_ab.json 2> gpurun_out/bench_ab.err
  python - "$cfg" <<'PY'
import json, sys
try:
    d = json.load(open("gpurun_out/bench_ab.json"))
    print("%-34s %8.0f rays/s %.4f ms; e2e %8.0f" % (sys.argv[1], d["value"], d["ms_per_step"], d["e2e"]["value"]))
except Exception as e:
    print(sys.argv[1], "no json:", e); print(open("gpurun_out/bench_ab.err").read()[-1500:])
PY
done 2>&1 | tee gpurun_out/shard_loss_ab.log
