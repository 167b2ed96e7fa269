#!/bin/bash
# A/B of NeuSRenderer.ray_streams on the GPU box: parity test, then the main bench measurement with 1 / 2 / 3 / 4 shards.
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
echo "== ray-streams test"; timeout 300 python -m pytest tests/test_gpu_product_default.py -q -x 2>&1 | tail -25 | tee gpurun_out/tests_streams.log
for k in 1 2 3 4; do
  echo "== bench --ray-streams $k"
  timeout 200 python bench.py --ray-streams $k --no-cpu-baseline --large-rays 0 --fit-rays 0 --grid-res 0 > gpurun_out/bench_rs$k.json 2> gpurun_out/bench_rs$k.err
  python - <<PY
import json
try:
    d = json.load(open("gpurun_out/bench_rs$k.json"))
    print("ray_streams=$k", round(d["value"]), "rays/s", round(d["ms_per_step"], 4), "ms; e2e", round(d["e2e"]["value"]), "graph", d["config"]["cuda_graph"], "clocks", d["clocks"])
except Exception as e:
    print("no json:", e)
PY
  tail -2 gpurun_out/bench_rs$k.err
done
