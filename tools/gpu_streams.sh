#!/bin/bash
# A/B of the ray-shard layout of NeuSRenderer.ray_streams on the GPU box (main bench measurement only).
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
for cfg in "--ray-streams 3" "--ray-shards 148,148,216" "--ray-shards 216,148,148" "--ray-shards 296,216" "--ray-shards 148,148,148,68" "--ray-shards 148,216,148" "--ray-shards 222,290" "--ray-streams 3 --rays 444" "--ray-streams 3 --rays 592"; do
  timeout 200 python bench.py $cfg --no-cpu-baseline --no-roofline --large-rays 0 --fit-rays 0 --grid-res 0 > gpurun_out/bench_ab.json 2> gpurun_out/bench_ab.err
  python - "$cfg" <<'PY'
import json, sys
try:
    d = json.load(open("gpurun_out/bench_ab.json"))
    print("%-34s %8.0f rays/s %.4f ms; e2e %8.0f" % (sys.argv[1], d["value"], d["ms_per_step"], d["e2e"]["value"]))
except Exception as e:
    print(sys.argv[1], "no json:", e)
PY
done 2>&1 | tee gpurun_out/streams_ab.log
