#!/bin/bash
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
echo "== graph repro"; for a in "shard 512 3"; do HONERF_PRECISION=tc_mixed16 timeout 60 python tools/dbg_graph.py $a 2>&1 | grep -E "^ok|HANG|Error|instance" | cut -c1-600 | head -8; done
echo "== mixed16 operator"; timeout 300 python -m pytest tests/test_gpu_mixed16.py -q -s -k "not dw16" > gpurun_out/r2g_m16.log 2>&1; grep -E "^n=|^worst|passed|failed|Error" gpurun_out/r2g_m16.log | cut -c1-600
echo "== bench-shape parity under mixed16"; HONERF_PRECISION=tc_mixed16 timeout 400 python -m pytest tests/test_gpu_bench_shape.py -q -s > gpurun_out/r2g_shape.log 2>&1; grep -E "^n_rays=|^n=|^worst|^rays|passed|failed|Error|d_pts" gpurun_out/r2g_shape.log | cut -c1-600
echo "== bench mixed16"; timeout 300 python bench.py --precision tc_mixed16 --no-cpu-baseline --large-rays 0 --fit-rays 0 --grid-res 0 > gpurun_out/r2g_bench_m16.json 2> gpurun_out/r2g_bench_m16.err; python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2g_bench_m16.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['e2e']['value'], d['gpu_launches'])
for f in d['roofline']['families']: print(f['kernel'], f['launches_per_step'], round(f['ms_per_step'],3), round(f['frac'] or 0,3))
PY
echo "== wait breakdown"; timeout 120 python tools/prof_m16.py 2>&1 | grep inst
echo "== ncu full (mixed16)"
PROF_PRECISION=tc_mixed16 timeout 500 ncu --set full --clock-control none --import-source on -k regex:"trunk16|nsweep16|bwd16|dw16" -s 4 -c 4 -o gpurun_out/r2g_prof python tools/prof_target.py > gpurun_out/r2g_ncu.log 2>&1
ls -la gpurun_out/r2g_prof.ncu-rep
