#!/bin/bash
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
echo "== full gpu suite"; timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -30 > gpurun_out/r2i_suite.log; tail -12 gpurun_out/r2i_suite.log | cut -c1-400
echo "== bench (default flags)"; timeout 600 python bench.py > gpurun_out/r2i_bench.json 2> gpurun_out/r2i_bench.err; python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2i_bench.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['e2e']['value'], d['gpu_launches'], d['clocks'])
for f in d['roofline']['families']: print(f['kernel'], f['launches_per_step'], round(f['ms_per_step'],3), round(f['frac'] or 0,3))
print('roofline', {k:d['roofline'][k] for k in ('kernel','achieved','frac','traffic','traffic_source')})
for k in ('sdf_grid','fitting_step','forward_only','large_batch','cpu_baseline','compositor'): print(k, json.dumps(d[k])[:700])
PY
tail -3 gpurun_out/r2i_bench.err | cut -c1-300
