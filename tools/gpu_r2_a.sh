#!/bin/bash
# round 2, call A: the new parity tests at the benchmarked shape + one `ncu --set full` pass over the chain kernels
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
echo "== new parity tests"; timeout 900 python -m pytest tests/test_gpu_bench_shape.py tests/test_gpu_hand_fit_default.py -q -s 2>&1 | tail -120 | tee gpurun_out/r2a_tests.log
echo "== ncu full (chain kernels, one 65 536-point fwd+bwd)"
timeout 500 ncu --set full --clock-control none -k regex:"sdf_bwd_kernel|sdf_fwd_kernel|dw_kernel|color_fwd|color_bwd|sdf_only" -s 8 -c 7 -o gpurun_out/r2a_prof python tools/prof_target.py > gpurun_out/r2a_ncu.log 2>&1
ls -la gpurun_out/r2a_prof.ncu-rep
