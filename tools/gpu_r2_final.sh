#!/bin/bash
# round 2, profile capture of the final build: ncu full over one launch of every chain-kernel family, the launch list of the
# bench command, and the hand-field chain kernels
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
echo "== ncu full (object chain kernels, tc_mixed16)"
PROF_PRECISION=tc_mixed16 timeout 600 ncu --set full --clock-control none --import-source on -k regex:"trunk16|nsweep16|bwd16_kernel|dw16_kernel|dw_kernel|color_fwd|color_bwd|sdf_only" -s 8 -c 8 -f -o gpurun_out/r02_prof python tools/prof_target.py > gpurun_out/r02_ncu.log 2>&1
ls -la gpurun_out/r02_prof.ncu-rep
echo "== launch list of the bench command"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/r02_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --large-rays 0 --fit-rays 0 --grid-res 0 > gpurun_out/r02_launch_bench.json 2> gpurun_out/r02_launch_bench.err
wc -l gpurun_out/r02_launches.csv
echo "== ncu full (hand chain kernels)"
PROF_HAND_COLOR=0 timeout 600 ncu --set full --clock-control none --import-source on -k regex:"hand_trunk16|hand_nsweep16|hand_bwd16|halo_bwd_tiled|halo_normal_tiled" -s 6 -c 6 -f -o gpurun_out/r02_hand_prof python tools/prof_hand.py > gpurun_out/r02_hand_ncu.log 2>&1
ls -la gpurun_out/r02_hand_prof.ncu-rep
