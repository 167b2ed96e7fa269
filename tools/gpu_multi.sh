#!/bin/bash
# N-GPU bench exactly as the driver launches it (torchrun, one rank per GPU), both arms.
N=${1:-2}
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
echo "== bench --gpus $N"
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err
tail -c 1200 gpurun_out/bench_n$N.json; tail -3 gpurun_out/bench_n$N.err
echo "== reference arm --gpus $N"
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29518 bench.py --impl reference --gpus $N --steps 3 --warmup 1 > gpurun_out/bench_ref_n$N.json 2> gpurun_out/bench_ref_n$N.err
tail -c 600 gpurun_out/bench_ref_n$N.json
