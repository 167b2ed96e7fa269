#!/bin/bash
# round 2: the peer-memory gradient exchange (csrc/peer.cu) on N GPUs of one box -- tests, kernel timing, step A/B against NCCL.
#   gpurun --gpus 2 -- 'bash tools/gpu_peer.sh 2'
N=${1:-2}
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
echo "== tests (world 1; 2 / 4 / 8 processes on one GPU; 2 GPUs over NCCL)"
timeout 400 python -m pytest tests/test_gpu_peer.py -x -q 2>&1 | tail -3
echo "== exchange + Adam, us per launch (tools/prof_peer.py)"
timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 \
    tools/prof_peer.py > gpurun_out/prof_peer_${N}gpu.json 2> gpurun_out/prof_peer_${N}gpu.err
cat gpurun_out/prof_peer_${N}gpu.json
echo "== step A/B"
for ex in peer nccl peer nccl; do
    timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
        bench.py --gpus $N --steps 30 --warmup 3 --exchange $ex --no-roofline --grid-res 0 --fit-rays 0 --strong-rays 0 \
        2> gpurun_out/peer_bench_${N}gpu_$ex.err | tee gpurun_out/peer_bench_${N}gpu_$ex.json |
        python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$ex', d['ms_per_step'], d['e2e']['ms_per_step'], d['config']['exchange_fallback'])"
done
