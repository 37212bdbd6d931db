#!/bin/bash
# Multi-GPU session: sharded parity check and the weak-scaling bench line on N GPUs of one box.
# Usage: bash scripts/gpu_round_multi.sh <N> <tag>
N=${1:-2}; TAG=${2:-r1}
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
    tests/sharded_gpu_check.py > gpurun_out/sharded_check_${TAG}_x$N.log 2>&1; echo "sharded check rc=$?"
tail -5 gpurun_out/sharded_check_${TAG}_x$N.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 \
    bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/bench_${TAG}_x$N.json 2> gpurun_out/bench_${TAG}_x$N.err; echo "bench rc=$?"
tail -c 3000 gpurun_out/bench_${TAG}_x$N.json; tail -5 gpurun_out/bench_${TAG}_x$N.err
