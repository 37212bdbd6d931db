#!/bin/bash
# One GPU session: bench line, ncu launch list of the same command, full captures of the top kernels.
# Usage (under gpurun): bash scripts/gpu_round.sh <tag>
TAG=${1:-r1c}
mkdir -p gpurun_out
python bench.py --steps 10 --warmup 3 > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err
tail -c 2500 gpurun_out/bench_$TAG.json; tail -3 gpurun_out/bench_$TAG.err
ncu --metrics gpu__time_duration.sum --clock-control none -k 'regex:xdist|minplus|lt_|edt_|uf_|flood' -c 400 --csv \
    --log-file gpurun_out/launches_$TAG.csv python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/launches_$TAG.log 2>&1
bash scripts/ncu_one.sh ${TAG}_edt 0 4
bash scripts/ncu_one.sh ${TAG}_k5 19 3
bash scripts/ncu_one.sh ${TAG}_bit 25 3
