#!/bin/bash
# GPU session A: full GPU test-suite, full-size config runs, bench line.  Usage: bash scripts/gpu_round2.sh <tag>
TAG=${1:-r1d}
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/pytest_$TAG.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest_$TAG.log
timeout 900 python scripts/configs_bench.py --tag $TAG > gpurun_out/configs_$TAG.log 2>&1; echo "configs rc=$?"; tail -c 6000 gpurun_out/configs_$TAG.log
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; echo "bench rc=$?"
tail -c 2500 gpurun_out/bench_$TAG.json; tail -3 gpurun_out/bench_$TAG.err
