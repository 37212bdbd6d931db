#!/bin/bash
# Full GPU session: tests, full-size configs, bench (ours + reference arm), ncu launch list of the bench command,
# full ncu captures of the top kernels with summaries.   Usage (under gpurun): bash scripts/gpu_round3.sh <tag>
TAG=${1:-r1n}
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/pytest_$TAG.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_$TAG.log
timeout 900 python scripts/configs_bench.py --tag $TAG > gpurun_out/configs_$TAG.log 2>&1; echo "configs rc=$?"
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; echo "bench rc=$?"
tail -c 1200 gpurun_out/bench_$TAG.json
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref_$TAG.json 2> gpurun_out/bench_ref_$TAG.err; echo "ref rc=$?"
cat gpurun_out/bench_ref_$TAG.json
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k 'regex:xdist|minplus|lt_|edt_|uf_|flood' -c 500 --csv \
    --log-file gpurun_out/launches_$TAG.csv python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/launches_$TAG.log 2>&1
python scripts/launch_shares.py gpurun_out/launches_$TAG.csv gpurun_out/bench_$TAG.json > gpurun_out/launch_shares_$TAG.txt; cat gpurun_out/launch_shares_$TAG.txt
bash scripts/ncu_one.sh ${TAG}_edt 0 6
bash scripts/ncu_one.sh ${TAG}_k5 21 3
bash scripts/ncu_one.sh ${TAG}_bit 27 3
python scripts/ncu_summary.py gpurun_out/prof_${TAG}_edt.raw.csv gpurun_out/prof_${TAG}_k5.raw.csv gpurun_out/prof_${TAG}_bit.raw.csv > gpurun_out/ncu_summary_$TAG.txt
cat gpurun_out/ncu_summary_$TAG.txt
