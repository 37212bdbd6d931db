#!/bin/bash
# ncu_r2.sh <tag> : round-2 profiling pass on one GPU (see /opt/skills/guides/B200_PROFILING.md)
#   1. launch list of the bench command (gpu__time_duration only; shares, not absolutes)
#   2. `--set full` captures of every kernel family of one local_thickness(1024^3, sizes=25) call
# Launch order of one call among our kernels: xdist<EDT> 0, minplus16 y 1, minplus y (gated) 2, minplus16 z 3,
# minplus z (gated) 4, fix_inf 5, classify 6, then per byte radius k: xdist<LT> 7+3k, lt_y2|lt_y3 8+3k, zsweep 9+3k
# (7 radii), wmask 28, packn 29, bitball 30..45, expand 46.
TAG=${1:-r2}
K='regex:xdist|minplus|lt_y|zsweep|classify|expand|point|bitball|lt_pack|wmask|fix_inf'
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/${TAG}_launches_bench.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu --no-check > gpurun_out/${TAG}_bench_under_ncu.log 2>&1
run() { # name skip count
  ncu --set full --clock-control none --import-source on -k "$K" -s $2 -c $3 -f -o gpurun_out/prof_${TAG}_$1 \
      python scripts/ncu_target.py 1024 > gpurun_out/ncu_${TAG}_$1.log 2>&1
  ncu -i gpurun_out/prof_${TAG}_$1.ncu-rep --page raw --csv > gpurun_out/prof_${TAG}_$1.raw.csv 2>/dev/null
  ncu -i gpurun_out/prof_${TAG}_$1.ncu-rep --page source --csv > gpurun_out/prof_${TAG}_$1.src.csv 2>/dev/null
  rm -f gpurun_out/prof_${TAG}_$1.ncu-rep
}
run edt 0 7
run k0 7 3
run k5 22 3
run bit 28 3
run expand 46 1
ls -la gpurun_out | grep ${TAG}
