#!/bin/bash
# Full ncu captures of selected launches of one local_thickness call (see scripts/ncu_target.py).
# Launch order among our kernels: xdist<EDT>, minplus y, minplus z, classify, then per radius k:
# xdist<LT> (4+3k), lt_y2 (5+3k), lt_zsweep (6+3k).  Usage: ncu_capture.sh <tag> [size]
TAG=${1:-r1}; SIZE=${2:-1024}
K='regex:xdist|minplus|lt_y2|zsweep|classify|expand|point'
mkdir -p gpurun_out
run() { # name skip count
  ncu --set full --clock-control none --import-source on -k "$K" -s $2 -c $3 -f -o gpurun_out/prof_${TAG}_$1 \
      python scripts/ncu_target.py $SIZE > gpurun_out/ncu_${TAG}_$1.log 2>&1
  ncu -i gpurun_out/prof_${TAG}_$1.ncu-rep --page raw --csv > gpurun_out/prof_${TAG}_$1.raw.csv 2>/dev/null
}
run edt 0 4
run k5 19 3
run k20 64 3
ls -la gpurun_out
