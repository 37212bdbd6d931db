#!/bin/bash
# ncu_r3.sh <tag> : profiling pass on one GPU (see /opt/skills/guides/B200_PROFILING.md)
#   1. launch list of the bench command (gpu__time_duration only; shares, not absolutes)
#   2. `--set full` captures of every kernel family of one local_thickness(1024^3, sizes=25) call, selected by name
#   3. the flood kernels of one porosimetry(1024^3, sizes=25) call (launch list only)
TAG=${1:-r3}
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/${TAG}_launches_bench.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu --no-check > gpurun_out/${TAG}_bench_under_ncu.log 2>&1
run() { # name regex count
  ncu --set full --clock-control none --import-source on -k "regex:$2" -c $3 -f -o gpurun_out/prof_${TAG}_$1 \
      python scripts/ncu_target.py 1024 > gpurun_out/ncu_${TAG}_$1.log 2>&1
  ncu -i gpurun_out/prof_${TAG}_$1.ncu-rep --page raw --csv > gpurun_out/prof_${TAG}_$1.raw.csv 2>/dev/null
  ncu -i gpurun_out/prof_${TAG}_$1.ncu-rep --page source --csv > gpurun_out/prof_${TAG}_$1.src.csv 2>/dev/null
  rm -f gpurun_out/prof_${TAG}_$1.ncu-rep
}
run edt 'xdist_kernel|minplus16|classify|fix_inf' 5
run byte 'xdist_bits|lt_y2|lt_y3|zsweep' 21
run misc 'packn|wmask|expand' 4
run bit 'bitball' 16
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}_launches_poro.csv \
    python scripts/ncu_target.py 1024 poro > gpurun_out/${TAG}_poro_under_ncu.log 2>&1
ls -la gpurun_out | grep ${TAG}
