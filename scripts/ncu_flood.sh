#!/bin/bash
# ncu_flood.sh <tag> [size] [skip]: full capture of a few flood launches (collect / union / resolve) of one
# access-limited porosimetry call, raw + source CSV
TAG=$1; SIZE=${2:-1024}; SKIP=${3:-480}
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k 'regex:uf_collect|uf_union' -s $SKIP -c 4 -f -o gpurun_out/prof_${TAG}_a \
    python scripts/ncu_target.py $SIZE poro50 > gpurun_out/ncu_${TAG}_a.log 2>&1
ncu --set full --clock-control none --import-source on -k 'regex:uf_resolve' -c 1 -f -o gpurun_out/prof_${TAG}_b \
    python scripts/ncu_target.py $SIZE poro50 > gpurun_out/ncu_${TAG}_b.log 2>&1
for s in a b; do
  ncu -i gpurun_out/prof_${TAG}_$s.ncu-rep --page raw --csv > gpurun_out/prof_${TAG}_$s.raw.csv 2>/dev/null
  ncu -i gpurun_out/prof_${TAG}_$s.ncu-rep --page source --csv > gpurun_out/prof_${TAG}_$s.src.csv 2>/dev/null
  rm -f gpurun_out/prof_${TAG}_$s.ncu-rep
  python scripts/ncu_summary.py gpurun_out/prof_${TAG}_$s.raw.csv
done
