#!/bin/bash
# ncu_flood.sh <tag> [size]: full capture of the flood kernels of one porosimetry call (the union launch of a busy radius)
TAG=$1; SIZE=${2:-768}
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k 'regex:uf_prelink|uf_emit|uf_sort|uf_compress|uf_resolve' -c 5 -f -o gpurun_out/prof_$TAG \
    python scripts/per_launch_poro.py $SIZE 25 faces once > gpurun_out/ncu_$TAG.log 2>&1
ncu --set full --clock-control none --import-source on -k 'regex:uf_union_rec' -s 12 -c 2 -f -o gpurun_out/prof_${TAG}_u \
    python scripts/per_launch_poro.py $SIZE 25 faces once >> gpurun_out/ncu_$TAG.log 2>&1
for t in $TAG ${TAG}_u; do
  ncu -i gpurun_out/prof_$t.ncu-rep --page raw --csv > gpurun_out/prof_$t.raw.csv 2>/dev/null
  ncu -i gpurun_out/prof_$t.ncu-rep --page source --csv > gpurun_out/prof_$t.src.csv 2>/dev/null
  ncu -i gpurun_out/prof_$t.ncu-rep --page details --csv > gpurun_out/prof_$t.det.csv 2>/dev/null
  rm -f gpurun_out/prof_$t.ncu-rep
done
ls -la gpurun_out | grep $TAG
