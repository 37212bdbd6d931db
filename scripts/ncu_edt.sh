#!/bin/bash
# ncu_edt.sh <tag> [size]: full capture of the kernels of one standalone EDT (x pass, y pass, z pass), raw + source CSV
TAG=$1; SIZE=${2:-1024}
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k 'regex:xdist|minplus' -c 6 -f -o gpurun_out/prof_$TAG \
    python scripts/ncu_target.py $SIZE edt > gpurun_out/ncu_$TAG.log 2>&1
ncu -i gpurun_out/prof_$TAG.ncu-rep --page raw --csv > gpurun_out/prof_$TAG.raw.csv 2>/dev/null
ncu -i gpurun_out/prof_$TAG.ncu-rep --page source --csv > gpurun_out/prof_$TAG.src.csv 2>/dev/null
rm -f gpurun_out/prof_$TAG.ncu-rep
python scripts/ncu_summary.py gpurun_out/prof_$TAG.raw.csv
