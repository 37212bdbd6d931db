#!/bin/bash
# ncu_one.sh <tag> <skip> <count> [size] : full capture of <count> launches after <skip>, plus raw + source CSV
TAG=$1; SKIP=$2; CNT=$3; SIZE=${4:-1024}
K='regex:xdist|minplus|lt_y2|zsweep|classify|expand|point|bitball|lt_pack|wmask'
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k "$K" -s $SKIP -c $CNT -f -o gpurun_out/prof_$TAG \
    python scripts/ncu_target.py $SIZE > gpurun_out/ncu_$TAG.log 2>&1
ncu -i gpurun_out/prof_$TAG.ncu-rep --page raw --csv > gpurun_out/prof_$TAG.raw.csv 2>/dev/null
ncu -i gpurun_out/prof_$TAG.ncu-rep --page source --csv > gpurun_out/prof_$TAG.src.csv 2>/dev/null
rm -f gpurun_out/prof_$TAG.ncu-rep
ls -la gpurun_out | grep $TAG
