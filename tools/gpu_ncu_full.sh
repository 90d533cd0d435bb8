#!/bin/bash
# One `ncu --set full` capture of selected kernels from a short bench run.
# Usage: bash tools/gpu_ncu_full.sh <tag> <kernel regex> [skip] [count] [tokens]
tag=$1; rx=$2; skip=${3:-12}; count=${4:-2}; tokens=${5:-8}
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"$rx" -s $skip -c $count -f -o gpurun_out/full_$tag \
  python bench.py --steps 1 --warmup 0 --tokens $tokens --no-e2e --no-cpu-baseline --no-profile --no-graph > gpurun_out/full_$tag.log 2>&1
echo "rc $?"
ncu -i gpurun_out/full_$tag.ncu-rep --page raw --csv > gpurun_out/full_${tag}_raw.csv 2>/dev/null
ncu -i gpurun_out/full_$tag.ncu-rep --page details > gpurun_out/full_${tag}_details.txt 2>/dev/null
ls -la gpurun_out | grep full_$tag
