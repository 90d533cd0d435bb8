#!/bin/bash
# ncu launch list of a few decode steps only (warm L2: --cache-control none), summarised per kernel.
# Usage: bash tools/gpu_decode_list.sh <tag> [skip] [count]
tag=${1:-dec}; skip=${2:-2500}; count=${3:-360}
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none -s $skip -c $count --csv \
  --log-file gpurun_out/declist_$tag.csv \
  python bench.py --steps 1 --warmup 0 --tokens 40 --no-e2e --no-cpu-baseline --no-profile --no-graph > gpurun_out/declist_$tag.log 2>&1
echo "rc $?"
python tools/ncu_summarize.py gpurun_out/declist_$tag.csv > gpurun_out/declist_$tag.md
cat gpurun_out/declist_$tag.md
