#!/bin/bash
# ncu evidence for the round (run on the GPU box via gpurun): launch list of a bench step + one full capture of
# the dominant decode kernels.  Usage: bash tools/gpu_profile.sh <tag>
tag=${1:-r01}
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 60000 --csv \
  --log-file gpurun_out/launches_$tag.csv \
  python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --no-profile > gpurun_out/ncu_list_$tag.log 2>&1
echo "launch list rc $?"
python tools/ncu_summarize.py gpurun_out/launches_$tag.csv > gpurun_out/launches_$tag.md
gzip -f gpurun_out/launches_$tag.csv
timeout 900 ncu --set full --clock-control none --import-source on \
  -k regex:'decode_cross_attn|decode_self_attn|gemm_tc_kernel|sample_step' -s 600 -c 24 -o gpurun_out/prof_$tag -f \
  python bench.py --steps 1 --warmup 0 --tokens 24 --no-e2e --no-cpu-baseline --no-profile --no-graph > gpurun_out/ncu_full_$tag.log 2>&1
echo "full capture rc $?"
ncu -i gpurun_out/prof_$tag.ncu-rep --page raw --csv > gpurun_out/prof_${tag}_raw.csv 2>/dev/null
ls -la gpurun_out
