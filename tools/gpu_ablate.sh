#!/bin/bash
# In-graph cost of each decode-step kernel class by ablation (CXRM_ABLATE drops launches; outputs are garbage).
mkdir -p gpurun_out
for abl in none gemm ln self cross sample "gemm,ln,sample,embed" "self,cross"; do
  v=$abl; [ "$abl" = none ] && v=""
  CXRM_ABLATE="$v" timeout 120 python bench.py --steps 2 --warmup 2 --no-cpu-baseline --no-gpu-eager --no-e2e --no-profile 2>/dev/null | \
    python -c "import sys,json; d=json.loads(sys.stdin.read()); print('%-36s ms_per_step %.2f steps %d' % ('$abl', d['ms_per_step'], d['config']['decode_steps_executed']))"
done | tee gpurun_out/ablate.txt
