#!/bin/bash
# Same-box A/B: a reference build (snapshot in _ab/, not tracked) against the working tree, alternating runs.
# Usage: bash tools/gpu_ab.sh [rounds] ; extra settings for the working-tree runs via AB_ENV="K=V K=V"
mkdir -p gpurun_out
run() {  # label, dir, env...
  local label=$1 dir=$2; shift 2
  (cd $dir && env "$@" timeout 300 python bench.py --steps 3 --warmup 2 --no-cpu-baseline --no-profile --no-e2e 2>&1 | tail -1 | \
    python -c "import sys,json; d=json.loads(sys.stdin.read()); print('%-14s ms_per_step %.2f reports/s %.1f launches %d sm_mhz %s' % ('$label', d['ms_per_step'], d['value'], d['gpu_launches'], d['clocks']['sm_mhz']))")
}
for i in $(seq 1 ${1:-2}); do
  run ref _ab A=1
  run new . A=1 $AB_ENV
  if [ -n "$AB_ALT" ]; then run new-alt . $AB_ALT; fi
done | tee gpurun_out/ab.txt
