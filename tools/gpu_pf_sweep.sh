#!/bin/bash
# Sweep of the decode-step L2 prefetch (CXRM_PF_CROSS / CXRM_PF_SELF fractions, CXRM_PF_CTAS grid): ms per SCST step.
mkdir -p gpurun_out
for v in "0 0 148" "0.5 1 148" "0.5 0 148" "0 1 148" "0.3 0.6 148" "0.7 1 148" "1 1 148" "0.5 1 74" "0.5 1 296" "0.5 1 32"; do
  set -- $v
  CXRM_PF_CROSS=$1 CXRM_PF_SELF=$2 CXRM_PF_CTAS=$3 timeout 300 python bench.py --steps 3 --warmup 2 --no-cpu-baseline --no-e2e --no-profile 2>&1 | tail -1 | \
    python -c "import sys,json; d=json.loads(sys.stdin.read()); print('cross $1 self $2 ctas $3  ms_per_step %.2f reports/s %.1f' % (d['ms_per_step'], d['value']))"
done | tee gpurun_out/pf_sweep.txt
