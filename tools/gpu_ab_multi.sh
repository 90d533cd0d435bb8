#!/bin/bash
# Same-box comparison of several build snapshots (directories given as arguments), alternating, 2 rounds.
mkdir -p gpurun_out
for i in 1 2; do
  for d in "$@"; do
    (cd $d && timeout 300 python bench.py --steps 3 --warmup 2 --no-cpu-baseline --no-profile --no-e2e 2>&1 | tail -1 | \
      python -c "import sys,json; d=json.loads(sys.stdin.read()); print('%-10s ms_per_step %.2f reports/s %.1f' % ('$d', d['ms_per_step'], d['value']))")
  done
done | tee gpurun_out/ab_multi.txt
