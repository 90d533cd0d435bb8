#!/bin/bash
# Separate the decode attention kernels' memory side from their consumer math: CXRM_ATTN_NOCOMPUTE keeps the TMA
# pipeline and skips the MMAs; prefetch mode 2 (real loads, serialised in profile mode) makes the K/V L2-resident.
mkdir -p gpurun_out
for v in "0 0 2 0" "0 0 2 1" "0.5 1 2 1" "0.5 1 2 0"; do
  set -- $v
  if [ "$4" = 1 ]; then export CXRM_ATTN_NOCOMPUTE=1; else unset CXRM_ATTN_NOCOMPUTE; fi
  CXRM_PF_MODE=$3 CXRM_PF_CROSS=$1 CXRM_PF_SELF=$2 timeout 300 python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --profile-out gpurun_out/pf_probe.json > /dev/null 2>&1
  python - "$1" "$2/mode$3/nocompute$4" <<'PY'
import json, sys
d = json.load(open("gpurun_out/pf_probe.json"))["breakdown"]
f = lambda k: (1000 * d[k]["ms"] / d[k]["n"], d[k]["n"]) if k in d else (0, 0)
print("cross %s self %s: prefetch %.1f us x%d  cross_attn %.1f us  self_attn %.1f us" % (
    sys.argv[1], sys.argv[2], *f("decode.prefetch"), f("decode.cross_attn")[0], f("decode.self_attn")[0]))
PY
done | tee gpurun_out/pf_probe2.txt
