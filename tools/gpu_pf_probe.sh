#!/bin/bash
# Does L2 residency speed the decode attention kernels up?  Event-profiled step (prefetch serialised on the main
# stream) for several prefetch fractions: mean us of the prefetch, cross- and self-attention launches.
mkdir -p gpurun_out
for v in "0 0 0" "1 1 0" "1 1 1" "1 1 2" "0.5 1 1" "0.5 1 2" "0.25 0.5 2"; do
  set -- $v
  CXRM_PF_MODE=$3 CXRM_PF_CROSS=$1 CXRM_PF_SELF=$2 timeout 300 python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --profile-out gpurun_out/pf_probe.json > /dev/null 2>&1
  python - "$1" "$2/mode$3" <<'PY'
import json, sys
d = json.load(open("gpurun_out/pf_probe.json"))["breakdown"]
f = lambda k: (1000 * d[k]["ms"] / d[k]["n"], d[k]["n"]) if k in d else (0, 0)
print("cross %s self %s: prefetch %.1f us x%d  cross_attn %.1f us  self_attn %.1f us  gemm768 %.1f us  ln %.1f us" % (
    sys.argv[1], sys.argv[2], *f("decode.prefetch"), f("decode.cross_attn")[0], f("decode.self_attn")[0],
    f("decode.gemm[768x768]")[0], f("decode.layernorm")[0]))
PY
done | tee gpurun_out/pf_probe.txt
