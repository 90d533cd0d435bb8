#!/bin/bash
# quick same-box comparison of environment-variable variants: bash tools/gpu_quick_bench.sh "VAR=1 VAR2=2" "..." ...
# prints value, ms/step and phase times per variant ("" = defaults)
for v in "$@"; do
  out=$(env $v python bench.py --steps 3 --warmup 2 --no-cpu-baseline --no-gpu-eager --no-e2e --no-profile 2>/dev/null | tail -1)
  python - "$v" "$out" <<'PY'
import json, sys
d = json.loads(sys.argv[2])
print(f"[{sys.argv[1] or 'default'}] {d['value']:.2f} reports/s  {d['ms_per_step']:.2f} ms/step  phases {d['phase_ms']}")
PY
done
