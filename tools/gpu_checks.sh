#!/bin/bash
# Run the GPU test files one process each (a CUDA fault in one file must not poison the rest), with hard timeouts.
# Usage (on the GPU box, via gpurun):  bash tools/gpu_checks.sh [pytest -k expression]
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.sm,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
for f in tests/test_kernels_gpu.py tests/test_parity_gpu.py; do
  name=$(basename $f .py)
  echo "=== $f" | tee -a gpurun_out/summary.txt
  timeout 900 python -m pytest $f -m gpu -q -rA --no-header -p no:cacheprovider ${1:+-k "$1"} > gpurun_out/$name.log 2>&1
  echo "exit $?" | tee -a gpurun_out/summary.txt
  grep -E "^(PASSED|FAILED|ERROR)|passed|failed|error" gpurun_out/$name.log | tail -80 | tee -a gpurun_out/summary.txt
done
