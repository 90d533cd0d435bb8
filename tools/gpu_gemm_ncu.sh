#!/bin/bash
# ncu durations + tensor-pipe activity of every encoder GEMM shape (tools/bench_gemm.py, one launch per shape)
# Usage: bash tools/gpu_gemm_ncu.sh <tag> [env assignments...]
tag=$1; shift
mkdir -p gpurun_out
env "$@" WARM=1 ITERS=1 REPS=1 timeout 600 ncu --metrics gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_tensor.sum,sm__throughput.avg.pct_of_peak_sustained_elapsed,lts__t_bytes.sum,dram__bytes_read.sum,dram__bytes_write.sum \
  --clock-control none -k regex:gemm_tc --csv --log-file gpurun_out/gemm_ncu_$tag.csv python tools/bench_gemm.py > gpurun_out/gemm_ncu_$tag.log 2>&1
echo "rc $?"
python - <<PY
import csv, collections
rows=[l for l in open("gpurun_out/gemm_ncu_$tag.csv") if not l.startswith("==")]
rd=csv.DictReader(rows)
per=collections.OrderedDict()
for r in rd:
    k=r["ID"]
    per.setdefault(k,{"name":r["Kernel Name"][:60],"grid":r["Grid Size"]})[r["Metric Name"]]=r["Metric Value"]
for k,v in per.items():
    print(k, v["name"], v["grid"], v.get("gpu__time_duration.sum"), "tensor%", v.get("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"), "lts MB", float(v.get("lts__t_bytes.sum","0").replace(",",""))/1e6 if v.get("lts__t_bytes.sum") else None, "dram rd/wr MB", v.get("dram__bytes_read.sum"), v.get("dram__bytes_write.sum"))
PY
