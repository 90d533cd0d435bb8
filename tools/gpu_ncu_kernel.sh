#!/bin/bash
# one `ncu --set full` capture of a kernel (regex) inside the decode loop + the hottest source lines by sampled stalls
# Usage: bash tools/gpu_ncu_kernel.sh <kernel-regex> <tag> [skip]
k=$1; tag=$2; skip=${3:-8}
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:$k -s $skip -c 1 -f -o gpurun_out/k_$tag \
  python tools/trace_step.py 16 > gpurun_out/k_$tag.log 2>&1
echo "ncu rc $?"
ncu -i gpurun_out/k_$tag.ncu-rep --page details --csv 2>/dev/null | python -c "
import csv,sys
for r in csv.DictReader(sys.stdin):
    n=r['Metric Name']
    if n in ('Duration','Executed Ipc Active','Issue Slots Busy','No Eligible','Registers Per Thread','Theoretical Occupancy','Achieved Occupancy','L1/TEX Hit Rate','Avg. Active Threads Per Warp','One or More Eligible','Warp Cycles Per Issued Instruction','Branch Efficiency','Dynamic Shared Memory Per Block','Static Shared Memory Per Block'):
        print(f\"{n}: {r['Metric Value']} {r['Metric Unit']}\")
"
ncu -i gpurun_out/k_$tag.ncu-rep --page source --csv 2>/dev/null > gpurun_out/k_${tag}_source.csv
python - <<PY
import csv
rows=list(csv.DictReader(open("gpurun_out/k_${tag}_source.csv")))
print(len(rows), "source rows; columns:", [c for c in rows[0].keys()][:12] if rows else None)
key=[c for c in rows[0].keys() if c.startswith("# Samples") or c=="Warp Stall Sampling (All Samples)"]
kk=key[0] if key else None
def val(r):
    try: return float(r[kk].replace(",",""))
    except: return 0.0
tot=sum(val(r) for r in rows)
print("total samples", tot, "column", kk)
top=sorted(rows,key=val,reverse=True)[:40]
for r in top:
    print(f"{val(r):8.0f} {100*val(r)/max(tot,1):5.1f}%  {r.get('Source','')[:130]}")
PY
rm -f gpurun_out/k_$tag.ncu-rep
