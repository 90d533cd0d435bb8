#!/bin/bash
# ncu launch list of ONE encoder pass (100 images): per-kernel time and DRAM bytes, summarised over the last pass
# Usage: bash tools/gpu_enc_ncu.sh <tag> [env assignments...]
tag=$1; shift
mkdir -p gpurun_out
env "$@" ENC_REPS=1 timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active --clock-control none \
  --csv --log-file gpurun_out/enclist_$tag.csv python tools/enc_time.py > gpurun_out/enclist_$tag.log 2>&1
echo "rc $?"
python - <<PY
import csv, collections, re
rows=[l for l in open("gpurun_out/enclist_$tag.csv") if not l.startswith("==")]
per=collections.OrderedDict()
for r in csv.DictReader(rows):
    d=per.setdefault(r["ID"],{"name":r["Kernel Name"],"grid":r["Grid Size"]})
    d[r["Metric Name"]]=(float(r["Metric Value"].replace(",","")), r["Metric Unit"])
ks=list(per.values())
# the last encoder pass = the launches after the last-but-one im2col_pixels
idx=[i for i,k in enumerate(ks) if "im2col_pixels" in k["name"]]
n_pass=2
start=idx[-n_pass] if len(idx)>=n_pass else 0
ks=ks[start:]
def us(k):
    v,u=k["gpu__time_duration.sum"]; return v/1e3 if u.startswith("ns") else v
def mb(k,m):
    v,u=k.get(m,(0,"byte")); f={"byte":1e-6,"Kbyte":1e-3,"Mbyte":1,"Gbyte":1e3}.get(u,1e-6); return v*f
agg=collections.OrderedDict()
for k in ks:
    nm=re.sub(r"^void ","",k["name"]); nm=re.sub(r"\(.*","",nm); nm=nm.replace("cxrm::<unnamed>::","").replace("cxrm::","")
    a=agg.setdefault(nm,[0,0.0,0.0,0.0,0.0]); a[0]+=1; a[1]+=us(k); a[2]+=mb(k,"dram__bytes_read.sum"); a[3]+=mb(k,"dram__bytes_write.sum")
    a[4]+=us(k)*k.get("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",(0,"%"))[0]
tot=sum(a[1] for a in agg.values())
out=["# ncu launch list of one encoder pass (100 images, cold-cache, serialised): %d launches, %.2f ms\n"%(len(ks),tot/1e3),
     "| kernel | launches | total us | mean us | DRAM rd MB | DRAM wr MB | GB/s | tensor pipe % (time-weighted) |","|---|---:|---:|---:|---:|---:|---:|---:|"]
for nm,a in sorted(agg.items(), key=lambda kv:-kv[1][1]):
    out.append("| \`%s\` | %d | %.0f | %.1f | %.0f | %.0f | %.0f | %.1f |"%(nm[:70],a[0],a[1],a[1]/a[0],a[2],a[3],(a[2]+a[3])/a[1]*1e3 if a[1] else 0,a[4]/a[1] if a[1] else 0))
open("gpurun_out/enclist_$tag.md","w").write("\n".join(out)+"\n")
print("\n".join(out))
PY
gzip -f gpurun_out/enclist_$tag.csv
