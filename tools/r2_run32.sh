#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 ncu --clock-control none --metrics gpu__time_duration.sum --csv --log-file gpurun_out/r2_cfg4_launches.csv python tools/bench_configs.py cfg4 32 1500 > gpurun_out/r2_cfg32.log 2>&1; tail -2 gpurun_out/r2_cfg32.log
python - <<'PY'
import csv, re, collections
rows=list(csv.reader(open("gpurun_out/r2_cfg4_launches.csv")))
hi=[i for i,r in enumerate(rows) if r and r[0]=="ID"][0]
h=rows[hi]; data=rows[hi+1:]
ki=h.index("Kernel Name"); vi=h.index("Metric Value"); ui=h.index("Metric Unit")
opt=[i for i,r in enumerate(data) if "optimizer" in r[ki]]
print("launches", len(data), "optimizer at", opt)
step=data[opt[0]+1:opt[1]+1] if len(opt)>1 else data
agg={}
tot=0
for r in step:
    v=float(r[vi].replace(",","")); unit=r[ui]
    ms = v/1e6 if unit.startswith("ns") else (v/1e3 if unit.startswith("us") else v)
    key=re.sub(r"\(.*","",r[ki]).replace("void ","").replace("b2::","")[:60]
    a=agg.setdefault(key,[0,0.0]); a[0]+=1; a[1]+=ms; tot+=ms
print("step launches", len(step), "total ms", round(tot,1))
for k,(c,ms) in sorted(agg.items(), key=lambda kv:-kv[1][1])[:25]:
    print("%-62s %5d %9.2f ms" % (k,c,ms))
PY
