#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
B2_WIDE_ONLY=1 timeout 600 python tools/bench_wide.py > gpurun_out/r2_wide38.log 2>&1; cat gpurun_out/r2_wide38.log
B2_WIDE_ONLY=1 WIDE_T=300 timeout 600 ncu --clock-control none --metrics gpu__time_duration.sum --csv --log-file gpurun_out/r2_wide38_launches.csv python tools/bench_wide.py > /dev/null 2>&1
python - <<'PY'
import csv, re
rows=list(csv.reader(open("gpurun_out/r2_wide38_launches.csv")))
hi=[i for i,r in enumerate(rows) if r and r[0]=="ID"][0]
h=rows[hi]; data=rows[hi+1:]
ki=h.index("Kernel Name"); vi=h.index("Metric Value"); ui=h.index("Metric Unit")
last=data[-40:]
for r in last:
    v=float(r[vi].replace(",","")); unit=r[ui]
    ms = v/1e6 if unit.startswith("ns") else (v/1e3 if unit.startswith("us") else v)
    print("%-70s %8.3f ms" % (re.sub(r"\(.*","",r[ki])[:70], ms))
PY
