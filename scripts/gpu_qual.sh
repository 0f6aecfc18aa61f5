#!/bin/bash
# quality cut-off path: kernel-only rate and the per-kernel split (ncu launch list)
set -u
mkdir -p gpurun_out
TAG=${1:-r2q}
KBENCH_QUAL=0.01 KBENCH_MD5=1 python scripts/kbench.py 50000000 qual: 2>&1 | tail -1
KBENCH_QUAL=0.01 KBENCH_MD5=0 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/${TAG}_qual_launches.csv \
  python scripts/kbench.py 20000000 qual: > gpurun_out/${TAG}_qual_ncu.log 2>&1
python3 - <<PY
import csv,collections
t=collections.defaultdict(lambda:[0,0.0])
for r in csv.DictReader(l for l in open("gpurun_out/${TAG}_qual_launches.csv") if not l.startswith("==")):
    n=r["Kernel Name"].split("(")[0]; t[n][0]+=1; t[n][1]+=float(r["Metric Value"].replace(",",""))/1e6
for n,(c,ms) in sorted(t.items(), key=lambda x:-x[1][1]): print("%-60s n=%3d %9.2f ms"%(n[:60],c,ms))
PY
