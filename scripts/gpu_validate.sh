#!/bin/bash
# one gpurun call: the whole GPU suite (incl. full-size md5 parity), smoke, N=1 bench + reference arm, ncu launch list
set -u
mkdir -p gpurun_out
TAG=${1:-r2}
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/${TAG}_gpu.txt 2>&1; nproc >> gpurun_out/${TAG}_gpu.txt
timeout 2400 python -m pytest tests -x -q -m gpu > gpurun_out/${TAG}_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${TAG}_pytest_gpu.log
tail -5 gpurun_out/${TAG}_pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${TAG}_smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/${TAG}_smoke.log
tail -3 gpurun_out/${TAG}_smoke.log
timeout 900 python bench.py > gpurun_out/${TAG}_bench_n1.json 2> gpurun_out/${TAG}_bench_n1.err; echo "bench rc=$?"
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/${TAG}_bench_ref.json 2> gpurun_out/${TAG}_bench_ref.err
python3 - <<PY
import json
d=[json.loads(l) for l in open("gpurun_out/${TAG}_bench_n1.json") if l.startswith("{")][-1]
r=[json.loads(l) for l in open("gpurun_out/${TAG}_bench_ref.json") if l.startswith("{")][-1]
print("value %.2f G/s  ms/step %.1f  frac %.3f  e2e %.2f G/s (%.1f ms)  reference %.1f M/s  e2e ratio %.0fx" % (d["value"]/1e9, d["ms_per_step"], d["roofline"]["frac"], d["e2e"]["value"]/1e9, d["e2e"]["ms_per_step"], r["value"]/1e6, d["e2e"]["value"]/r["value"]))
x=d["extra"]
print("cli", x.get("cli")); 
for k,v in (x.get("other_configs") or {}).items(): print(k, {a:(round(b/1e9,2) if a in ("value","positions_per_s") else b) for a,b in v.items() if a in ("value","ms_per_step","roofline_frac","failed")})
c=x.get("config5") or {}
print("config5", {a:c.get(a) for a in ("value","ms_per_step","frac","failed")}, (c.get("insert_kernel_cold") or {}).get("novel",{}).get("inserts_per_s"), (c.get("insert_kernel_cold") or {}).get("found",{}).get("inserts_per_s"))
PY
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/${TAG}_launches.csv \
  python bench.py --reads 10000000 --steps 2 --warmup 1 --no-cpu-baseline --no-extras > gpurun_out/${TAG}_ncu_launch_bench.log 2>&1
grep -c "mcx_" gpurun_out/${TAG}_launches.csv
