#!/bin/bash
# per-kernel durations of the MCX_SPILL variant (launch list; times are serialised / cold-cache)
set -u
mkdir -p gpurun_out
MCX_SPILL=1 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/launches_spill.csv \
  python bench.py --reads 10000000 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_spill_bench.log 2>&1
python - <<'PY'
import csv, collections
rows = [r for r in csv.reader(open("gpurun_out/launches_spill.csv")) if len(r) > 10]
hdr = rows[0]; ki = hdr.index("Kernel Name"); vi = hdr.index("Metric Value"); ui = hdr.index("Metric Unit")
for r in rows[1:]:
    print(r[ki][:60], r[vi], r[ui])
PY
