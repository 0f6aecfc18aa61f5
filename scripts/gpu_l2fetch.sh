#!/bin/bash
mkdir -p gpurun_out
for g in "" 32 64 128; do
  echo "== L2FETCH=$g"
  L2FETCH=$g timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:mcx_insert_tuples --csv --log-file gpurun_out/r2m_l2fetch_$g.csv python scripts/insert_bench.py 30 64 2>&1 | grep -E "cudaLimit|G inserts" | cut -c1-150
  grep -E "dram__bytes|gpu__time" gpurun_out/r2m_l2fetch_$g.csv | awk -F'","' '{print $NF, $(NF-2)}' | tr -d '"' | paste - - - | head -3
done
