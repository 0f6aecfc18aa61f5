#!/bin/bash
# the cold-table, exchange-bound regime of configs[4] on N GPUs: R reads per GPU of a 3 Gbp genome (600 M reads over 8 GPUs = 75 M each)
set -u
N=${1:-8}; R=${2:-37500000}
mkdir -p gpurun_out
MCX_BENCH_GENOME=3000000000 MCX_MULTI_BIN_FRAC=1.4 MCX_MULTI_BATCH_READS=2000000 MCX_MULTI_PROFILE=1 timeout ${INNER_TIMEOUT:-1200} python -m torch.distributed.run --nnodes=1 --nproc-per-node $N \
  --master-addr 127.0.0.1 --master-port 29620 bench.py --gpus $N --reads $R --steps 2 --warmup 1 > gpurun_out/r2y_config5_n$N.json 2> gpurun_out/r2y_config5_n$N.err; echo "bench rc=$?"
python3 - <<PY
import json
d=[json.loads(l) for l in open("gpurun_out/r2y_config5_n$N.json") if l.startswith("{")][-1]
r=d["roofline"]
print("config5-shaped N=%d: %d reads/GPU: value %.2f G k-mers/s  %.1f ms/step  frac %.3f  e2e %.2f G/s  tuples/step %.0fM  nvlink %.1f GB/step  distinct %d  parity %s" % (
  d["n_gpus"], d["config"]["reads_per_gpu"], d["value"]/1e9, d["ms_per_step"], r["frac"], d["e2e"]["value"]/1e9, r["tuples_per_step"]/1e6, r["nvlink_bytes_per_step"]/1e9, d["extra"]["distinct_kmers_total"], (d.get("parity") or {}).get("ok")))
PY
grep -E "stage ms" gpurun_out/r2y_config5_n$N.err | sed 's/; tuples sent.*//' | cut -c1-260 | head -3
grep -E "parity|Error|error" gpurun_out/r2y_config5_n$N.err | head -5
