#!/bin/bash
set -u
N=${1:-2}; R=${2:-20000000}
mkdir -p gpurun_out
run() { # name, env...
  name=$1; shift
  env "$@" MCX_MULTI_PROFILE=1 MCX_MULTI_PARITY=0 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29618 bench.py --gpus $N --reads $R --steps 2 --warmup 1 > gpurun_out/r2o_${name}_n$N.json 2> gpurun_out/r2o_${name}_n$N.err
  python - <<PY
import json
d=json.load(open("gpurun_out/r2o_${name}_n$N.json"))
print("%-14s N=%d value %.1f G/s  %.1f ms/step  e2e %.1f G/s  tuples/step %.0fM" % ("$name", d["n_gpus"], d["value"]/1e9, d["ms_per_step"], d["e2e"]["value"]/1e9, d["roofline"].get("tuples_per_step",0)/1e6))
PY
  grep "rank 0 stage ms" gpurun_out/r2o_${name}_n$N.err | sed 's/; tuples sent.*//' | cut -c1-300
}
run routed_overlap MCX_MULTI_OVERLAP=1
run routed_serial MCX_MULTI_OVERLAP=0
run nccl_bins MCX_MULTI_EXCHANGE=nccl
timeout 600 python -m pytest tests/test_gpu_multi.py -x -q -m gpu 2>&1 | tail -2
