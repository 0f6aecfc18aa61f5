#!/bin/bash
# final N-GPU record: (N=2 only) IPC parity tests + shard-mode CLI tests, then the bench at N ranks as the driver launches it
set -u
N=${1:-2}; TAG=${2:-r2fin}
mkdir -p gpurun_out
if [ "$N" = 2 ]; then
  timeout 1200 python -m pytest tests/test_gpu_multi.py tests/test_zy_gpu_cli_new.py -x -q -m gpu -k "routed or shard" > gpurun_out/${TAG}_pytest_multi.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${TAG}_pytest_multi.log; tail -3 gpurun_out/${TAG}_pytest_multi.log
fi
MCX_MULTI_PROFILE=1 timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29621 bench.py --gpus $N --steps 3 --warmup 3 > gpurun_out/${TAG}_bench_n$N.json 2> gpurun_out/${TAG}_bench_n$N.err; echo "bench rc=$?"
python3 - <<PY
import json
d=[json.loads(l) for l in open("gpurun_out/${TAG}_bench_n$N.json") if l.startswith("{")][-1]
print("N=%d value %.1f G/s  %.1f ms/step  frac %.3f  e2e %.1f G/s (%.1f ms)  tuples/step %.0fM  parity %s" % (d["n_gpus"], d["value"]/1e9, d["ms_per_step"], d["roofline"]["frac"], d["e2e"]["value"]/1e9, d["e2e"]["ms_per_step"], d["roofline"].get("tuples_per_step",0)/1e6, (d.get("parity") or {}).get("ok")))
PY
grep -E "stage ms" gpurun_out/${TAG}_bench_n$N.err | sed 's/; tuples sent.*//' | cut -c1-260 | head -3
