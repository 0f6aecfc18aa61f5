#!/bin/bash
# N-GPU call: IPC parity test, then the bench at N ranks (with per-stage profile on stderr)
set -u
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv > gpurun_out/r2n_gpus_$N.txt
timeout 900 python -m pytest tests/test_gpu_multi.py -x -q -m gpu > gpurun_out/r2n_pytest_multi_$N.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2n_pytest_multi_$N.log; tail -4 gpurun_out/r2n_pytest_multi_$N.log
MCX_MULTI_PROFILE=1 timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29617 bench.py --gpus $N --steps 3 --warmup 2 > gpurun_out/r2n_bench_n$N.json 2> gpurun_out/r2n_bench_n$N.err; echo "bench rc=$?"
cat gpurun_out/r2n_bench_n$N.json | cut -c1-1800
grep -E "parity|stage ms" gpurun_out/r2n_bench_n$N.err | cut -c1-600 | head -12
