#!/bin/bash
# N-GPU bench (N = $1, default 2): aggregated exchange at the full per-GPU workload, and the unaggregated baseline on a smaller one
set -u
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/topo_$N.txt 2>&1
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus $N --steps 3 --warmup 3 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err; echo "rc=$?"
cat gpurun_out/bench_n$N.json; tail -5 gpurun_out/bench_n$N.err
MCX_MULTI_AGGREGATE=0 timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29532 bench.py --gpus $N --steps 2 --warmup 1 --reads 10000000 > gpurun_out/bench_n${N}_unaggregated.json 2> gpurun_out/bench_n${N}_unaggregated.err; echo "rc=$?"
cat gpurun_out/bench_n${N}_unaggregated.json; tail -5 gpurun_out/bench_n${N}_unaggregated.err
