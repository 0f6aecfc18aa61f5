#!/bin/bash
# last gpurun call of round 1: the whole GPU suite (no -x: list every failure), smoke, N=1 bench
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
nproc >> gpurun_out/gpu.txt
timeout 420 python -m pytest tests -q -m gpu -k "pcr" > gpurun_out/pytest_gpu_pcr.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu_pcr.log
tail -5 gpurun_out/pytest_gpu_pcr.log
timeout 600 python -m pytest tests -q -m gpu -k "not pcr" > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/smoke.log
tail -3 gpurun_out/smoke.log
timeout 400 python bench.py > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "bench rc=$?"
cat gpurun_out/bench_n1.json
