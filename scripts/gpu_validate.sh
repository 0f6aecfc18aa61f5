#!/bin/bash
# one gpurun call: GPU parity tests, smoke, N=1 bench + reference arm, ncu launch list, one ncu --set full capture,
# command line at scale (phases, configs[1])
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
nproc >> gpurun_out/gpu.txt
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/smoke.log
tail -3 gpurun_out/smoke.log
timeout 600 python bench.py > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "bench rc=$?"
cat gpurun_out/bench_n1.json
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err
cat gpurun_out/bench_ref.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches.csv \
  python bench.py --reads 10000000 --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_launch_bench.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:mcx_build_fused_kernel -s 1 -c 1 -f -o gpurun_out/fused_full \
  python bench.py --reads 10000000 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_full_bench.log 2>&1
bash scripts/gpu_cli_phases.sh > gpurun_out/cli_phases.txt 2>&1; tail -12 gpurun_out/cli_phases.txt
bash scripts/gpu_cli_config2.sh > gpurun_out/cli_config2.txt 2>&1; tail -4 gpurun_out/cli_config2.txt
# two-pass design probe
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o /tmp/part_bench scripts/part_bench.cu && timeout 120 /tmp/part_bench 1200 > gpurun_out/part_bench.txt 2>&1; cat gpurun_out/part_bench.txt
ls -la gpurun_out
