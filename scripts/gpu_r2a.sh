#!/bin/bash
# round 2, call A: parity of the warp kernel, kernel A/B on the bench workload, cold insert kernel, partition probe
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1; nproc >> gpurun_out/gpu.txt
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_abi.py -x -q -m gpu > gpurun_out/r2a_pytest_parity.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2a_pytest_parity.log
tail -5 gpurun_out/r2a_pytest_parity.log
timeout 900 python scripts/kbench.py 20000000 > gpurun_out/r2a_kbench_20M.txt 2>&1; cat gpurun_out/r2a_kbench_20M.txt | tail -8
timeout 600 python scripts/kbench.py 50000000 fused:MCX_KERNEL=fused warp:MCX_KERNEL=warp warp2x20:MCX_KERNEL=warp,MCX_CLASSES=2,MCX_FRONT_BITS=20 > gpurun_out/r2a_kbench_50M.txt 2>&1; tail -4 gpurun_out/r2a_kbench_50M.txt
timeout 300 python scripts/insert_bench.py 30 256 > gpurun_out/r2a_insert_bench.txt 2>&1; cat gpurun_out/r2a_insert_bench.txt | tail -4
timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct --clock-control none -k regex:mcx_insert_tuples --csv --log-file gpurun_out/r2a_insert_ncu.csv python scripts/insert_bench.py 30 64 > gpurun_out/r2a_insert_ncu.log 2>&1; tail -4 gpurun_out/r2a_insert_ncu.csv
timeout 120 ./scripts/part_bench 1200 > gpurun_out/r2a_part_bench.txt 2>&1; cat gpurun_out/r2a_part_bench.txt
ls -la gpurun_out | tail -12
