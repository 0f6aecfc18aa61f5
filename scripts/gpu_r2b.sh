#!/bin/bash
set -u
mkdir -p gpurun_out
python scripts/wdebug.py > gpurun_out/r2b_wdebug.txt 2>&1; cat gpurun_out/r2b_wdebug.txt
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_abi.py -x -q -m gpu > gpurun_out/r2b_pytest_parity.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2b_pytest_parity.log
tail -5 gpurun_out/r2b_pytest_parity.log
timeout 900 python scripts/kbench.py 20000000 > gpurun_out/r2b_kbench_20M.txt 2>&1; cat gpurun_out/r2b_kbench_20M.txt | tail -8
timeout 600 python scripts/kbench.py 50000000 fused:MCX_KERNEL=fused warp:MCX_KERNEL=warp warp2x20:MCX_KERNEL=warp,MCX_CLASSES=2,MCX_FRONT_BITS=20 > gpurun_out/r2b_kbench_50M.txt 2>&1; tail -4 gpurun_out/r2b_kbench_50M.txt
