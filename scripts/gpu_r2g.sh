#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/r2g_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2g_pytest_gpu.log
tail -6 gpurun_out/r2g_pytest_gpu.log
KBENCH_MD5=1 python scripts/kbench.py 50000000 fused: > gpurun_out/r2g_kbench_50M.txt 2>&1; cat gpurun_out/r2g_kbench_50M.txt | tail -2
KBENCH_MD5=0 python scripts/kbench.py 20000000 fused: > gpurun_out/r2g_kbench_20M.txt 2>&1; cat gpurun_out/r2g_kbench_20M.txt | tail -2
cp mccortex_b200/lib/libmcxgpu.so /tmp/keep.so; cp mccortex_b200/lib/libmcxgpu_minb4.so mccortex_b200/lib/libmcxgpu.so
KBENCH_MD5=0 python scripts/kbench.py 20000000 fused_minb4: > gpurun_out/r2g_kbench_20M_minb4.txt 2>&1; cat gpurun_out/r2g_kbench_20M_minb4.txt | tail -2
cp /tmp/keep.so mccortex_b200/lib/libmcxgpu.so
KBENCH_K=63 KBENCH_MD5=0 python scripts/kbench.py 20000000 k63: > gpurun_out/r2g_kbench_20M_k63.txt 2>&1; cat gpurun_out/r2g_kbench_20M_k63.txt | tail -2
