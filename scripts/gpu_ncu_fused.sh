#!/bin/bash
# ncu of the build kernel on a 10M-read launch of the bench workload: --set full capture + launch list
set -u
mkdir -p gpurun_out
TAG=${1:-r2}
KBENCH_MD5=0 timeout 900 ncu --set full --clock-control none --import-source on -k regex:mcx_build_fused_kernel -s 2 -c 1 -f -o gpurun_out/${TAG}_fused_full \
  python scripts/kbench.py 10000000 fused: > gpurun_out/${TAG}_ncu_fused.log 2>&1
tail -2 gpurun_out/${TAG}_ncu_fused.log
