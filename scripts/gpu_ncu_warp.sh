#!/bin/bash
# ncu --set full of the warp kernel (and the fused one for comparison) on a 10M-read launch
set -u
mkdir -p gpurun_out
KBENCH_MD5=0 timeout 900 ncu --set full --clock-control none --import-source on -k regex:mcx_build_warp_kernel -s 2 -c 1 -f -o gpurun_out/r2c_warp_full \
  python scripts/kbench.py 10000000 warp:MCX_KERNEL=warp > gpurun_out/r2c_ncu_warp.log 2>&1
tail -3 gpurun_out/r2c_ncu_warp.log
KBENCH_MD5=0 timeout 900 ncu --set full --clock-control none --import-source on -k regex:mcx_build_fused_kernel -s 2 -c 1 -f -o gpurun_out/r2c_fused_full \
  python scripts/kbench.py 10000000 fused:MCX_KERNEL=fused > gpurun_out/r2c_ncu_fused.log 2>&1
tail -3 gpurun_out/r2c_ncu_fused.log
ls -la gpurun_out/*.ncu-rep
