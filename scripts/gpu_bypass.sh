#!/bin/bash
set -u
cp mccortex_b200/lib/libmcxgpu.so /tmp/keep.so
for lib in /tmp/keep.so mccortex_b200/lib/libmcxgpu_nobypass.so; do
  cp $lib mccortex_b200/lib/libmcxgpu.so 2>/dev/null
  n=$(basename $lib .so)
  KBENCH_MD5=1 python scripts/kbench.py 20000000 $n-config2: 2>&1 | tail -1 | cut -c1-200
  MCX_BENCH_GENOME=3000000000 KBENCH_MD5=1 python scripts/kbench.py 10000000 $n-cold3G: 2>&1 | tail -1 | cut -c1-200
  MCX_BENCH_GENOME=100000000 KBENCH_MD5=1 python scripts/kbench.py 20000000 $n-100Mbp: 2>&1 | tail -1 | cut -c1-200
done
cp /tmp/keep.so mccortex_b200/lib/libmcxgpu.so
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu 2>&1 | tail -2
