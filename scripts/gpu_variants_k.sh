#!/bin/bash
set -u
cp mccortex_b200/lib/libmcxgpu.so /tmp/keep.so
for K in 31 63; do
  KBENCH_K=$K KBENCH_MD5=1 python scripts/kbench.py 20000000 base: 2>&1 | tail -1
  for f in mccortex_b200/lib/libmcxgpu_*.so; do
    n=$(basename $f .so); n=${n#libmcxgpu_}
    cp $f mccortex_b200/lib/libmcxgpu.so
    KBENCH_K=$K KBENCH_MD5=1 python scripts/kbench.py 20000000 $n: 2>&1 | tail -1
    cp /tmp/keep.so mccortex_b200/lib/libmcxgpu.so
  done
done
