#!/bin/bash
# time alternative builds of libmcxgpu.so (mccortex_b200/lib/libmcxgpu_<name>.so) on the bench workload
set -u
mkdir -p gpurun_out
cp mccortex_b200/lib/libmcxgpu.so /tmp/keep.so
R=${1:-20000000}
KBENCH_MD5=${MD5:-0} python scripts/kbench.py $R base: 2>&1 | tail -1
for f in mccortex_b200/lib/libmcxgpu_*.so; do
  n=$(basename $f .so); n=${n#libmcxgpu_}
  cp $f mccortex_b200/lib/libmcxgpu.so
  KBENCH_MD5=${MD5:-0} python scripts/kbench.py $R $n: 2>&1 | tail -1
done
cp /tmp/keep.so mccortex_b200/lib/libmcxgpu.so
