#!/bin/bash
# k=63: hot-pass depth (libmcxgpu_<name>.so variants) x front-table size
set -u
mkdir -p gpurun_out
cp mccortex_b200/lib/libmcxgpu.so /tmp/keep.so
run() { KBENCH_K=63 KBENCH_MD5=0 python scripts/kbench.py ${R:-50000000} $1_b21:MCX_FRONT_BITS=21 $1_b22:MCX_FRONT_BITS=22 2>&1 | tail -2; }
run base
for f in mccortex_b200/lib/libmcxgpu_*.so; do
  n=$(basename $f .so); n=${n#libmcxgpu_}
  cp $f mccortex_b200/lib/libmcxgpu.so
  run $n
done
cp /tmp/keep.so mccortex_b200/lib/libmcxgpu.so
