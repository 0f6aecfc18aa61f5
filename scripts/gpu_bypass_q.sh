#!/bin/bash
cp mccortex_b200/lib/libmcxgpu.so /tmp/keep.so
for lib in /tmp/keep.so mccortex_b200/lib/libmcxgpu_nobypass.so /tmp/keep.so mccortex_b200/lib/libmcxgpu_nobypass.so; do
  cp $lib mccortex_b200/lib/libmcxgpu.so 2>/dev/null
  KBENCH_MD5=0 python scripts/kbench.py 50000000 $(basename $lib .so)-config2: 2>&1 | tail -1 | cut -c1-120
done
cp /tmp/keep.so mccortex_b200/lib/libmcxgpu.so
