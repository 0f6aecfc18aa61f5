#!/bin/bash
# k=63 build kernel: front table on / off (same records), the k=31 kernel beside it, then the C-ABI parity tests
set -u
mkdir -p gpurun_out
R=${1:-50000000}
KBENCH_K=63 KBENCH_MD5=1 python scripts/kbench.py $R front: nofront:MCX_FRONT_BITS=0 2>&1 | tail -2
KBENCH_K=47 KBENCH_MD5=1 python scripts/kbench.py 20000000 front: nofront:MCX_FRONT_BITS=0 2>&1 | tail -2
KBENCH_K=31 KBENCH_MD5=1 python scripts/kbench.py $R front: 2>&1 | tail -1
if [ "${2:-}" = tests ]; then python -m pytest tests/test_gpu_parity.py -x -q -m gpu 2>&1 | tail -5; fi
