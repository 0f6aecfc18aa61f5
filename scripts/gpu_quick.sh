#!/bin/bash
# parity tests + a short N=1 bench (20M reads) + optional env sweeps: scripts/gpu_quick.sh [label]
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -4
run() { python bench.py --reads ${READS:-20000000} --steps 3 --warmup 2 --no-cpu-baseline 2>&1 | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('value %.2f G/s  kernel_ms %.1f e2e %.2f G/s'%(d['value']/1e9, d['roofline']['kernel_ms'], d['e2e']['value']/1e9))"; }
echo "== default"; run
echo "== G=1"; MCX_G=1 run
echo "== ceiling (1 Mbp genome, no errors)"; MCX_BENCH_GENOME=1000000 MCX_BENCH_PERR=0 run
