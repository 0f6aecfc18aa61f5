#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -4
run() { python bench.py --reads ${READS:-20000000} --steps 3 --warmup 2 --no-cpu-baseline 2>&1 | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('value %.2f G/s  kernel_ms %.1f e2e %.2f G/s'%(d['value']/1e9, d['roofline']['kernel_ms'], d['e2e']['value']/1e9))"; }
for mb in 3 4 5; do for gg in 1 2; do echo "== minb $mb G $gg"; MCX_MINB=$mb MCX_G=$gg run; done; done
echo "== ceiling G=1"; MCX_G=1 MCX_BENCH_GENOME=1000000 MCX_BENCH_PERR=0 run
echo "== perr 0 genome 4.6M G=1"; MCX_G=1 MCX_BENCH_PERR=0 run
MCX_G=1 timeout 900 ncu --set full --clock-control none --import-source on -k regex:mcx_build_fused_kernel -s 1 -c 1 -f -o gpurun_out/fused_v3 python bench.py --reads 10000000 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_v3.log 2>&1; tail -1 gpurun_out/ncu_v3.log | cut -c1-200
