run() { python bench.py --reads 20000000 --steps 2 --warmup 1 --no-cpu-baseline 2>&1 | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('value %.2f G/s  kernel_ms %.1f e2e %.2f G/s'%(d['value']/1e9, d['roofline']['kernel_ms'], d['e2e']['value']/1e9))"; }
for mb in 4 5; do echo "== v2 ceiling minb $mb"; MCX_MINB=$mb MCX_BENCH_GENOME=1000000 MCX_BENCH_PERR=0 run; done
for mb in 4 5 6; do echo "== v2 front 2^21 sets (64MB) minb $mb"; MCX_MINB=$mb run; done
echo "== v2 front 2^20 sets (32MB) minb 5"; MCX_FRONT_BITS=20 MCX_MINB=5 run
echo "== v2 front 2^22 sets (128MB) minb 5"; MCX_FRONT_BITS=22 MCX_MINB=5 run
MCX_MINB=5 ncu --set full --clock-control none --import-source on -k regex:mcx_build_fused -s 1 -c 1 -o gpurun_out/prof_front2_r1c python bench.py --reads 20000000 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_full_d.log 2>&1; tail -1 gpurun_out/ncu_full_d.log
MCX_MINB=5 MCX_BENCH_GENOME=1000000 MCX_BENCH_PERR=0 ncu --set full --clock-control none --import-source on -k regex:mcx_build_fused -s 1 -c 1 -o gpurun_out/prof_ceiling2_r1c python bench.py --reads 20000000 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_full_e.log 2>&1; tail -1 gpurun_out/ncu_full_e.log
