run() { python bench.py --reads ${READS:-20000000} --steps 3 --warmup 2 --no-cpu-baseline 2>&1 | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('value %.2f G/s  kernel_ms %.1f'%(d['value']/1e9, d['roofline']['kernel_ms']))" 2>&1 | tail -1; }
for mb in 2 3; do for gg in 1 2 4; do echo "== minb $mb G $gg"; MCX_MINB=$mb MCX_G=$gg run; done; done
echo "== ceiling minb 3 G 4"; MCX_MINB=3 MCX_G=4 MCX_BENCH_GENOME=1000000 MCX_BENCH_PERR=0 run
echo "== ceiling minb 2 G 4"; MCX_MINB=2 MCX_G=4 MCX_BENCH_GENOME=1000000 MCX_BENCH_PERR=0 run
echo "== perr0 minb 3 G 4"; MCX_MINB=3 MCX_G=4 MCX_BENCH_PERR=0 run
