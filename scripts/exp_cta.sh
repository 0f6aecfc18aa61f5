timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu 2>&1 | tail -2
run() { python bench.py --reads ${READS:-20000000} --steps 3 --warmup 2 --no-cpu-baseline 2>&1 | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('value %.2f G/s  kernel_ms %.1f'%(d['value']/1e9, d['roofline']['kernel_ms']))" 2>&1 | tail -1; }
for mb in 3 4; do for gg in 1 2; do echo "== minb $mb G $gg"; MCX_MINB=$mb MCX_G=$gg run; done; done
echo "== ceiling minb 3 G 2"; MCX_MINB=3 MCX_G=2 MCX_BENCH_GENOME=1000000 MCX_BENCH_PERR=0 run
