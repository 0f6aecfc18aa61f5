run() { python bench.py --reads 20000000 --steps 2 --warmup 1 --no-cpu-baseline 2>&1 | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('value %.2f G/s  kernel_ms %.1f e2e %.2f G/s'%(d['value']/1e9, d['roofline']['kernel_ms'], d['e2e']['value']/1e9))"; }
for mb in 4 5 6; do echo "== ceiling minb $mb"; MCX_MINB=$mb MCX_BENCH_GENOME=1000000 MCX_BENCH_PERR=0 run; done
for mb in 4 5 6; do echo "== front64 minb $mb"; MCX_MINB=$mb run; done
for mb in 5; do echo "== front96 minb $mb"; MCX_FRONT_MB=96 MCX_MINB=$mb run; done
echo "== front64 4way minb 5"; MCX_FRONT_WAYS=4 MCX_MINB=5 run
echo "== nofront minb 5"; MCX_FRONT_MB=0 MCX_MINB=5 run
