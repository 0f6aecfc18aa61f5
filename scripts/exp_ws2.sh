run() { timeout 300 python bench.py --reads ${READS:-20000000} --steps 3 --warmup 2 --no-cpu-baseline 2>&1 | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('value %.2f G/s  kernel_ms %.1f'%(d['value']/1e9, d['roofline']['kernel_ms']))" 2>&1 | tail -1; }
for v in 2 3 4 5 6; do
  echo "== WS variant $v real";    MCX_WS=$v run
  echo "== WS variant $v ceiling"; MCX_WS=$v MCX_BENCH_GENOME=1000000 MCX_BENCH_PERR=0 run
done
