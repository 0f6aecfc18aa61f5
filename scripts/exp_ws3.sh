export MCX_WS=3
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu 2>&1 | tail -2
run() { timeout 300 python bench.py --reads ${READS:-20000000} --steps 3 --warmup 2 --no-cpu-baseline 2>&1 | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('value %.2f G/s  kernel_ms %.1f'%(d['value']/1e9, d['roofline']['kernel_ms']))" 2>&1 | tail -1; }
for v in 3 4 5 6 2; do
  echo "== WS variant $v real";    MCX_WS=$v run
done
echo "== WS variant 3 perr0";   MCX_WS=3 MCX_BENCH_PERR=0 run
echo "== WS variant 3 ceiling"; MCX_WS=3 MCX_BENCH_GENOME=1000000 MCX_BENCH_PERR=0 run
