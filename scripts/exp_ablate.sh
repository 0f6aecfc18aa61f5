run() { python bench.py --reads ${READS:-20000000} --steps 3 --warmup 2 --no-cpu-baseline 2>&1 | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('value %.2f G/s  kernel_ms %.1f'%(d['value']/1e9, d['roofline']['kernel_ms']))" 2>&1 | tail -1; }
export MCX_MINB=${MINB:-3} MCX_G=${G:-2}
echo "== full"; run
echo "== no table access (front end + k-mers + hash only)"; MCX_L2_HINTS=8 run
echo "== load + compare only"; MCX_L2_HINTS=16 run
echo "== RED only"; MCX_L2_HINTS=32 run
export MCX_MINB=4
echo "== minb4 no table access"; MCX_L2_HINTS=8 run
echo "== minb4 load + compare only"; MCX_L2_HINTS=16 run
echo "== minb4 RED only"; MCX_L2_HINTS=32 run
