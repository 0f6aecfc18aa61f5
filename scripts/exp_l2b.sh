run() { python bench.py --reads ${READS:-20000000} --steps 3 --warmup 2 --no-cpu-baseline 2>&1 | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('value %.2f G/s  kernel_ms %.1f'%(d['value']/1e9, d['roofline']['kernel_ms']))" 2>&1 | tail -1; }
echo "== base"; run
for h in 4 64 128 68 132 5 69 133 135; do echo "== hints $h"; MCX_L2_HINTS=$h run; done
echo "== S=20"; MCX_FRONT_BITS=20 run
