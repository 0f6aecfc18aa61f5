run() { python bench.py --reads 20000000 --steps 2 --warmup 1 --no-cpu-baseline 2>&1 | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('value %.2f G/s  kernel_ms %.1f e2e %.2f G/s'%(d['value']/1e9, d['roofline']['kernel_ms'], d['e2e']['value']/1e9))"; }
for cfg in "3 2" "3 4" "4 1" "4 2" "4 4" "5 1" "5 2" "6 1"; do set -- $cfg; echo "== minb $1 inflight $2"; MCX_MINB=$1 MCX_G=$2 run; done
for cfg in "4 2" "4 4" "5 2" "3 4"; do set -- $cfg; echo "== ceiling minb $1 inflight $2"; MCX_MINB=$1 MCX_G=$2 MCX_BENCH_GENOME=1000000 MCX_BENCH_PERR=0 run; done
