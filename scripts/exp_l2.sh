run() { python bench.py --reads ${READS:-20000000} --steps 3 --warmup 2 --no-cpu-baseline 2>&1 | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('value %.2f G/s  kernel_ms %.1f e2e %.2f G/s'%(d['value']/1e9, d['roofline']['kernel_ms'], d['e2e']['value']/1e9))"; }
export MCX_MINB=${MINB:-3} MCX_G=${G:-2}
python -c "
import ctypes
l=ctypes.CDLL('libcudart.so.12') if False else None
"
echo "== base"; run
for h in 1 2 4 3 5 7; do echo "== hints $h"; MCX_L2_HINTS=$h run; done
for mb in 64 96 110; do echo "== persist $mb MB"; MCX_L2_PERSIST_MB=$mb run; done
echo "== persist 96 + hints 3"; MCX_L2_PERSIST_MB=96 MCX_L2_HINTS=3 run
echo "== S=20"; MCX_FRONT_BITS=20 run
echo "== S=22"; MCX_FRONT_BITS=22 run
echo "== S=20 hints 7"; MCX_FRONT_BITS=20 MCX_L2_HINTS=7 run
