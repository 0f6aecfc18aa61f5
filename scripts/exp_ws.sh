export MCX_WS=1
timeout 120 python -c "
import random, sys
sys.path.insert(0,'.'); sys.path.insert(0,'tests')
import mccortex_b200 as M
from oracle import oracle as O
from conftest import rand_reads, oracle_records
rng = random.Random(1)
reads = rand_reads(rng, 500, 150, 20000)
recs, ost = oracle_records(O, reads, 31)
g = M.Graph(31, 1, 1 << 20)
g.add_lines(''.join(r + '\n' for r in reads).encode())
st = g.sync()
got, n, rb = g.export_records()
print('WS smoke', got == recs, st.num_kmers_loaded == ost.num_kmers_loaded, st.num_kmers_novel == ost.num_kmers_novel, st.contigs_parsed == ost.contigs_parsed)
"
echo "smoke rc=$?"
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu 2>&1 | tail -3
run() { timeout 300 python bench.py --reads ${READS:-20000000} --steps 3 --warmup 2 --no-cpu-baseline 2>&1 | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('value %.2f G/s  kernel_ms %.1f'%(d['value']/1e9, d['roofline']['kernel_ms']))" 2>&1 | tail -1; }
echo "== WS real"; run
echo "== WS ceiling"; MCX_BENCH_GENOME=1000000 MCX_BENCH_PERR=0 run
echo "== WS perr0"; MCX_BENCH_PERR=0 run
export MCX_WS=0
echo "== fused real"; run
