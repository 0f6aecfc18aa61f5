# experiment: front table sizes, L2 persistence; sector counts per load mode
M="lts__t_sectors_srcunit_tex_op_read.sum,lts__t_sectors_srcunit_tex_op_red.sum,dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,lts__t_sector_hit_rate.pct"
run() { python bench.py --reads 20000000 --steps 2 --warmup 1 --no-cpu-baseline 2>&1 | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('value %.2f G/s  kernel_ms %.1f e2e %.2f G/s distinct %s'%(d['value']/1e9, d['roofline']['kernel_ms'], d['e2e']['value']/1e9, d['extra']['distinct_kmers']))"; }
prof() { ncu --metrics $M --clock-control none -k regex:mcx_build_fused -s 1 -c 1 --csv python bench.py --reads 20000000 --steps 1 --warmup 1 --no-cpu-baseline 2>/dev/null | grep -E "mcx_build_fused" | python -c "
import sys,csv
for r in csv.reader(sys.stdin): print('   ', r[-3], r[-2], r[-1])"; }
echo "== front off, LD_MODE 0"; MCX_FRONT_MB=0 run; MCX_FRONT_MB=0 prof
echo "== front off, LD_MODE 3"; MCX_FRONT_MB=0 MCX_LD_MODE=3 prof
for mb in 32 64 96; do
  echo "== front $mb MB"; MCX_FRONT_MB=$mb run
  echo "== front $mb MB + L2 persist"; MCX_FRONT_MB=$mb MCX_L2_PERSIST=1 run
done
echo "== front 64 MB profile"; MCX_FRONT_MB=64 prof
echo "== front 64 MB + persist profile"; MCX_FRONT_MB=64 MCX_L2_PERSIST=1 prof
