M="lts__t_sectors_srcunit_tex_op_read.sum,lts__t_sectors_srcunit_tex_op_red.sum,dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,lts__t_sector_hit_rate.pct"
for mode in 0 1 2 3 4; do
  echo "== LD_MODE $mode"
  MCX_LD_MODE=$mode python bench.py --reads 20000000 --steps 2 --warmup 1 --no-cpu-baseline 2>&1 | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('value %.2f G/s  kernel_ms %.1f e2e %.2f G/s'%(d['value']/1e9, d['roofline']['kernel_ms'], d['e2e']['value']/1e9))"
  MCX_LD_MODE=$mode ncu --metrics $M --clock-control none -k regex:mcx_build_fused -s 1 -c 1 --csv python bench.py --reads 20000000 --steps 1 --warmup 1 --no-cpu-baseline 2>/dev/null | grep -E "mcx_build_fused" | awk -F'","' '{print $(NF-2), $(NF-1), $NF}'
done
for f in 32 128; do
  echo "== L2FETCH $f (LD_MODE 1)"
  MCX_L2FETCH=$f MCX_LD_MODE=1 python bench.py --reads 20000000 --steps 2 --warmup 1 --no-cpu-baseline 2>&1 | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('value %.2f G/s  kernel_ms %.1f'%(d['value']/1e9, d['roofline']['kernel_ms']))"
  MCX_L2FETCH=$f MCX_LD_MODE=1 ncu --metrics $M --clock-control none -k regex:mcx_build_fused -s 1 -c 1 --csv python bench.py --reads 20000000 --steps 1 --warmup 1 --no-cpu-baseline 2>/dev/null | grep -E "mcx_build_fused" | awk -F'","' '{print $(NF-2), $(NF-1), $NF}'
done
