#!/bin/bash
# N GPUs, one process: `mccortex-b200 build -D 0,..,N-1 --shard` -- CLI tests, then configs[1] x N/… at scale with cmp against the single-device build
set -u
N=${1:-2}; R=${2:-50000000}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_zy_gpu_cli_new.py tests/test_gpu_multi.py -x -q -m gpu -k "shard or routed" 2>&1 | tail -3
D=/dev/shm/mcx_shard; mkdir -p $D
BIN=mccortex_b200/bin
$BIN/mcx-synth 4600000 0 $R 150 0.001 1 > $D/r.fa
NK=$(( (4600000 + R * 150 / 1000 * 31) * 4 / 3 + 1000000 ))
DEV=$(seq -s, 0 $((N-1)))
t0=$(date +%s.%N)
MCX_TIMING=1 $BIN/mccortex-b200 build -f -q -m 100G -n $NK -k 31 -S --sample s --seq $D/r.fa $D/one.ctx 2> $D/one.log; rc1=$?
t1=$(date +%s.%N)
MCX_TIMING=1 $BIN/mccortex-b200 build -f -q -D $DEV --shard -m 100G -n $NK -k 31 -S --sample s --seq $D/r.fa $D/shard.ctx 2> $D/shard.log; rc2=$?
t2=$(date +%s.%N)
MCX_TIMING=1 $BIN/mccortex-b200 build -f -q -D $DEV -m 100G -n $NK -k 31 -S --sample s --seq $D/r.fa $D/repl.ctx 2> $D/repl.log; rc3=$?
t3=$(date +%s.%N)
cmp -s $D/one.ctx $D/shard.ctx && c12=identical || c12=DIFFER
cmp -s $D/one.ctx $D/repl.ctx && c13=identical || c13=DIFFER
python3 - <<PY
occ = $R * 120
print("reads %d: 1 GPU %.2f s (rc %d) = %.2f G k-mers/s | %d GPUs --shard %.2f s (rc %d) = %.2f G/s, cmp %s | %d GPUs replicas %.2f s (rc %d) = %.2f G/s, cmp %s" % (
  $R, $t1 - $t0, $rc1, occ / ($t1 - $t0) / 1e9, $N, $t2 - $t1, $rc2, occ / ($t2 - $t1) / 1e9, "$c12", $N, $t3 - $t2, $rc3, occ / ($t3 - $t2) / 1e9, "$c13"))
PY
grep phase $D/shard.log | tail -12
md5sum $D/one.ctx | cut -c1-32
rm -rf $D
