#!/bin/bash
# configs[4] subsample parity through the command line: R reads of the 3 Gbp genome, reference (CPU) vs `--shard` on N GPUs
set -u
N=${1:-2}; R=${2:-9375000}
mkdir -p gpurun_out
D=/dev/shm/mcx_c5; mkdir -p $D
BIN=mccortex_b200/bin; REF=oracle/_ref/mccortex31
T=$(nproc); [ $T -gt 32 ] && T=32
free -g | head -2
t0=$(date +%s.%N)
$BIN/mcx-synth 3000000000 0 $R 150 0.001 1 5 > $D/r.fa
t1=$(date +%s.%N)
NK=$(( R * 120 * 4 / 3 ))
$REF build -f -q -t $T -m 120G -n $NK -k 31 -S --sample s --seq $D/r.fa $D/ref.ctx > $D/ref.log 2>&1; rrc=$?
t2=$(date +%s.%N)
DEV=$(seq -s, 0 $((N-1)))
MCX_TIMING=1 $BIN/mccortex-b200 build -f -q -D $DEV --shard -m 150G -n $NK -k 31 -S --sample s --seq $D/r.fa $D/shard.ctx 2> $D/shard.log; grc=$?
t3=$(date +%s.%N)
$BIN/mccortex-b200 build -f -q -m 150G -n $NK -k 31 -S --sample s --seq $D/r.fa $D/one.ctx 2> $D/one.log; orc=$?
t4=$(date +%s.%N)
cmp -s $D/ref.ctx $D/shard.ctx && c1=identical || c1=DIFFER
cmp -s $D/ref.ctx $D/one.ctx && c2=identical || c2=DIFFER
python3 - <<PY
import os
occ = $R * 120
print("configs[4] subsample: %d reads of a 3 Gbp genome (synth %.1f s): reference -t $T %.1f s (rc $rrc) = %.1f M k-mers/s | %d GPUs --shard %.2f s (rc $grc) = %.2f G k-mers/s, cmp %s | 1 GPU %.2f s (rc $orc), cmp %s | .ctx %d bytes" % (
  $R, $t1 - $t0, $t2 - $t1, occ / ($t2 - $t1) / 1e6, $N, $t3 - $t2, occ / ($t3 - $t2) / 1e9, "$c1", $t4 - $t3, "$c2", os.path.getsize("$D/ref.ctx") if os.path.exists("$D/ref.ctx") else -1))
PY
tail -2 $D/ref.log; grep phase $D/shard.log | tail -6
md5sum $D/ref.ctx | cut -c1-32
rm -rf $D
