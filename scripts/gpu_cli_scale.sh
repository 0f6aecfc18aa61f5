#!/bin/bash
# the drop-in command line at scale, against the compiled reference on the same box: byte-identical sorted .ctx
# (cmp) and wall-clock of the whole process, file -> .ctx.   usage: scripts/gpu_cli_scale.sh [reads] [k]
set -u
R=${1:-5000000}; K=${2:-31}
D=/dev/shm/mcx_scale; mkdir -p $D gpurun_out
BIN=mccortex_b200/bin; REF=oracle/_ref/mccortex31; [ "$K" -gt 31 ] && REF=oracle/_ref/mccortex63
$BIN/mcx-synth 4600000 0 $R 150 0.001 1 > $D/reads.fa
ls -la $D/reads.fa
NK=$(( (4600000 + R * 150 / 1000 * K) * 4 / 3 + 1000000 ))
t0=$(date +%s.%N)
$BIN/mccortex-b200 build -q -f -m 100G -n $NK -k $K -S --sample s --seq $D/reads.fa $D/gpu.ctx; rc1=$?
t1=$(date +%s.%N)
$REF build -q -f -t $(nproc) -m 100G -n $NK -k $K -S --sample s --seq $D/reads.fa $D/ref.ctx; rc2=$?
t2=$(date +%s.%N)
cmp $D/gpu.ctx $D/ref.ctx; rc3=$?
python3 - <<PY
import os
g=$t1-$t0; r=$t2-$t1
occ=$R*(150-$K+1)
print("reads %d k=%d: %d k-mer occurrences, .ctx %d bytes, cmp rc=%d (build rc %d / %d)" % ($R, $K, occ, os.path.getsize("$D/gpu.ctx"), $rc3, $rc1, $rc2))
print("mccortex-b200 build: %.2f s (%.1f M k-mers/s, whole process: FASTA parse on one host thread, H2D, kernels, sorted export, write)" % (g, occ/g/1e6))
print("reference build -t %d: %.2f s (%.1f M k-mers/s)  => %.1fx" % (os.cpu_count(), r, occ/r/1e6, r/g))
PY
rm -rf $D
