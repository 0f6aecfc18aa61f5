#!/bin/bash
# configs[1] through the command line: 50 M x 150 bp reads (7.7 GB FASTA in /dev/shm) -> sorted .ctx, phases on stderr
set -u
R=${1:-50000000}; K=31
D=/dev/shm/mcx_scale; mkdir -p $D gpurun_out
BIN=mccortex_b200/bin
t0=$(date +%s.%N)
$BIN/mcx-synth 4600000 0 $R 150 0.001 1 > $D/reads.fa
t1=$(date +%s.%N)
ls -la $D/reads.fa; nproc
NK=$(( (4600000 + R * 150 / 1000 * K) * 4 / 3 + 1000000 ))
MCX_TIMING=1 $BIN/mccortex-b200 build -f -m 100G -n $NK -k $K -S --sample s --seq $D/reads.fa $D/gpu.ctx 2> $D/log.txt; rc=$?
t2=$(date +%s.%N)
grep "phase\|occupancy" $D/log.txt
python3 - <<PY
import os
occ = $R * (150 - $K + 1)
w = $t2 - $t1
print("reads %d k=%d rc=%d: %d k-mer occurrences, .ctx %d bytes, whole process %.2f s = %.1f M k-mers/s (synth %.1f s)" % ($R, $K, $rc, occ, os.path.getsize("$D/gpu.ctx"), w, occ / w / 1e6, $t1 - $t0))
PY
md5sum $D/gpu.ctx | cut -c1-32
rm -rf $D
