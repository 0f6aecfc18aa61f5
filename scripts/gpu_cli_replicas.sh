#!/bin/bash
# build -D 0,1,..: whole-process wall clock of the replica mode against one device (run under `gpurun --gpus N`).
# usage: scripts/gpu_cli_replicas.sh [reads] [ndevices]
set -u
R=${1:-50000000}; N=${2:-2}; K=31
D=/dev/shm/mcx_scale; mkdir -p $D gpurun_out
BIN=mccortex_b200/bin
$BIN/mcx-synth 4600000 0 $R 150 0.001 1 > $D/reads.fa
NK=$(( (4600000 + R * 150 / 1000 * K) * 4 / 3 + 1000000 ))
DEVS=$(seq -s, 0 $((N - 1)))
for devs in 0 $DEVS; do
  t0=$(date +%s.%N)
  MCX_TIMING=1 $BIN/mccortex-b200 build -q -f -m 100G -n $NK -k $K -S -D $devs --sample s --seq $D/reads.fa $D/out_$devs.ctx 2> $D/phases.txt; rc=$?
  t1=$(date +%s.%N)
  echo "== -D $devs rc=$rc wall $(python3 -c "print('%.3f' % ($t1-$t0))") s  md5 $(md5sum < $D/out_$devs.ctx | cut -c1-12)"
  grep "^\[phase\] [a-z]" $D/phases.txt
done
rm -rf $D
