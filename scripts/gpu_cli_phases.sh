#!/bin/bash
# wall clock of the host driver's phases (MCX_TIMING=1) at scale; no reference run.  usage: [reads31] [reads63]
set -u
R1=${1:-5000000}; R2=${2:-2000000}
D=/dev/shm/mcx_scale; mkdir -p $D gpurun_out
BIN=mccortex_b200/bin
one() { # reads k
  local R=$1 K=$2
  $BIN/mcx-synth 4600000 0 $R 150 0.001 1 > $D/reads.fa
  local NK=$(( (4600000 + R * 150 / 1000 * K) * 4 / 3 + 1000000 ))
  for rep in 1 2; do
    local t0=$(date +%s.%N)
    MCX_TIMING=1 $BIN/mccortex-b200 build -q -f -m 100G -n $NK -k $K -S --sample s --seq $D/reads.fa $D/gpu.ctx 2> $D/phases.txt; local rc=$?
    local t1=$(date +%s.%N)
    echo "== reads $R k=$K run $rep rc=$rc wall $(python3 -c "print('%.3f' % ($t1-$t0))") s  md5 $(md5sum < $D/gpu.ctx | cut -c1-12)"
    grep "^\[phase\]" $D/phases.txt
  done
}
one $R1 31
one $R2 63
rm -rf $D
