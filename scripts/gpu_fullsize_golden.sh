#!/bin/bash
# Full-size byte parity, once, recorded: the compiled reference (`oracle/_ref/mccortex31|63 build -S`,
# CPU) and `mccortex-b200 build -S` (GPU) on the stated inputs of configs 1a, 1b, 2, 3, 4; `cmp` of the two files and the
# md5 / size of the reference's -> gpurun_out/fullsize_md5.json (committed as tests/golden/fullsize_md5.json).
set -u
D=/dev/shm/mcx_full; mkdir -p $D gpurun_out
BIN=mccortex_b200/bin; REF=oracle/_ref
T=$(nproc); [ $T -gt 32 ] && T=32
OUT=gpurun_out/fullsize_md5.json
echo "{" > $OUT
first=1
one() { # name k nslots ref-args... (inputs are given as --sample/--seq arguments valid for both programs)
  name=$1; k=$2; n=$3; shift 3
  refbin=$REF/mccortex31; [ $k -gt 31 ] && refbin=$REF/mccortex63
  t0=$(date +%s.%N)
  $refbin build -f -q -t $T -m 60G -n $n -k $k -S "$@" $D/ref.ctx > $D/ref.log 2>&1; rrc=$?
  t1=$(date +%s.%N)
  $BIN/mccortex-b200 build -f -q -m 100G -n $n -k $k -S "$@" $D/gpu.ctx > $D/gpu.log 2>&1; grc=$?
  t2=$(date +%s.%N)
  if cmp -s $D/ref.ctx $D/gpu.ctx; then same=identical; else same=DIFFER; fi
  rmd5=$(md5sum $D/ref.ctx | cut -c1-32); gmd5=$(md5sum $D/gpu.ctx | cut -c1-32)
  rsz=$(stat -c %s $D/ref.ctx); gsz=$(stat -c %s $D/gpu.ctx)
  [ $first -eq 0 ] && echo "," >> $OUT; first=0
  python3 - >> $OUT <<PY
import json
print(' %s: %s' % (json.dumps("$name"), json.dumps({"k": $k, "nslots": $n, "md5": "$rmd5", "bytes": $rsz, "gpu_md5": "$gmd5", "gpu_bytes": $gsz,
  "cmp": "$same", "ref_rc": $rrc, "gpu_rc": $grc, "ref_seconds": round($t1 - $t0, 1), "gpu_seconds": round($t2 - $t1, 2), "ref_threads": $T,
  "args": "build -k $k -n $n -S " + " ".join("""$*""".replace("$D/", "").split())})), end="")
PY
  echo "$name: ref rc=$rrc $(printf %.0f $(echo "$t1 - $t0" | bc -l 2>/dev/null || echo 0)) s, gpu rc=$grc, cmp $same, md5 $rmd5 size $rsz"
  rm -f $D/ref.ctx $D/gpu.ctx
}
WHAT=${1:-all}
# configs[0] 1a: one 1,000,000-base random record wrapped at 80 columns; 1b: 6,536 x 150 bp reads of a 100,000-base genome
python3 - <<PY
import ctypes as C, os
L = C.CDLL("mccortex_b200/lib/libmcxsynth.so")
L.mcx_synth_genome.argtypes = [C.c_void_p, C.c_uint64, C.c_uint64]
g = C.create_string_buffer(1000000); L.mcx_synth_genome(g, 1000000, 0x1A)
s = "".join("ACGT"[b] for b in g.raw[:1000000])
open("$D/c1a.fa", "w").write(">one\n" + "\n".join(s[i:i + 80] for i in range(0, len(s), 80)) + "\n")
PY
$BIN/mcx-synth 100000 0 6536 150 0.001 1 > $D/c1b.fa
one config1a_1Mbase_record_80col 31 2000000 --sample s --seq $D/c1a.fa
one config1b_6536_reads 31 1000000 --sample s --seq $D/c1b.fa
if [ "$WHAT" != "small" ]; then
  $BIN/mcx-synth 4600000 0 50000000 150 0.001 1 > $D/r.fa
  one config2_50M_reads_k31 31 331633333 --sample s --seq $D/r.fa
  one config3_50M_reads_k63 63 669000000 --sample s --seq $D/r.fa
  rm -f $D/r.fa
  for c in 0 1 2 3; do $BIN/mcx-synth 4600000 $((c * 25000000)) 25000000 150 0.001 1 > $D/s$c.fa; done
  one config4_4x25M_reads_4_colours 31 663000000 --sample s0 --seq $D/s0.fa --sample s1 --seq $D/s1.fa --sample s2 --seq $D/s2.fa --sample s3 --seq $D/s3.fa
fi
echo "" >> $OUT; echo "}" >> $OUT
rm -rf $D
cat $OUT
