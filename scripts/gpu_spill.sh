#!/bin/bash
# MCX_SPILL experiment: parity test of the spill path, then the N=1 bench with it (baseline: profiles/r1j_bench_n1.json)
set -u
mkdir -p gpurun_out
timeout 200 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "spill" > gpurun_out/pytest_spill.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_spill.log
tail -4 gpurun_out/pytest_spill.log
run() { # name, env...
  local name=$1; shift
  env "$@" timeout 150 python bench.py --no-cpu-baseline > gpurun_out/bench_$name.json 2> gpurun_out/bench_$name.err; echo "$name rc=$?"
  python - <<PY
import json
try:
    d = json.load(open("gpurun_out/bench_$name.json"))
    print("$name", "value %.2f G/s" % (d["value"] / 1e9), "ms/step %.1f" % d["ms_per_step"], "e2e %.2f G/s" % (d["e2e"]["value"] / 1e9), "launches", d.get("gpu_launches"))
except Exception as e:
    print("$name: no result", e)
PY
}
for v in "$@"; do
  case $v in
    s1024) run spill2_1024 MCX_SPILL=1 ;;
    s3584) run spill2_3584 MCX_SPILL=1 MCX_SPILL_SPAN_MB=3584 ;;
    base) run base MCX_SPILL=0 ;;
  esac
done
