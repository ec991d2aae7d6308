#!/bin/bash
# compute-sanitizer racecheck / synccheck / memcheck of a small run (batch 2, plain launches so that every kernel is attributed):
# the default plans of both precision modes (batch 2 = a latency plan: split-K project kernel, gate-folded project weights)
# and the two opt-in fused kernels
set -u
mkdir -p gpurun_out
export PYTHONPATH=.
run() {  # name, tool, precision, env...
  local name=$1 tool=$2 prec=$3; shift 3
  env "$@" timeout 1200 compute-sanitizer --tool $tool --print-limit 30 python tools/one_step.py $prec 2 2 0 > gpurun_out/r2_sanitizer_$name.log 2>&1
  echo "== $name: $(grep -e 'SUMMARY' -e 'launches/step' gpurun_out/r2_sanitizer_$name.log | tr '\n' ' ')"
}
for tool in racecheck synccheck memcheck; do
  run ${tool}_fast $tool fast X=1
done
run racecheck_parity racecheck parity X=1
run synccheck_parity synccheck parity X=1
for tool in racecheck synccheck memcheck; do
  run ${tool}_fast_mbfuse $tool fast HMDPOSE_MBFUSE=1
  run ${tool}_fast_expdw $tool fast HMDPOSE_EXPDW=1
done
