#!/bin/bash
# compute-sanitizer racecheck / synccheck / memcheck of a small run (batch 2, plain launches so that every kernel is attributed)
set -u
mkdir -p gpurun_out
export PYTHONPATH=.
for tool in racecheck synccheck memcheck; do
  for prec in fast parity; do
    timeout 1200 compute-sanitizer --tool $tool --print-limit 30 python tools/one_step.py $prec 2 2 0 > gpurun_out/r2_sanitizer_${tool}_$prec.log 2>&1
    echo "== $tool $prec: $(grep -e 'SUMMARY' -e 'launches/step' gpurun_out/r2_sanitizer_${tool}_$prec.log | tr '\n' ' ')"
  done
done
HMDPOSE_MBFUSE=1 timeout 1200 compute-sanitizer --tool racecheck --print-limit 30 python tools/one_step.py fast 2 2 0 > gpurun_out/r2_sanitizer_racecheck_fast_mbfuse.log 2>&1
echo "== racecheck fast mbfuse: $(grep -e 'SUMMARY' -e 'launches/step' gpurun_out/r2_sanitizer_racecheck_fast_mbfuse.log | tr '\n' ' ')"
