#!/bin/bash
# Round-2 profiling pass (one GPU): per-launch DRAM traffic of one whole step (all launches, not hand-picked ones),
# `ncu --set full` of the parity-mode GEMM, compute-sanitizer racecheck / synccheck of a small run.
set -u
mkdir -p gpurun_out
export PYTHONPATH=.
# 1. every launch of the 4th step (3 warm-up steps skipped): duration + DRAM bytes
L=$(python tools/one_step.py fast 16 1 | sed -n 's/launches\/step \([0-9]*\).*/\1/p')
echo "launches per step: $L"
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none \
  --launch-skip $((3 * L)) --launch-count $L --csv --log-file gpurun_out/r2_step_traffic.csv python tools/one_step.py fast 16 4 > gpurun_out/r2_step_traffic.log 2>&1
tail -2 gpurun_out/r2_step_traffic.log
# 2. full capture of the parity-mode GEMM (3 launches) and of the fast GEMM (3 launches)
timeout 900 ncu --set full --clock-control none --import-source on -k regex:^gemm_tf32_kernel --launch-skip 70 -c 3 -f -o gpurun_out/r2_ncu_gemm_tf32 \
  python tools/one_step.py parity 16 3 > gpurun_out/r2_ncu_gemm_tf32.log 2>&1
tail -1 gpurun_out/r2_ncu_gemm_tf32.log
# 3. compute-sanitizer (slow: batch 2, plain launches so that every kernel is attributed)
for tool in racecheck synccheck; do
  timeout 900 compute-sanitizer --tool $tool --print-limit 20 python tools/one_step.py fast 2 2 0 > gpurun_out/r2_sanitizer_$tool.log 2>&1
  tail -3 gpurun_out/r2_sanitizer_$tool.log
done
timeout 900 compute-sanitizer --tool racecheck --print-limit 20 python tools/one_step.py parity 2 2 0 > gpurun_out/r2_sanitizer_racecheck_parity.log 2>&1
tail -3 gpurun_out/r2_sanitizer_racecheck_parity.log
