#!/bin/bash
# Round-2 measurement pass on one GPU: the default bench line, the reference arm, the ncu launch list of the bench
# command (shares only: cold-cache, serialised) and the per-launch DRAM traffic of one whole step.
set -u
mkdir -p gpurun_out
export PYTHONPATH=.
python bench.py > gpurun_out/r2_bench.json 2> gpurun_out/r2_bench.err
tail -c 400 gpurun_out/r2_bench.json
python bench.py --impl reference --steps 10 --warmup 2 > gpurun_out/r2_bench_reference_cpu.json 2>> gpurun_out/r2_bench.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_ncu_launch_list.csv \
  python bench.py --steps 2 --warmup 1 --headline-only --no-cpu-baseline > gpurun_out/r2_launches_bench.log 2>&1
L=$(python tools/one_step.py fast 16 1 | sed -n 's/launches\/step \([0-9]*\).*/\1/p')
echo "launches per step: $L"
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none \
  --launch-skip $((3 * L)) --launch-count $L --csv --log-file gpurun_out/r2_step_traffic.csv python tools/one_step.py fast 16 4 > gpurun_out/r2_step_traffic.log 2>&1
tail -1 gpurun_out/r2_step_traffic.log
python tools/gpu_check.py insitu > /dev/null 2>&1
cp gpurun_out/check_insitu.txt gpurun_out/r2_steps_b16_insitu.txt
