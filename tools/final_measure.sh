#!/bin/bash
# One measurement pass on the GPU box: bench lines, latency harness, per-step timings, ncu launch list and
# `ncu --set full` captures of the three heaviest kernels.  Everything lands in gpurun_out/ (copied to profiles/ by hand).
set -u
mkdir -p gpurun_out
export PYTHONPATH=.
python bench.py > gpurun_out/final_bench.json 2> gpurun_out/final_bench.err
tail -c 1500 gpurun_out/final_bench.json
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/final_bench_reference.json 2>> gpurun_out/final_bench.err
python tools/make_blob.py /tmp/hmdpose_phi0.blob > /dev/null 2>&1
{
  for prec in 1 0; do
    tools/pinvoke_harness hmd_ego_pose_b200/lib/libhmdpose.so /tmp/hmdpose_phi0.blob 256 3000 300 $prec
  done
  tools/pinvoke_harness hmd_ego_pose_b200/lib/libhmdpose.so /tmp/hmdpose_phi0.blob 256 3000 300 1 1
} > gpurun_out/final_latency.json 2>&1
cat gpurun_out/final_latency.json
python tools/gpu_check.py insitu > /dev/null 2>&1
python tools/gpu_check.py steps > /dev/null 2>&1
# launch list of the bench command (cold-cache, serialised: shares only)
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/final_launches.csv \
  python bench.py --steps 2 --warmup 1 > gpurun_out/final_launches_bench.log 2>&1
ls -la gpurun_out | grep final
