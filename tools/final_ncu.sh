#!/bin/bash
# `ncu --set full` captures of the heaviest kernels (three launches each keeps gpurun_out/ under the 64 MiB limit)
set -u
mkdir -p gpurun_out
export PYTHONPATH=.
for k in ${KERNELS:-gemm_tc2_kernel sepconv3_kernel dw3_kernel sepconv_kernel}; do
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:^$k -c 3 -f -o gpurun_out/final_ncu_$k \
    python bench.py --inflight 1 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/final_ncu_$k.log 2>&1
done
ls -la gpurun_out | grep final
