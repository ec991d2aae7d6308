#!/bin/bash
# 8-GPU survival evidence: the bench under torchrun several times in a row, every rank's exit code kept.
set -u
mkdir -p gpurun_out
N=${1:-8}; LOOPS=${2:-4}
for i in $(seq 1 $LOOPS); do
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29510 + i)) \
    bench.py --gpus $N --headline-only --no-cpu-baseline > gpurun_out/r2_bench${N}_$i.json 2> gpurun_out/r2_bench${N}_$i.err
  echo "loop $i rc=$? $(head -c 160 gpurun_out/r2_bench${N}_$i.json)"
done
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29540 \
  bench.py --gpus $N --no-cpu-baseline > gpurun_out/r2_bench${N}_full.json 2> gpurun_out/r2_bench${N}_full.err
echo "full rc=$? $(head -c 160 gpurun_out/r2_bench${N}_full.json)"
