set -u
mkdir -p gpurun_out
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 8 --headline-only --no-cpu-baseline > gpurun_out/r2_bench8_last.json 2> gpurun_out/r2_bench8_last.err
echo "rc=$? $(grep '^{' gpurun_out/r2_bench8_last.json | head -c 200)"
