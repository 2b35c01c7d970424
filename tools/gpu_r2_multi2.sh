#!/bin/bash
# multi-GPU step only (no extras): usage gpu_r2_multi2.sh N [steps]
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
N=$1; STEPS=${2:-10}
export PYTHONUNBUFFERED=1
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps $STEPS --warmup 5 --no-extras --no-cpu-baseline > gpurun_out/r2x_bench_n$N.json 2> gpurun_out/r2x_bench_n$N.err
tail -c 3000 gpurun_out/r2x_bench_n$N.json; grep -v "Warn\|warn" gpurun_out/r2x_bench_n$N.err | tail -5
