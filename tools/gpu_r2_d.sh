#!/bin/bash
# round-2 GPU trip D: per-rank-of-8 sized step on one GPU (fixed cost of the step), graph vs fixed schedule
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
for PX in 256 512; do
for G in 1 0; do
  echo "== bench pixels=$PX graph=$G"
  NEFII_BENCH_PIXELS=$PX NEFII_TRACE_GRAPH=$G timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-extras 2> gpurun_out/r2d_$PX_$G.err | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('ms/step %.2f  e2e ms %.2f  value %.0f launches/step %.0f gemm TF %.1f share %.3f rays/step %.0f clocks %s' % (d['ms_per_step'], d['e2e']['ms_per_step'], d['value'], d['gpu_launches']/d['steps'], d['roofline']['achieved'], d['roofline']['kernel_share_of_step'], d['config']['rays_per_step'], d['clocks']))
"
done
done
echo "== torch profiler of a 256-pixel step (CPU + CUDA time by op)"
NEFII_BENCH_PIXELS=256 timeout 600 python tools/profile_small_step.py 2>&1 | grep -v Warn | tail -60
