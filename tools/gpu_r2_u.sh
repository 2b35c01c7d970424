#!/bin/bash
# cross-step prefetch of the primary trace: bit-identity test, A/B of the 1/8-batch and full-batch steps over the SM bound
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
echo "== tests"
timeout 600 python -m pytest tests/test_dense_mlp_gpu.py tests/test_pipeline_gpu.py -q -x 2>&1 | tail -6
run() {
  echo "== bench pixels=$1 prefetch=$2 sms=$3"
  NEFII_BENCH_PIXELS=$1 NEFII_BENCH_PREFETCH=$2 NEFII_PREFETCH_SMS=$3 timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-extras 2> gpurun_out/r2u_$1_$2_$3.err | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('ms/step %.2f  e2e ms %.2f  value %.0f gemm TF %.1f share %.3f launches %d captures %d clocks %s' % (d['ms_per_step'], d['e2e']['ms_per_step'], d['value'], d['roofline']['achieved'], d['roofline']['kernel_share_of_step'], d['gpu_launches'], d['trace_graph_captures_in_timed_region'], d['clocks']))
" || tail -5 gpurun_out/r2u_$1_$2_$3.err
}
run 256 0 0
run 256 1 96
run 256 1 128
run 256 1 0
run 2048 0 0
run 2048 1 128
