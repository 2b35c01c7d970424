#!/bin/bash
# padding rays written with fills (no synchronous H2D in the middle of a forward): tracer / pipeline tests, step A/B
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
echo "== tests"
timeout 900 python -m pytest tests/test_tracer_gpu.py tests/test_pipeline_gpu.py tests/test_parity_fullsize_gpu.py tests/test_mis_gpu.py -q -x 2>&1 | tail -4
run() {
  echo "== bench pixels=$1"
  NEFII_BENCH_PIXELS=$1 timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-extras 2> gpurun_out/r2y_$1.err | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('ms/step %.2f  e2e ms %.2f  value %.0f gemm TF %.1f share %.3f launches %d captures %d clocks %s' % (d['ms_per_step'], d['e2e']['ms_per_step'], d['value'], d['roofline']['achieved'], d['roofline']['kernel_share_of_step'], d['gpu_launches'], d['trace_graph_captures_in_timed_region'], d['clocks']))
" || tail -5 gpurun_out/r2y_$1.err
}
run 256
run 256
run 2048
