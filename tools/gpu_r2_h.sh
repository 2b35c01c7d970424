#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
echo "== gpu suite"
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -8
echo "== tracer tests, fixed schedule"
NEFII_TRACE_GRAPH=0 timeout 900 python -m pytest tests/test_tracer_gpu.py -q -x 2>&1 | tail -3
for PX in 256 512; do
for Q in 12288 0; do
  echo "== bench pixels=$PX quad_rows=$Q"
  NEFII_BENCH_PIXELS=$PX NEFII_TRACE_QUAD_ROWS=$Q timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-extras 2> gpurun_out/r2h_$PX_$Q.err | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('ms/step %.2f  e2e ms %.2f  value %.0f gemm TF %.1f share %.3f clocks %s' % (d['ms_per_step'], d['e2e']['ms_per_step'], d['value'], d['roofline']['achieved'], d['roofline']['kernel_share_of_step'], d['clocks']))
"
done
done
