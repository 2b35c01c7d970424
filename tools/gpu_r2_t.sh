#!/bin/bash
# native dense-stack sequencer: its tests, A/B of the 1/8-batch and full-batch steps, host enqueue profile
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
echo "== dense stack tests"
timeout 600 python -m pytest tests/test_dense_mlp_gpu.py tests/test_pipeline_gpu.py -q -x 2>&1 | tail -6
for PX in 256 2048; do
for NAT in 1 0; do
  echo "== bench pixels=$PX native=$NAT"
  NEFII_BENCH_PIXELS=$PX NEFII_DENSE_NATIVE=$NAT timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-extras 2> gpurun_out/r2t_${PX}_$NAT.err | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('ms/step %.2f  e2e ms %.2f  value %.0f gemm TF %.1f share %.3f launches %d clocks %s' % (d['ms_per_step'], d['e2e']['ms_per_step'], d['value'], d['roofline']['achieved'], d['roofline']['kernel_share_of_step'], d['gpu_launches'], d['clocks']))
"
done
done
echo "== host profile 256 px"
PX=256 timeout 300 python tools/diag_host.py 2>&1 | head -60
