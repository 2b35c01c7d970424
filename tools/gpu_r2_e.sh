#!/bin/bash
# round-2 GPU trip E: new tests (small ops, full-size parity, tightened pipeline tolerances), split-N effect on small steps
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
echo "== gpu suite"
timeout 1500 python -m pytest tests -m gpu -q -s 2>&1 | grep -v "Warn\|warn" | grep -E "passed|failed|FAILED|Error|error|assert|camera rays|pipeline train|grad rel|forward_with_point|FULL|depth|sg_rgb_values|secondary mask|grad rel" | tail -60
for PX in 256 2048; do
  echo "== bench pixels=$PX"
  NEFII_BENCH_PIXELS=$PX timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-extras 2> gpurun_out/r2e_$PX.err | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('ms/step %.2f  e2e ms %.2f  value %.0f launches/step %.0f gemm TF %.1f share %.3f rays/step %.0f clocks %s' % (d['ms_per_step'], d['e2e']['ms_per_step'], d['value'], d['gpu_launches']/d['steps'], d['roofline']['achieved'], d['roofline']['kernel_share_of_step'], d['config']['rays_per_step'], d['clocks']))
"
done
