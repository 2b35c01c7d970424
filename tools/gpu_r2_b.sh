#!/bin/bash
# round-2 GPU trip B: truncation-compensation calibration, tier A/B at bench level
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
echo "== trunc comp"
timeout 600 python tools/diag_gpu.py trunccomp 2>&1 | grep -v Warn | tail -40
for T in "1,8" "4,8" "1,4" "2,8"; do
  echo "== bench tiers $T"
  NEFII_TRACE_TIERS=$T timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2> gpurun_out/r2b_bench_$T.err | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('ms/step %.2f  e2e ms %.2f  frame %.1f  sec %.1f  gemm TF %.1f share %.3f clocks %s' % (d['ms_per_step'], d['e2e']['ms_per_step'], d['ms_per_frame_800x800'], d['ms_per_secondary_training_pass'], d['roofline']['achieved'], d['roofline']['kernel_share_of_step'], d['clocks']))
"
done
