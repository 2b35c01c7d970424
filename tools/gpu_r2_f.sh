#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
echo "== gpu suite summary"
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -4
echo "== ncu launch list of a 256-pixel step (fixed schedule)"
PROFILE_PIXELS=256 NEFII_TRACE_GRAPH=0 timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/r2f_launches_256.csv python tools/profile_step.py > gpurun_out/r2f_ncu_step.log 2>&1; tail -2 gpurun_out/r2f_ncu_step.log; wc -l gpurun_out/r2f_launches_256.csv
