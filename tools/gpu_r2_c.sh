#!/bin/bash
# round-2 GPU trip C: defaults = k_flush 8 + truncation compensation.  Suite, parity at full size, bench, ncu.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
echo "== whole gpu suite"
timeout 1200 python -m pytest tests -m gpu -q -x 2>&1 | tail -15
echo "== trunccomp check (defaults on)"
timeout 300 python tools/diag_gpu.py kflush 2>&1 | grep KFLUSH
echo "== fullsize parity diag (defaults), then comp off for comparison"
timeout 900 python tools/diag_gpu.py fullsize 2>&1 | grep -v Warn | tail -60
NEFII_GEMM_TRUNC_COMP=0 FULL_PX=2048 timeout 900 python tools/diag_gpu.py fullsize 2>&1 | grep -v Warn | head -16
echo "== bench"
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r2c_bench.json 2> gpurun_out/r2c_bench.err; tail -c 6000 gpurun_out/r2c_bench.json; tail -5 gpurun_out/r2c_bench.err
echo "== ncu launch list (fixed schedule)"
NEFII_TRACE_GRAPH=0 timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/r2c_launches.csv python tools/profile_step.py > gpurun_out/r2c_ncu_step.log 2>&1; tail -3 gpurun_out/r2c_ncu_step.log; wc -l gpurun_out/r2c_launches.csv
echo "== ncu full of the hidden-layer GEMM"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_split -s 5 -c 1 -o gpurun_out/r2c_gemm python tools/diag_gpu.py gemmprof > gpurun_out/r2c_ncu_gemm.log 2>&1; tail -3 gpurun_out/r2c_ncu_gemm.log
ncu -i gpurun_out/r2c_gemm.ncu-rep --page raw --csv > gpurun_out/r2c_gemm_raw.csv 2>/dev/null; wc -c gpurun_out/r2c_gemm_raw.csv
