#!/bin/bash
# round-2 final GPU trip: suite, bench, ncu launch lists (full step and 1/8 step, fixed schedule), ncu --set full of the layer GEMM
# (fp16-split planes) and of the render_with_sg kernels.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
echo "== whole gpu suite"
timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -4
echo "== bench"
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r2z_bench.json 2> gpurun_out/r2z_bench.err; tail -c 3000 gpurun_out/r2z_bench.json; tail -3 gpurun_out/r2z_bench.err
echo "== ncu launch list (fixed schedule)"
NEFII_TRACE_GRAPH=0 timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/r2z_launches.csv python tools/profile_step.py > gpurun_out/r2z_ncu_step.log 2>&1; tail -2 gpurun_out/r2z_ncu_step.log; wc -l gpurun_out/r2z_launches.csv
PROFILE_PIXELS=256 NEFII_TRACE_GRAPH=0 timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/r2z_launches_256.csv python tools/profile_step.py > gpurun_out/r2z_ncu_step256.log 2>&1; wc -l gpurun_out/r2z_launches_256.csv
echo "== ncu full of the hidden-layer GEMM (fp16-split planes)"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_split -s 5 -c 1 -o gpurun_out/r2z_gemm python tools/diag_gpu.py gemmprof > gpurun_out/r2z_ncu_gemm.log 2>&1; tail -2 gpurun_out/r2z_ncu_gemm.log
ncu -i gpurun_out/r2z_gemm.ncu-rep --page raw --csv > gpurun_out/r2z_gemm_raw.csv 2>/dev/null; wc -c gpurun_out/r2z_gemm_raw.csv
echo "== ncu full of render_with_sg forward / backward"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:sg_render -c 2 -o gpurun_out/r2z_sg python tools/profile_shading.py > gpurun_out/r2z_ncu_sg.log 2>&1; tail -2 gpurun_out/r2z_ncu_sg.log
ncu -i gpurun_out/r2z_sg.ncu-rep --page raw --csv > gpurun_out/r2z_sg_raw.csv 2>/dev/null; wc -c gpurun_out/r2z_sg_raw.csv
rm -f gpurun_out/r2z_gemm.ncu-rep gpurun_out/r2z_sg.ncu-rep
