#!/bin/bash
# round-2 GPU trip A: tracer v2 (fixed schedule, then graph mode), full suite, full-size parity diagnostics, short bench
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
echo "== tracer tests, fixed schedule"
NEFII_TRACE_GRAPH=0 timeout 600 python -m pytest tests/test_tracer_gpu.py -x -q 2>&1 | tail -15
echo "== tracer tests, graph mode"
NEFII_TRACE_GRAPH=1 timeout 600 python -m pytest tests/test_tracer_gpu.py -x -q 2>&1 | tail -25
echo "== whole gpu suite (graph mode default)"
timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -40
echo "== fullsize parity diag"
FULL_TIERS="1,0;1,8;4,4" timeout 900 python tools/diag_gpu.py fullsize 2>&1 | grep -v Warning | tail -80
echo "== bench (fixed schedule)"
NEFII_TRACE_GRAPH=0 timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r2a_bench_static.json 2> gpurun_out/r2a_bench_static.err; tail -c 3000 gpurun_out/r2a_bench_static.json; tail -5 gpurun_out/r2a_bench_static.err
echo "== bench (graph mode)"
NEFII_TRACE_GRAPH=1 timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r2a_bench_graph.json 2> gpurun_out/r2a_bench_graph.err; tail -c 3000 gpurun_out/r2a_bench_graph.json; tail -5 gpurun_out/r2a_bench_graph.err
