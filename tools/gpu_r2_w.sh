#!/bin/bash
# late round-2 validation: whole GPU suite, smoke, the default bench line, ncu launch list of bench.py itself
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
echo "== whole gpu suite"
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -5
echo "== smoke"
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
echo "== bench (default flags)"
timeout 900 python bench.py > gpurun_out/r2w_bench.json 2> gpurun_out/r2w_bench.err; tail -c 6000 gpurun_out/r2w_bench.json; grep -v "Warn\|warn" gpurun_out/r2w_bench.err | tail -3
echo "== reference arm"
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2w_ref.json 2> gpurun_out/r2w_ref.err; tail -c 1500 gpurun_out/r2w_ref.json
