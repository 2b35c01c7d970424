#!/bin/bash
# 8-GPU step with the cross-step prefetch, both placements: usage gpu_r2_multi3.sh N
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
N=$1
export PYTHONUNBUFFERED=1
for MODE in 1 2; do
NEFII_BENCH_PREFETCH=$MODE timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2951$MODE bench.py --gpus $N --steps 10 --warmup 5 --no-extras --no-cpu-baseline > gpurun_out/r2x_prefetch${MODE}_n$N.json 2> gpurun_out/r2x_prefetch${MODE}_n$N.err
python - gpurun_out/r2x_prefetch${MODE}_n$N.json <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1], d['n_gpus'], 'ms/step %.2f value %.3f M  e2e %.2f ms %.3f M captures %s' % (d['ms_per_step'], d['value']/1e6, d['e2e']['ms_per_step'], d['e2e']['value']/1e6, d['trace_graph_captures_in_timed_region']))
except Exception as e:
    print("failed", e)
PY
grep -v "Warn\|warn" gpurun_out/r2x_prefetch${MODE}_n$N.err | tail -3
done
