#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -x > gpurun_out/r2_pytest_3.log 2>&1; tail -6 gpurun_out/r2_pytest_3.log
run() { name=$1; shift; python bench.py "$@" > gpurun_out/r2_bench_$name.json 2> gpurun_out/r2_bench_$name.err; echo "$name rc=$?"; }
run ddim128_v2 --config smoke128x64-ddim --steps 5 --warmup 2 --profile
grep " ms " gpurun_out/r2_bench_ddim128_v2.err | head -8
run graph_b64 --steps 10 --warmup 3 --cuda-graph --no-cpu-baseline --no-rollout --no-e2e --no-roofline
run graph_b8 --steps 10 --warmup 3 --cuda-graph --batch 8 --no-cpu-baseline --no-rollout --no-e2e --no-roofline
run eager_b8 --steps 10 --warmup 3 --batch 8 --no-cpu-baseline --no-rollout --no-e2e --no-roofline
run graph_b16 --steps 10 --warmup 3 --cuda-graph --batch 16 --no-cpu-baseline --no-rollout --no-e2e --no-roofline
run eager_b16 --steps 10 --warmup 3 --batch 16 --no-cpu-baseline --no-rollout --no-e2e --no-roofline
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r2_bench_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1]); print(f, round(d['ms_per_step'],3), round(d['value'],4), d['gpu_launches'], d['config'].get('graph_launches'))
    except Exception as e: print(f,'ERR',e)
PY
