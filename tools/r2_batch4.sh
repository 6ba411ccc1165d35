#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_sampler_gpu.py tests/test_evaluate_gpu.py tests/test_distributed_gpu.py -m gpu -q > gpurun_out/r2_pytest_4.log 2>&1; tail -6 gpurun_out/r2_pytest_4.log
run() { name=$1; shift; python bench.py "$@" > gpurun_out/r2_bench_$name.json 2> gpurun_out/r2_bench_$name.err; echo "$name rc=$?"; }
run graph_b64 --steps 10 --warmup 3 --cuda-graph --no-cpu-baseline --no-rollout --no-e2e --no-roofline
run graph_b8 --steps 10 --warmup 3 --cuda-graph --batch 8 --no-cpu-baseline --no-rollout --no-e2e --no-roofline
run graph_b16 --steps 10 --warmup 3 --cuda-graph --batch 16 --no-cpu-baseline --no-rollout --no-e2e --no-roofline
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r2_bench_*_b*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1]); print(f, round(d['ms_per_step'],3), round(d['value'],4), d['gpu_launches'], d['config'].get('graph_launches'))
    except Exception as e: print(f,'ERR',e)
PY
tail -3 gpurun_out/r2_bench_graph_b8.err
