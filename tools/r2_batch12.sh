#!/bin/bash
mkdir -p gpurun_out
timeout 2400 python -m pytest tests/ -q -m gpu -x > gpurun_out/r2_pytest_full.log 2>&1; tail -8 gpurun_out/r2_pytest_full.log
timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2_bench_v3.json 2> gpurun_out/r2_bench_v3.err; python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2_bench_v3.json').read().strip().splitlines()[-1]); print(d['ms_per_step'], d['value'], d['e2e']['value'], d['roofline']['achieved'], d['clocks'], d.get('rollout'))
PY
timeout 600 python tools/profile_step.py 64 > gpurun_out/r2_step_profile_v3.txt 2>&1; head -12 gpurun_out/r2_step_profile_v3.txt; tail -1 gpurun_out/r2_step_profile_v3.txt
