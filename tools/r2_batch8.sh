#!/bin/bash
mkdir -p gpurun_out
for d in 0 4; do DPC_TC_DEBUG=$d timeout 200 python tools/time_conv.py 16 all 2>&1 | grep TFLOP; done
timeout 600 python tools/profile_step.py 64 > gpurun_out/r2_step_profile_v2.txt 2>&1; head -16 gpurun_out/r2_step_profile_v2.txt; tail -1 gpurun_out/r2_step_profile_v2.txt
timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-rollout > gpurun_out/r2_bench_v2.json 2> gpurun_out/r2_bench_v2.err; python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2_bench_v2.json').read().strip().splitlines()[-1]); print(d['ms_per_step'], d['value'], d['e2e'], d['roofline']['achieved'])
PY
