#!/bin/bash
timeout 900 python -m pytest tests/test_burgers_sampler.py -m gpu -q 2>&1 | tail -4
timeout 600 python bench.py --config burgers --steps 3 --warmup 1 > gpurun_out/r2_bench_burgers_tc2.json 2> gpurun_out/r2_bench_burgers_tc2.err; python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2_bench_burgers_tc2.json').read().strip().splitlines()[-1]); print(d['ms_per_step'], d['value'], d['gpu_launches'])
PY
tail -3 gpurun_out/r2_bench_burgers_tc2.err
