#!/bin/bash
# round-2 first GPU call: parity tests at the benchmarked shape + baseline benches (micro-batch sweep, small per-GPU batch)
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest_1.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest_1.log
tail -5 gpurun_out/r2_pytest_1.log
for mb in 0 2 4 8; do
  python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-rollout --no-e2e --micro-batch $mb > gpurun_out/r2_bench_mb$mb.json 2> gpurun_out/r2_bench_mb$mb.err
done
for b in 8 16 32; do
  python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-rollout --no-e2e --batch $b > gpurun_out/r2_bench_b$b.json 2> gpurun_out/r2_bench_b$b.err
done
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r2_bench_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1]); print(f, d['ms_per_step'], d['value'], d['gpu_launches'])
    except Exception as e: print(f,'ERR',e)
PY
