#!/bin/bash
# round-2: every BASELINE.json configuration on one GPU + the default bench line
mkdir -p gpurun_out
run() { name=$1; shift; python bench.py "$@" > gpurun_out/r2_bench_$name.json 2> gpurun_out/r2_bench_$name.err; echo "$name rc=$?"; tail -c 600 gpurun_out/r2_bench_$name.json | head -c 400; echo; }
run default --steps 10 --warmup 3
run jellyfish128 --config jellyfish128 --steps 5 --warmup 2
run smoke16 --config smoke16+rollout --steps 10 --warmup 3
run burgers --config burgers --steps 2 --warmup 1
run ddim128 --config smoke128x64-ddim --steps 5 --warmup 2
run smoke32 --config smoke256x8 --steps 10 --warmup 3
run 3xtf32 --precision 3xtf32 --steps 3 --warmup 2 --no-cpu-baseline --no-rollout
python tools/time_jelly_nets.py 8 128 > gpurun_out/r2_time_jelly_nets.log 2>&1; head -3 gpurun_out/r2_time_jelly_nets.log
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r2_bench_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1]); print(f, round(d['ms_per_step'],3), round(d['value'],4), d['gpu_launches'], (d.get('e2e') or {}).get('value'), (d.get('roofline') or {}).get('frac'))
    except Exception as e: print(f,'ERR',e)
PY
