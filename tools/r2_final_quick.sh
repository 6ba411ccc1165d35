#!/bin/bash
# refresh of the headline evidence after a kernel change: full GPU tests, smoke(), default bench, step profile, launch list
mkdir -p gpurun_out
timeout 1800 python -m pytest tests/ -q -m gpu > gpurun_out/r2f_pytest_gpu.log 2>&1; tail -2 gpurun_out/r2f_pytest_gpu.log
timeout 600 python __graft_entry__.py smoke 2>&1 | tail -1
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r2f_bench_default.json 2> gpurun_out/r2f_bench_default.err
timeout 900 python bench.py --config jellyfish128 --steps 5 --warmup 2 > gpurun_out/r2f_bench_jellyfish128.json 2> gpurun_out/r2f_bench_jellyfish128.err
timeout 900 python bench.py --config smoke128x64-ddim --steps 5 --warmup 2 > gpurun_out/r2f_bench_ddim128.json 2> gpurun_out/r2f_bench_ddim128.err
timeout 600 python tools/profile_step.py 64 > gpurun_out/r2f_step_profile.txt 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/r2f_launches_bench.csv python bench.py --steps 2 --warmup 1 --no-cuda-graph --no-cpu-baseline --no-e2e --no-rollout > gpurun_out/r2f_ncu_bench.log 2>&1
python tools/launch_list_summary.py gpurun_out/r2f_launches_bench.csv > gpurun_out/r2f_launches_bench.md; gzip -f gpurun_out/r2f_launches_bench.csv
ncu --set full --clock-control none -k regex:"linattn_apply_pipe|linattn_context_tc" -c 2 -o gpurun_out/r2f_full_linblock -f python tools/run_kernels_once.py linblock 8 > gpurun_out/r2f_ncu_linblock.log 2>&1
python tools/ncu_summary.py gpurun_out/r2f_full_linblock.ncu-rep > gpurun_out/r2f_full_linblock.txt 2>&1; rm -f gpurun_out/r2f_full_linblock.ncu-rep
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r2f_bench_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1]); print(f, round(d['ms_per_step'],3), round(d['value'],4), (d.get('e2e') or {}).get('value'), (d.get('roofline') or {}).get('frac'), (d.get('cpu_baseline') or {}).get('value'))
    except Exception as e: print(f,'ERR',e)
PY
head -8 gpurun_out/r2f_step_profile.txt; tail -1 gpurun_out/r2f_step_profile.txt
