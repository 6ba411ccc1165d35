#!/bin/bash
# round-2 final evidence: full GPU test suite, every BASELINE.json configuration, the reference arm, ncu of the changed kernels
mkdir -p gpurun_out
T0=$(date +%s); lap() { echo "[lap] $1 $(( $(date +%s) - T0 )) s"; }
timeout 1800 python -m pytest tests/ -q -m gpu > gpurun_out/r2f_pytest_gpu.log 2>&1; tail -3 gpurun_out/r2f_pytest_gpu.log; lap pytest
timeout 600 python __graft_entry__.py smoke 2>&1 | tail -2; lap smoke
run() { name=$1; shift; timeout 900 python bench.py "$@" > gpurun_out/r2f_bench_$name.json 2> gpurun_out/r2f_bench_$name.err; echo "$name rc=$?"; }
run default --steps 10 --warmup 3; lap default
run reference --impl reference --steps 2 --warmup 1; lap reference
run smoke16 --config smoke16+rollout --steps 10 --warmup 3; lap smoke16
run jellyfish128 --config jellyfish128 --steps 5 --warmup 2; lap jelly
run burgers --config burgers --steps 2 --warmup 1; lap burgers
run ddim128 --config smoke128x64-ddim --steps 5 --warmup 2; lap ddim
run smoke32 --config smoke256x8 --steps 10 --warmup 3; lap smoke32
run 3xtf32 --precision 3xtf32 --steps 3 --warmup 2 --no-cpu-baseline --no-rollout; lap 3xtf32
timeout 600 python tools/profile_step.py 64 > gpurun_out/r2f_step_profile.txt 2>&1; lap profile
ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/r2f_launches_bench.csv python bench.py --steps 2 --warmup 1 --no-cuda-graph --no-cpu-baseline --no-e2e --no-rollout > gpurun_out/r2f_ncu_bench.log 2>&1
python tools/launch_list_summary.py gpurun_out/r2f_launches_bench.csv > gpurun_out/r2f_launches_bench.md; gzip -f gpurun_out/r2f_launches_bench.csv; lap launch-list
ncu --set full --clock-control none --import-source on -k regex:temporal_block -s 1 -c 1 -o gpurun_out/r2f_full_tblock -f python tools/run_kernels_once.py tblock 8 > gpurun_out/r2f_ncu_tblock.log 2>&1
python tools/ncu_summary.py gpurun_out/r2f_full_tblock.ncu-rep > gpurun_out/r2f_full_tblock.txt 2>&1; rm -f gpurun_out/r2f_full_tblock.ncu-rep
ncu --set full --clock-control none -k regex:smoke_rollout -s 1 -c 1 -o gpurun_out/r2f_full_rollout -f python tools/time_rollout.py 64 32 > gpurun_out/r2f_ncu_rollout.log 2>&1
ncu --set full --clock-control none -k regex:"guided_step|final_proj|layernorm_channels|time_proj|sinusoidal" -c 8 -o gpurun_out/r2f_full_stream2 -f python bench.py --scaling weak --no-cuda-graph --no-cpu-baseline --no-e2e --no-rollout --no-roofline --batch 8 --steps 1 --warmup 0 > gpurun_out/r2f_ncu_stream2.log 2>&1
python tools/ncu_summary.py gpurun_out/r2f_full_stream2.ncu-rep > gpurun_out/r2f_full_stream2.txt 2>&1; rm -f gpurun_out/r2f_full_stream2.ncu-rep
python tools/ncu_summary.py gpurun_out/r2f_full_rollout.ncu-rep > gpurun_out/r2f_full_rollout.txt 2>&1; rm -f gpurun_out/r2f_full_rollout.ncu-rep; lap ncu-full
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r2f_bench_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1]); print(f, round(d['ms_per_step'],3), round(d['value'],4), d.get('gpu_launches'), (d.get('e2e') or {}).get('value'), (d.get('roofline') or {}).get('frac'), (d.get('cpu_baseline') or {}).get('value'), (d.get('rollout') or {}).get('ms_total'))
    except Exception as e: print(f,'ERR',e)
PY
du -sh gpurun_out
