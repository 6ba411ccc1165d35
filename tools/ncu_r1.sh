#!/bin/bash
# Round-1 ncu evidence (run under gpurun): launch list of two bench steps at batch 16 and --set full captures of the hot kernels.
set -x
ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/launches_r1_v12.csv python bench.py --batch 16 --steps 2 --warmup 1 --no-cpu-baseline --no-e2e --no-rollout > gpurun_out/ncu_bench_v12.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:conv3d_tc_kernel -s 2 -c 1 -o gpurun_out/prof_conv64_pair -f python tools/run_kernels_once.py conv 16 > gpurun_out/ncu_conv_pair.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:linattn -s 3 -c 3 -o gpurun_out/prof_linblock -f python tools/run_kernels_once.py linblock 16 > gpurun_out/ncu_linblock.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:stem_conv -s 1 -c 1 -o gpurun_out/prof_stem -f python tools/run_kernels_once.py stem 16 > gpurun_out/ncu_stem.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:temporal_block -s 1 -c 1 -o gpurun_out/prof_tblock_v2 -f python tools/run_kernels_once.py tblock 16 > gpurun_out/ncu_tblock_v2.log 2>&1
ls -la gpurun_out/*.ncu-rep
