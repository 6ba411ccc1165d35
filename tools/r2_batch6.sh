#!/bin/bash
mkdir -p gpurun_out
for d in 0 1 2 4 5 7; do echo "DBG=$d"; DPC_TB_DBG=$d timeout 120 python tools/time_temporal_block.py 16 2>&1 | tail -1; done
ncu --set full --clock-control none --import-source on -k regex:temporal_block -s 1 -c 1 -o gpurun_out/r2_full_tblock16 -f python tools/run_kernels_once.py tblock 8 > gpurun_out/r2_ncu_tblock16.log 2>&1
ncu -i gpurun_out/r2_full_tblock16.ncu-rep --page raw --csv > gpurun_out/r2_full_tblock16_raw.csv 2>/dev/null
ncu -i gpurun_out/r2_full_tblock16.ncu-rep --page source --csv > gpurun_out/r2_full_tblock16_source.csv 2>/dev/null
python tools/ncu_summary.py gpurun_out/r2_full_tblock16.ncu-rep
ls -la gpurun_out
timeout 600 python bench.py --config burgers --steps 3 --warmup 1 > gpurun_out/r2_bench_burgers_graph.json 2> gpurun_out/r2_bench_burgers_graph.err; tail -c 700 gpurun_out/r2_bench_burgers_graph.json; tail -3 gpurun_out/r2_bench_burgers_graph.err
