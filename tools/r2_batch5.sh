#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py -m gpu -q -k "temporal_block" > gpurun_out/r2_pytest_5.log 2>&1; tail -8 gpurun_out/r2_pytest_5.log
timeout 300 python tools/time_temporal_block.py 64 2>&1 | tail -3
timeout 300 python tools/time_temporal_block.py 8 2>&1 | tail -3
timeout 600 python -m pytest tests/test_burgers_sampler.py -m gpu -q > gpurun_out/r2_pytest_5b.log 2>&1; tail -8 gpurun_out/r2_pytest_5b.log
timeout 600 python bench.py --config burgers --steps 3 --warmup 1 > gpurun_out/r2_bench_burgers_graph.json 2> gpurun_out/r2_bench_burgers_graph.err; tail -c 600 gpurun_out/r2_bench_burgers_graph.json; tail -3 gpurun_out/r2_bench_burgers_graph.err
