#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_burgers_unet.py tests/test_burgers_sampler.py tests/test_burgers.py -m gpu -q 2>&1 | tail -12
timeout 600 python bench.py --config burgers --steps 3 --warmup 1 > gpurun_out/r2_bench_burgers_tc.json 2> gpurun_out/r2_bench_burgers_tc.err; tail -c 900 gpurun_out/r2_bench_burgers_tc.json | head -c 700; echo; tail -3 gpurun_out/r2_bench_burgers_tc.err
timeout 600 python bench.py --config burgers --steps 1 --warmup 1 --no-cuda-graph --profile 2>&1 >/dev/null | grep -E " ms |total" | head -16
