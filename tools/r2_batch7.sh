#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py -m gpu -q -k "temporal_block" 2>&1 | tail -4
for d in 0 1 2; do echo "DBG=$d"; DPC_TB_DBG=$d timeout 120 python tools/time_temporal_block.py 16 2>&1 | tail -1; done
timeout 120 python tools/time_temporal_block.py 64 2>&1 | tail -1
