#!/bin/bash
timeout 900 python -m pytest tests/test_kernels_gpu.py -m gpu -q -k "linear_block or spatial_linear" 2>&1 | tail -4
timeout 120 python tools/time_linear_block.py 64 2>&1 | tail -1
DPC_SL_DBG=1 timeout 120 python tools/time_linear_block.py 16 2>&1 | grep -E "linattn context|fused" | tail -2
