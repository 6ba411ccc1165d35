#!/bin/bash
timeout 900 python -m pytest tests/test_kernels_gpu.py -m gpu -q -k "linear_block or spatial_linear" 2>&1 | tail -4
for pz in 1 0 1 0; do echo -n "PIPE=$pz "; DPC_SL_PIPE=$pz timeout 120 python tools/time_linear_block.py 64 2>&1 | tail -1; done
