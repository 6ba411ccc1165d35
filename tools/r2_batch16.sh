#!/bin/bash
timeout 600 python -m pytest tests/test_kernels_gpu.py -m gpu -q -k "temporal_block" 2>&1 | tail -3
DPC_TB_DBG=8 timeout 120 python tools/time_temporal_block.py 16 2>&1 | grep -E "pipe head|fused" | tail -5
for pz in 1 0; do echo "PIPE=$pz"; DPC_TB_PIPE=$pz timeout 120 python tools/time_temporal_block.py 64 2>&1 | tail -1; done
