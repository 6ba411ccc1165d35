#!/bin/bash
timeout 900 python -m pytest tests/test_kernels_gpu.py -m gpu -q -k "temporal_attention" 2>&1 | tail -3
timeout 600 python bench.py --config smoke128x64-ddim --steps 3 --warmup 2 --profile 2>&1 | grep -E " ms .*temporal_attention|ms_per_step" | cut -c1-220 | head -4
timeout 600 python tools/profile_step.py 64 2>&1 | grep -E "temporal_attention|total"
