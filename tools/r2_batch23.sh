#!/bin/bash
timeout 900 python -m pytest tests/test_unet_gpu.py tests/test_metric_shape_gpu.py tests/test_jellyfish_nets.py -m gpu -q 2>&1 | tail -3
timeout 600 python tools/profile_step.py 64 2>&1 | grep -E "spatial_linear_block|temporal_block|total"
DPC_SL_PIPE=0 timeout 600 python tools/profile_step.py 64 2>&1 | grep -E "spatial_linear_block|total"
