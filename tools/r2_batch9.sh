#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_rollout_gpu.py tests/test_evaluate_gpu.py -m gpu -q 2>&1 | tail -5
for c in 2 4; do DPC_ROLLOUT_CLUSTER=$c timeout 900 python -m pytest tests/test_rollout_gpu.py -m gpu -q 2>&1 | tail -2; done
DPC_ROLLOUT_CLUSTER=2 DPC_ROLLOUT_T2=512 timeout 900 python -m pytest tests/test_rollout_gpu.py -m gpu -q 2>&1 | tail -2
timeout 300 python tools/time_rollout.py 16 64
timeout 300 python tools/time_rollout.py 32 64
timeout 300 python tools/time_rollout.py 64 64
DPC_ROLLOUT_T2=512 timeout 300 python tools/time_rollout.py 64 64
DPC_ROLLOUT_CLUSTER=4 timeout 300 python tools/time_rollout.py 64 64
DPC_ROLLOUT_CLUSTER=8 timeout 300 python tools/time_rollout.py 64 64
