#!/bin/bash
timeout 900 python -m pytest tests/test_rollout_gpu.py -m gpu -q 2>&1 | tail -2
DPC_ROLLOUT_CLUSTER=2 timeout 900 python -m pytest tests/test_rollout_gpu.py -m gpu -q 2>&1 | tail -2
export DPC_ROLLOUT_PROF=1
timeout 300 python tools/time_rollout.py 16 32 2>&1 | tail -2
timeout 300 python tools/time_rollout.py 32 32 2>&1 | tail -2
timeout 300 python tools/time_rollout.py 64 32 2>&1 | tail -2
DPC_ROLLOUT_T2=512 timeout 300 python tools/time_rollout.py 64 32 2>&1 | tail -2
