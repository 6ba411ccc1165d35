#!/bin/bash
export DPC_ROLLOUT_PROF=1
for c in 2 4 8; do echo "CS=$c"; DPC_ROLLOUT_CLUSTER=$c timeout 300 python tools/time_rollout.py 16 32 2>&1 | grep -E "frame phases|rank 0|rollout B" | tail -3; done
