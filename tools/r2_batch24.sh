#!/bin/bash
DPC_SL_DBG=1 timeout 120 python tools/time_linear_block.py 16 2>&1 | grep -E "linattn context|fused" | tail -3
