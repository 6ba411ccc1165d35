#!/bin/bash
for d in 8; do echo "DBG=$d"; DPC_TB_DBG=$d timeout 120 python tools/time_temporal_block.py 16 2>&1 | grep "temporal block" | tail -5; done
