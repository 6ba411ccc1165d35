#!/bin/bash
timeout 600 python bench.py --config burgers --steps 1 --warmup 1 --no-cuda-graph --profile 2>&1 >/dev/null | grep -E " ms |total" | head -40
