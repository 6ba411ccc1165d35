#!/bin/bash
timeout 600 python bench.py --config smoke128x64-ddim --steps 2 --warmup 2 --profile 2>&1 >/dev/null | grep -E " ms |total" | head -24
