#!/bin/bash
for ts in 0 1 0 1; do echo -n "two_streams=$ts: "; DPC_TWO_STREAMS=$ts timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-rollout --no-e2e --no-roofline 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['ms_per_step'], d['value'], d['clocks']['sm_mhz'])"; done
