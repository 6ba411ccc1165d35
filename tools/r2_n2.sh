#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_distributed_gpu.py -m gpu -q 2>&1 | tail -3
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r2f_bench_n2.json 2> gpurun_out/r2f_bench_n2.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2f_bench_n2.json').read().strip().splitlines()[-1]); print(d['ms_per_step'], d['value'], d['n_gpus'], d['scaling'], d.get('weak'), d.get('gather'))
PY
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 1 --warmup 1 2>/dev/null | tail -c 400
