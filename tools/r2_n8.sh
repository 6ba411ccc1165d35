#!/bin/bash
mkdir -p gpurun_out
N=$1
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/r2f_bench_n$N.json 2> gpurun_out/r2f_bench_n$N.err
python - <<PY
import json
d=json.loads(open('gpurun_out/r2f_bench_n$N.json').read().strip().splitlines()[-1]); print(d['n_gpus'], d['ms_per_step'], d['value'], d['scaling'], d.get('weak'), d['clocks'])
PY
tail -3 gpurun_out/r2f_bench_n$N.err
