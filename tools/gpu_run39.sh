#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/bench_n2.log 2> gpurun_out/bench_n2.err; echo "N=2 rc=$?"
python - <<PY
import json
d=json.loads(open('gpurun_out/bench_n2.log').read().strip().splitlines()[-1])
print('N=2', d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'], 'acq', d['acq']['value'], d['acq']['e2e']['value'], d['avg_power'], d.get('track',{}).get('value'))
PY
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29534 bench.py --impl reference --gpus 2 --steps 1 --warmup 0 2>/dev/null | tail -1 | cut -c1-200
