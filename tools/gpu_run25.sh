#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for N in 8 4; do
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2951$N bench.py --gpus $N --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_n$N.log 2> gpurun_out/bench_n$N.err; echo "N=$N rc=$?"
python - <<PY
import json
d=json.loads(open('gpurun_out/bench_n$N.log').read().strip().splitlines()[-1])
print('N=$N', d['value'], d['ms_per_step'], d['roofline']['frac'], 'e2e', d['e2e']['value'], 'acq', d['acq']['value'], d['acq']['e2e']['value'], d['clocks'])
PY
done
