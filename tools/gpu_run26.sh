#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_acq.py -q -m gpu 2>&1 | tail -3
timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-track > gpurun_out/b1.log 2> gpurun_out/b1.err; echo "rc=$?"; tail -3 gpurun_out/b1.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/b1.log').read().strip().splitlines()[-1])
print(d['value'], d['e2e']['value'], 'acq', d['acq']['value'], 'acq e2e', d['acq']['e2e']['value'])
PY
