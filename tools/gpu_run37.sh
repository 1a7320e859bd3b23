#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python bench.py > gpurun_out/bench_default.log 2> gpurun_out/bench_default.err; echo "bench exit $?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_default.log').read().strip().splitlines()[-1])
print('synth', round(d['value']), d['ms_per_step'], d['roofline']['frac'], 'e2e', round(d['e2e']['value']), 'acq', d['acq']['value']/1e9, 'acq e2e', d['acq']['e2e']['value']/1e9, 'track', d['track'])
PY
