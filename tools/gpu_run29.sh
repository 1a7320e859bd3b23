#!/bin/bash
cd "$(dirname "$0")/.."
for i in 1 2; do
python bench.py > gpurun_out/bd$i.log 2> gpurun_out/bd$i.err
python - <<PY
import json
d=json.loads(open('gpurun_out/bd$i.log').read().strip().splitlines()[-1])
print('default run $i: synth', round(d['value']), 'e2e', round(d['e2e']['value']), 'acq', d['acq']['value']/1e9, 'acq e2e', d['acq']['e2e']['value']/1e9, 'guards', d['acq']['f64_guard_reruns'])
PY
done
python bench.py --no-cpu-baseline --no-track > gpurun_out/bd3.log 2> gpurun_out/bd3.err
python - <<PY
import json
d=json.loads(open('gpurun_out/bd3.log').read().strip().splitlines()[-1])
print('no-cpu no-track: synth', round(d['value']), 'e2e', round(d['e2e']['value']), 'acq', d['acq']['value']/1e9, 'acq e2e', d['acq']['e2e']['value']/1e9)
PY
