#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python bench.py --steps 3 --warmup 3 --acq-snapshots 296 --no-cpu-baseline > gpurun_out/b1.log 2> gpurun_out/b1.err; echo "rc=$?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/b1.log').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['roofline'], d['synth_kernel_ms'], d['synth_kernel_launches'])
print(d['acq']['roofline'])
PY
tail -3 gpurun_out/b1.err
