#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
timeout 600 python bench.py --steps 3 --warmup 3 --acq-snapshots 296 > gpurun_out/b1.log 2> gpurun_out/b1.err; echo "rc=$?"; tail -3 gpurun_out/b1.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/b1.log').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['roofline']['frac'], d['e2e']['value'], d.get('track'))
PY
