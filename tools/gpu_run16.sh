#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python tools/prof_synth.py e1c_8prn_20s_clean 5 2>&1 | tail -1
python tools/prof_synth.py e1c_prn3_20s_withdoppler 5 2>&1 | tail -1
timeout 600 python -m pytest tests/test_gpu_synth.py tests/test_gpu_synth_periodic.py -q -m gpu -x 2>&1 | tail -3
timeout 600 python bench.py --steps 3 --warmup 3 --acq-snapshots 296 --no-cpu-baseline > gpurun_out/b1.log 2> gpurun_out/b1.err; echo "rc=$?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/b1.log').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['roofline']['frac'], d['synth_kernel_ms'], d['e2e']['value'])
PY
