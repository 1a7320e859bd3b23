#!/bin/bash
mkdir -p gpurun_out
timeout 240 python -m pytest tests/test_gpu_acq.py -m gpu -q -x > gpurun_out/pytest_gpu_acq.log 2>&1; echo "pytest acq exit $?"; tail -5 gpurun_out/pytest_gpu_acq.log
timeout 120 python bench.py --steps 2 --warmup 2 --acq-snapshots 1024 --no-cpu-baseline > gpurun_out/bench_tm.log 2> gpurun_out/bench_tm.err; echo "bench tmem exit $?"
R4WB_ACQ_TMEM=0 timeout 120 python bench.py --steps 2 --warmup 2 --acq-snapshots 1024 --no-cpu-baseline > gpurun_out/bench_notm.log 2> gpurun_out/bench_notm.err; echo "bench no-tmem exit $?"
python - <<'PY'
import json
for f in ("gpurun_out/bench_tm.log","gpurun_out/bench_notm.log"):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, "acq", d["acq"]["value"]/1e9, d["acq"]["kernel_ms"], "guards", d["acq"]["f64_guard_reruns"])
    except Exception as e: print(f, "ERR", e)
PY
nvidia-smi --query-gpu=name,memory.used --format=csv
