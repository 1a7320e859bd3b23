#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_gpu_acq.py -m gpu -q -x > gpurun_out/pytest_gpu_acq.log 2>&1; echo "pytest acq exit $?"; tail -15 gpurun_out/pytest_gpu_acq.log
python bench.py --steps 2 --warmup 2 --acq-snapshots 256 --no-cpu-baseline > gpurun_out/bench_rf.log 2> gpurun_out/bench_rf.err; echo "bench rf exit $?"
R4WB_ACQ_ENGINE=smem python bench.py --steps 2 --warmup 2 --acq-snapshots 256 --no-cpu-baseline > gpurun_out/bench_smem.log 2> gpurun_out/bench_smem.err; echo "bench smem exit $?"
python - <<'PY'
import json
for f in ("gpurun_out/bench_rf.log","gpurun_out/bench_smem.log"):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, "synth Ms/s", round(d["value"]), "acq", d["acq"]["value"]/1e9, d["acq"]["kernel_ms"], "guards", d["acq"]["f64_guard_reruns"], d["acq"]["first_snapshot"])
    except Exception as e: print(f, "ERR", e)
PY
