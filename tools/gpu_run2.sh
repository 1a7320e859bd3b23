#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?"; tail -4 gpurun_out/pytest_gpu.log
python bench.py --steps 3 --warmup 3 --acq-snapshots 64 --no-cpu-baseline > gpurun_out/bench_k10.log 2> gpurun_out/bench_k10.err; echo "bench k10 exit $?"
R4WB_SYNTH_TILE_K=5 python bench.py --steps 3 --warmup 3 --acq-snapshots 64 --no-cpu-baseline > gpurun_out/bench_k5.log 2> gpurun_out/bench_k5.err; echo "bench k5 exit $?"
python - <<'PY'
import json
for f in ("gpurun_out/bench_k10.log","gpurun_out/bench_k5.log"):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, "synth Ms/s", round(d["value"]), "frac", round(d["roofline"]["frac"],4), "acq", d["acq"]["value"], d["acq"]["kernel_ms"], "e2e", d["e2e"]["value"])
    except Exception as e: print(f, "ERR", e)
PY
python bench.py --steps 1 --warmup 1 --acq-snapshots 32 --no-cpu-baseline > gpurun_out/plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:k_synth -s 1 -c 1 -o gpurun_out/prof_synth -f python bench.py --steps 1 --warmup 1 --acq-snapshots 32 --no-cpu-baseline > gpurun_out/ncu_synth.log 2>&1; echo "ncu exit $?"
