#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_gpu_synth.py -m gpu -q -x > gpurun_out/pytest_gpu_synth.log 2>&1; echo "pytest synth exit $?"; tail -3 gpurun_out/pytest_gpu_synth.log
for v in lut nolut; do
  if [ $v = nolut ]; then export R4WB_SYNTH_NO_LUT=1; fi
  python bench.py --steps 3 --warmup 3 --acq-snapshots 32 --no-cpu-baseline > gpurun_out/bench_$v.log 2> gpurun_out/bench_$v.err; echo "bench $v exit $?"
done
unset R4WB_SYNTH_NO_LUT
python - <<'PY'
import json
for f in ("gpurun_out/bench_lut.log","gpurun_out/bench_nolut.log"):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, "synth Ms/s", round(d["value"]), "frac", round(d["roofline"]["frac"],4))
    except Exception as e: print(f, "ERR", e)
PY
python bench.py --steps 1 --warmup 1 --acq-snapshots 32 --no-cpu-baseline > gpurun_out/plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:k_synth -s 1 -c 1 -o gpurun_out/prof_synth -f python bench.py --steps 1 --warmup 1 --acq-snapshots 32 --no-cpu-baseline > gpurun_out/ncu_synth.log 2>&1; echo "ncu exit $?"
