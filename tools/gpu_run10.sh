#!/bin/bash
# periodic synthesis path: parity + A/B, then a short bench
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_synth_periodic.py -x -q -m gpu > gpurun_out/pytest_periodic.log 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest_periodic.log
tail -30 gpurun_out/pytest_periodic.log
timeout 600 python bench.py --steps 3 --warmup 3 --acq-snapshots 296 --no-cpu-baseline > gpurun_out/bench_per.log 2> gpurun_out/bench_per.err
tail -3 gpurun_out/bench_per.log; tail -5 gpurun_out/bench_per.err
R4WB_SYNTH_PERIODIC=0 timeout 600 python bench.py --steps 3 --warmup 3 --acq-snapshots 296 --no-cpu-baseline > gpurun_out/bench_noper.log 2> gpurun_out/bench_noper.err
tail -3 gpurun_out/bench_noper.log
