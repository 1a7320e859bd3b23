#!/bin/bash
# record run (round 1, session 4): full GPU test suite, default bench, reference arm, launch list, full ncu of the hot kernels
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?"; tail -3 gpurun_out/pytest_gpu.log
python bench.py > gpurun_out/bench_default.log 2> gpurun_out/bench_default.err; echo "bench exit $?"
python bench.py --impl reference > gpurun_out/bench_ref.log 2> gpurun_out/bench_ref.err; echo "ref exit $?"
python bench.py --steps 1 --warmup 1 --acq-snapshots 296 --no-cpu-baseline > gpurun_out/plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv python bench.py --steps 1 --warmup 1 --acq-snapshots 296 --no-cpu-baseline > gpurun_out/ncu_launches.log 2>&1; echo "ncu launches exit $?"
ncu --set full --clock-control none --import-source on -k regex:'k_synth_periodic|k_periodic_fix|k_rf_fwd|k_rf_inv_peak_tm' -s 4 -c 4 -o gpurun_out/prof_all -f python bench.py --steps 1 --warmup 1 --acq-snapshots 296 --no-cpu-baseline > gpurun_out/ncu_all.log 2>&1; echo "ncu full exit $?"
