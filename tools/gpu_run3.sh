#!/bin/bash
mkdir -p gpurun_out
python bench.py --steps 1 --warmup 1 --acq-snapshots 32 --no-cpu-baseline > gpurun_out/plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:'k_rf_inv_peak' -s 1 -c 1 -o gpurun_out/prof_acq -f python bench.py --steps 1 --warmup 1 --acq-snapshots 32 --no-cpu-baseline > gpurun_out/ncu_acq.log 2>&1; echo "ncu exit $?"
