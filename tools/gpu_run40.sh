#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python tools/prof_new_kernels.py > gpurun_out/new_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:'k_track|k_synth_direct|k_compose' -c 3 -o gpurun_out/prof_new -f python tools/prof_new_kernels.py > gpurun_out/ncu_new.log 2>&1; echo "ncu exit $?"; tail -2 gpurun_out/new_plain.log
