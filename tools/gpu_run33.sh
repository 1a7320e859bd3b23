#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:'^k_synth$' -s 2 -c 1 -o gpurun_out/prof_gen -f python tools/prof_synth.py e1c_8prn_60s_cn34_orbital 1 > gpurun_out/ncu_gen.log 2>&1; echo "ncu exit $?"
