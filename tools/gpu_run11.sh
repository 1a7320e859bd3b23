#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python tools/prof_synth.py e1c_8prn_20s_clean 5 > gpurun_out/prof_synth_plain.log 2>&1; cat gpurun_out/prof_synth_plain.log
python tools/prof_synth.py e1c_prn3_20s_withdoppler 5 2>&1 | tail -1
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/per_launches.csv python tools/prof_synth.py e1c_8prn_20s_clean 1 > gpurun_out/ncu_per_l.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'k_synth_periodic' -s 2 -c 1 -o gpurun_out/prof_per -f python tools/prof_synth.py e1c_8prn_20s_clean 1 > gpurun_out/ncu_per.log 2>&1; echo "ncu exit $?"
