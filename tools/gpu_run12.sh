#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python tools/prof_synth.py e1c_8prn_20s_clean 5 2>&1 | tail -1
python tools/prof_synth.py e1c_prn3_20s_withdoppler 5 2>&1 | tail -1
timeout 900 python -m pytest tests/test_gpu_synth_periodic.py -x -q -m gpu 2>&1 | tail -5
