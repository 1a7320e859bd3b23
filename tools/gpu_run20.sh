#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_synth.py tests/test_gpu_synth_periodic.py -q -m gpu 2>&1 | tail -15
python tools/prof_synth.py e1c_8prn_20s_clean 3 2>&1 | tail -1
R4WB_SYNTH_PERIODIC=0 python tools/prof_synth.py e1c_8prn_20s_clean 3 2>&1 | tail -1
python tools/prof_synth.py e1c_8prn_60s_cn34_orbital 3 2>&1 | tail -1
