#!/bin/bash
cd "$(dirname "$0")/.."
timeout 900 python -m pytest tests/test_gpu_synth.py -q -m gpu -k "other_signals or unsupported" 2>&1 | tail -3
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:k_synth_direct python tools/prof_new_kernels.py 2>&1 | grep -E "gpu__time" | tail -2
