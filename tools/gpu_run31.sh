#!/bin/bash
cd "$(dirname "$0")/.."
timeout 1200 python -m pytest tests/test_gpu_synth.py tests/test_gpu_synth_periodic.py -q -m gpu -k "not 600s" 2>&1 | tail -8
