#!/bin/bash
cd "$(dirname "$0")/.."
timeout 900 python -m pytest tests/test_gpu_synth_periodic.py -q -m gpu -k other_sample_rates 2>&1 | grep -E "^E  |assert|path|FAILED|passed|failed" | head -60
