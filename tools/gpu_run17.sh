#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_tracking.py -x -q -m gpu 2>&1 | tail -30
