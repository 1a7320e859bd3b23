#!/bin/bash
cd "$(dirname "$0")/.."
timeout 900 python -m pytest tests/test_tracking.py -q -m gpu 2>&1 | tail -3
python tools/prof_track.py 296 500 2>&1 | grep channels | tail -1
python tools/prof_track.py 8 500 2>&1 | grep channels | tail -1
python tools/prof_track.py 1 500 2>&1 | grep channels | tail -1
python tools/prof_track.py 40 500 2>&1 | grep channels | tail -1
