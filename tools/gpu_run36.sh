#!/bin/bash
cd "$(dirname "$0")/.."
ncu --metrics gpu__time_duration.sum,launch__cluster_dim_x,launch__grid_size,sm__warps_active.avg.pct_of_peak_sustained_active --clock-control none -k regex:k_track python tools/prof_track.py 296 500 2>&1 | grep -E "k_track|gpu__time|cluster_dim_x|grid_size|warps_active" | tail -12
ncu --metrics gpu__time_duration.sum,launch__cluster_dim_x --clock-control none -k regex:k_track python tools/prof_track.py 8 500 2>&1 | grep -E "gpu__time|cluster_dim_x" | tail -4
