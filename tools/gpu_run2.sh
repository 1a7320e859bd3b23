mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -8 | tee gpurun_out/pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3 | tee gpurun_out/smoke.log
R4WB_BENCH_DEBUG=1 timeout 900 python bench.py --no-cpu-baseline > gpurun_out/bench_r2b.json 2> gpurun_out/bench_r2b.err; echo "bench rc=$?"; tail -c 600 gpurun_out/bench_r2b.err
python tools/prof_position.py 0 270 530 590 2>&1 | tee gpurun_out/position_r2b.log
timeout 900 ncu --set full --clock-control none --import-source on --kernel-name-base function -k "regex:k_synth_lat" -c 2 -f -o gpurun_out/prof_lat_r2b python tools/prof_position.py 530 > gpurun_out/ncu_lat_r2b.log 2>&1; echo "ncu rc=$?"
