mkdir -p gpurun_out
python tools/prof_position.py 0 530 2>&1 | tee gpurun_out/position_r2e.log
timeout 600 ncu --set full --clock-control none --import-source on --kernel-name-base function -k "regex:k_synth_lat" -c 1 -f -o gpurun_out/prof_lat_r2e python tools/prof_position.py 530 > gpurun_out/ncu_lat_r2e.log 2>&1; echo "ncu rc=$?"
timeout 1500 python -m pytest tests/test_gpu_synth.py tests/test_gpu_synth_periodic.py -x -q -m gpu 2>&1 | tail -5 | tee gpurun_out/pytest_gpu.log
