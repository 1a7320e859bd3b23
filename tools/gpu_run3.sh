mkdir -p gpurun_out
python tools/prof_position.py 0 530 2>&1 | tee gpurun_out/position_r2c.log
python tools/prof_prologue.py > gpurun_out/prologue_r2c.log 2>&1; cat gpurun_out/prologue_r2c.log
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/prologue_launches_r2c.csv python tools/prof_prologue.py > gpurun_out/prologue_ncu_r2c.log 2>&1; echo "ncu rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on --kernel-name-base function -k "regex:k_synth_lat" -c 1 -f -o gpurun_out/prof_lat_r2c python tools/prof_position.py 530 > gpurun_out/ncu_lat_r2c.log 2>&1; echo "ncu rc=$?"
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -8 | tee gpurun_out/pytest_gpu.log
