mkdir -p gpurun_out
python tools/prof_position.py 0 530 2>&1 | tee gpurun_out/position_r2m.log
timeout 900 python -m pytest tests/test_gpu_synth.py -x -q -m gpu -k "lattice or random_access or block_loop or clean_iq or integer_sink or device_output" 2>&1 | tail -4
