mkdir -p gpurun_out
for mix in 0 1; do echo "mix $mix"; R4WB_LAT_MIX_NOISE=$mix python tools/prof_position.py 0 530 2>&1 | tee -a gpurun_out/mix_r2l.log; done
R4WB_LAT_MIX_NOISE=1 timeout 600 ncu --set full --clock-control none --import-source on --kernel-name-base function -k "regex:k_synth_lat" -c 1 -f -o gpurun_out/prof_lat_r2l python tools/prof_position.py 530 > gpurun_out/ncu_lat_r2l.log 2>&1; echo "ncu rc=$?"
timeout 900 python -m pytest tests/test_gpu_synth.py -x -q -m gpu -k "lattice or random_access or block_loop or clean_iq or cn0" 2>&1 | tail -4
