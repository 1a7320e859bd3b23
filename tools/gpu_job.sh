mkdir -p gpurun_out
FLAGS="--steps 2 --warmup 1 --no-per-config --no-cpu-baseline --no-block-api --acq-snapshots 296"
python bench.py $FLAGS > gpurun_out/plain_r2i.json 2> gpurun_out/plain_r2i.err && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/launches_r2i.csv python bench.py $FLAGS > gpurun_out/ncu_launches_r2i.log 2>&1; echo "launch list rc=$?"; wc -l gpurun_out/launches_r2i.csv
timeout 900 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k "regex:k_synth_lat|k_rf_inv_peak_tm" -c 12 --csv --log-file gpurun_out/traffic_r2i.csv python bench.py $FLAGS --no-parity > gpurun_out/ncu_traffic_r2i.log 2>&1; echo "traffic rc=$?"; tail -4 gpurun_out/traffic_r2i.csv
timeout 900 python -m pytest tests/test_sink.py tests/test_gpu_synth.py -x -q -m gpu 2>&1 | tail -4
