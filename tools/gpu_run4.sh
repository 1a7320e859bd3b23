mkdir -p gpurun_out
for ns in 0 3000 6000 9000; do echo "stagger $ns"; R4WB_LAT_STAGGER_NS=$ns python tools/prof_position.py 0 2>&1 | tee -a gpurun_out/stagger_r2d.log; done
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -8 | tee gpurun_out/pytest_gpu.log
R4WB_BENCH_DEBUG=1 timeout 900 python bench.py --no-cpu-baseline --no-per-config > gpurun_out/bench_r2d.json 2> gpurun_out/bench_r2d.err; echo "bench rc=$?"; tail -c 300 gpurun_out/bench_r2d.err
