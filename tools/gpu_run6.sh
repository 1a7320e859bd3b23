mkdir -p gpurun_out
timeout 2400 python -m pytest tests -x -q -m gpu 2>&1 | tail -8 | tee gpurun_out/pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3 | tee gpurun_out/smoke.log
R4WB_BENCH_DEBUG=1 timeout 1200 python bench.py > gpurun_out/bench_r2f.json 2> gpurun_out/bench_r2f.err; echo "bench rc=$?"; tail -c 400 gpurun_out/bench_r2f.err
