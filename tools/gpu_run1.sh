mkdir -p gpurun_out
R4WB_BENCH_DEBUG=1 timeout 900 python bench.py > gpurun_out/bench_r2a.json 2> gpurun_out/bench_r2a.err; echo "bench rc=$?"; tail -c 300 gpurun_out/bench_r2a.err
python tools/prof_prologue.py > gpurun_out/prologue_plain.log 2>&1; cat gpurun_out/prologue_plain.log
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/prologue_launches.csv python tools/prof_prologue.py > gpurun_out/prologue_ncu.log 2>&1; echo "ncu rc=$?"
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -8 | tee gpurun_out/pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3 | tee gpurun_out/smoke.log
