#!/bin/bash
mkdir -p gpurun_out
( time python bench.py ) > gpurun_out/bench_default.log 2> gpurun_out/bench_default.err; echo "bench exit $?"; tail -3 gpurun_out/bench_default.err
( time python bench.py --impl reference ) > gpurun_out/bench_ref.log 2> gpurun_out/bench_ref.err; echo "ref exit $?"; tail -3 gpurun_out/bench_ref.err
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke exit $?"; tail -2 gpurun_out/smoke.log
nproc; lscpu | grep "Model name"
