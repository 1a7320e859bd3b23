cd "$(dirname "$0")/.."
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -5
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
python bench.py > gpurun_out/bench_h_n1.json 2> gpurun_out/bench_h_n1.err; tail -c 600 gpurun_out/bench_h_n1.json
