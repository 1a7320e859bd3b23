mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_synth.py -x -q -m gpu -k "block_view or block_loop" 2>&1 | tail -8 | tee gpurun_out/pytest_view.log
timeout 1200 python bench.py > gpurun_out/bench_s.json 2> gpurun_out/bench_s.err; echo "bench rc=$?"; tail -c 300 gpurun_out/bench_s.err
python - <<'P'
import json
d=json.loads(open('gpurun_out/bench_s.json').read().strip().splitlines()[-1])
print(json.dumps(d.get('e2e_block_api'),indent=1)); print(d['value'], d['e2e'])
P
