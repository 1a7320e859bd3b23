#!/bin/bash
mkdir -p gpurun_out
for k in 10 5; do
  R4WB_SYNTH_TILE_K=$k python bench.py --steps 3 --warmup 3 --acq-snapshots 32 --no-cpu-baseline > gpurun_out/bench_k$k.log 2> gpurun_out/bench_k$k.err; echo "bench k$k exit $?"
  python -c "
import json
d=json.loads(open('gpurun_out/bench_k$k.log').read().strip().splitlines()[-1]); print('k$k', round(d['value']))"
done
