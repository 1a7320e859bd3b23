#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_synth.py tests/test_gpu_synth_periodic.py -q -m gpu -x 2>&1 | tail -3
timeout 600 python bench.py --steps 3 --warmup 3 --acq-snapshots 296 --no-cpu-baseline 2> gpurun_out/b1.err | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('N=1', d['value'], d['ms_per_step'], d['roofline']['frac'], d['e2e']['value'], d['acq']['value'])"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 3 --warmup 3 --acq-snapshots 296 --no-cpu-baseline 2> gpurun_out/b2.err | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('N=2', d['value'], d['ms_per_step'], d['roofline']['frac'], d['e2e']['value'], d['acq']['value'])"
tail -3 gpurun_out/b2.err
