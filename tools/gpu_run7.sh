mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_multi_device.py -x -q -m gpu 2>&1 | tail -5 | tee gpurun_out/pytest_multi.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --no-per-config > gpurun_out/bench_r2g_n2.json 2> gpurun_out/bench_r2g_n2.err; echo "bench rc=$?"; tail -c 300 gpurun_out/bench_r2g_n2.err
python tools/prof_position.py 0 530 2>&1 | tee gpurun_out/position_r2g.log
