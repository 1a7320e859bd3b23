mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_multi_device.py -x -q -m gpu 2>&1 | tail -40 | tee gpurun_out/pytest_multi.log
