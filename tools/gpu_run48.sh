cd "$(dirname "$0")/.."
timeout 600 python -m pytest tests/test_gpu_acq.py tests/test_cabi.py tests/test_sink.py -q -m gpu -x 2>&1 | tail -4
bash tools/gpu_run47.sh
