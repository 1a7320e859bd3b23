cd "$(dirname "$0")/.."; timeout 1200 python -m pytest tests/test_gpu_acq.py -q -m gpu -k "all_prns" 2>&1 | tail -4
