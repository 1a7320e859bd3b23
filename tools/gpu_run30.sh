cd "$(dirname "$0")/.."; timeout 1200 python -m pytest tests/test_gpu_synth.py -q -m gpu -k "600s" 2>&1 | tail -8
