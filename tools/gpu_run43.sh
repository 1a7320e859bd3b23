cd "$(dirname "$0")/.."; timeout 900 python -m pytest tests/test_gpu_synth.py -q -m gpu -k "doppler_rate or antenna" 2>&1 | tail -6
