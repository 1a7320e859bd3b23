mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_synth.py -x -q -m gpu -k "whole_600s" 2>&1 | tail -6
