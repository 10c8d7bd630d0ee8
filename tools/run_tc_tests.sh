set -x
timeout 600 python -m pytest tests/test_gpu_conv_cond_tc.py -x -q -m gpu 2>&1 | tail -30
