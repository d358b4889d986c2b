timeout 1500 python -m pytest tests/test_dropin_gpu.py -x -q -m gpu 2>&1 | tail -6
