timeout 600 python -m pytest tests/test_chain_gpu.py tests/test_sbrdec_gpu.py -x -q -m gpu 2>&1 | tail -3
