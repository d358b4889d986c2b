timeout 1200 python -m pytest tests/test_esbr_hfgen_gpu.py tests/test_esbr_stage_gpu.py tests/test_dropin_gpu.py -x -q -m gpu 2>&1 | tail -6
