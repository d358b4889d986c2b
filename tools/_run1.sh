timeout 900 python -m pytest tests/test_sbr_sideinfo_gpu.py -x -q -m gpu > gpurun_out/sd_test.log 2>&1
head -60 gpurun_out/sd_test.log
