timeout 1200 python -m pytest tests/test_qmf_synth_gpu.py tests/test_chain_gpu.py tests/test_sbrdec_gpu.py tests/test_dropin_gpu.py -x -q -m gpu 2>&1 | tail -3
timeout 300 python bench.py --workload qmf_synth_hq --steps 20 --warmup 5 --no-cpu-baseline --no-extra-stages 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('synth', d['ms_per_step'], d['roofline']['frac'])"
