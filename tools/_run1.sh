# scratch driver for one gpurun call: GPU parity tests of the HE-AACv2 chain + a short bench line with the per-kernel table
timeout 1500 python -m pytest tests/test_sbrdec_gpu.py tests/test_chain_gpu.py tests/test_dropin_gpu.py -x -q -m gpu 2>&1 | tail -3
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-extra-stages --e2e-steps 1 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); k=d['kernels']; print('chain', d['ms_per_step'], d['value']); print({n:round(v['launch_ms'],4) for n,v in k.items()})"
