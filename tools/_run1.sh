timeout 1500 python -m pytest tests/test_envcalc_gpu.py tests/test_hfgen_gpu.py tests/test_peaklim_gpu.py tests/test_sbrdec_gpu.py tests/test_chain_gpu.py tests/test_dropin_gpu.py -x -q -m gpu 2>&1 | tail -3
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-extra-stages --e2e-steps 1 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); k=d['kernels']; print('chain', d['ms_per_step'], d['value'], 'envcalc', k['calc_sbrenvelope_hq_kernel']['launch_ms'], 'hfgen', k['hf_generator_hq_kernel']['launch_ms'])"
timeout 300 python bench.py --workload aac_lc_stereo_output --steps 10 --warmup 3 --no-cpu-baseline --no-extra-stages --e2e-steps 1 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); k=d['kernels']; print('lcout', d['ms_per_step'], {n:round(x['launch_ms'],3) for n,x in k.items()})"
