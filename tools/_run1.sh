timeout 1200 python -m pytest tests/test_sbrdec_lp_gpu.py tests/test_dropin_gpu.py -x -q -m gpu 2>&1 | tail -4
timeout 300 python bench.py --workload heaacv1_stereo_chain --steps 10 --warmup 3 --no-cpu-baseline --e2e-steps 2 2>gpurun_out/ab_x.err | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('heaacv1', round(d['ms_per_step'],3), round(d['value']/1e6,3), 'e2e', round(d['e2e']['value']/1e6,3)); print({n:round(v['launch_ms'],4) for n,v in d['kernels'].items()})" || tail -3 gpurun_out/ab_x.err
