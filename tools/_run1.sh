run() { timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-extra-stages --e2e-steps 1 2>gpurun_out/ab_x.err | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); k=d['kernels']; print('$1 chain', round(d['ms_per_step'],3), 'hfgen', round(k['hf_generator_hq_kernel']['launch_ms'],4))" || tail -3 gpurun_out/ab_x.err; }
run base
for v in 8 16; do XAAC_B200_LIB=$PWD/build/var/hf_$v.so run ahead_$v; done
XAAC_B200_LIB=$PWD/build/var/hf_8.so timeout 600 python -m pytest tests/test_hfgen_gpu.py tests/test_sbrdec_gpu.py -x -q -m gpu 2>&1 | tail -2
