run() { timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-extra-stages --e2e-steps 1 2>gpurun_out/ab_x.err | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); k=d['kernels']; print('$1 chain', round(d['ms_per_step'],3), 'synth', round(k['qmf_synth_hq_kernel']['launch_ms'],4), 'ps', round(k['ps_frame_kernel']['launch_ms'],4))" || tail -3 gpurun_out/ab_x.err; }
run base
for v in p e pe; do XAAC_B200_LIB=$PWD/build/var/var_$v.so run var_$v; done
run base2
