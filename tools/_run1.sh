timeout 1200 python -m pytest tests/test_sbrdec_gpu.py tests/test_chain_gpu.py -x -q -m gpu 2>&1 | tail -3
for uf in 0; do
XAAC_B200_SBR_UNFUSED=$uf timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-extra-stages --e2e-steps 3 2>gpurun_out/ab_$uf.err | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); k=d['kernels']; print('unfused=$uf chain', d['ms_per_step'], d['value'], 'e2e', d['e2e']['value']); print({n:round(v['launch_ms'],4) for n,v in k.items()})" || tail -5 gpurun_out/ab_$uf.err
done
