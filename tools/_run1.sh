run() { timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-extra-stages --e2e-steps 1 2>gpurun_out/ab_x.err | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$1 chain', round(d['ms_per_step'],3), round(d['value']/1e6,3))" || tail -3 gpurun_out/ab_x.err; }
run base
for c in 8192 16384 32768 65536; do for k in 2 3 4; do XAAC_B200_DEV_CHUNK=$c XAAC_B200_DEV_STREAMS=$k run "chunk=$c streams=$k"; done; done
