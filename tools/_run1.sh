run() { timeout 200 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-extra-stages --e2e-steps 10 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$1', 'value', round(d['value']/1e6,2), 'e2e', round(d['e2e']['value']/1e6,3))"; }
for c in 2048 4096 8192 16384; do XAAC_B200_HOST_CHUNK=$c run "pipe3 chunk $c"; done
for k in 2 4 6; do for c in 2048 4096 8192; do XAAC_B200_LIB=$PWD/build/var/pipe_$k.so XAAC_B200_HOST_CHUNK=$c run "pipe$k chunk $c"; done; done
