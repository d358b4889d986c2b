for c in 2048 4096 8192 16384 32768; do
  XAAC_B200_HOST_CHUNK=$c timeout 200 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-extra-stages --e2e-steps 10 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('chunk $c', 'value', round(d['value']/1e6,2), 'e2e', round(d['e2e']['value']/1e6,3))"
done
