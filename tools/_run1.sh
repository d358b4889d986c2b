for v in "" u20 u12; do
  if [ -n "$v" ]; then export XAAC_B200_LIB=$PWD/build/var/libxaac_b200_$v.so; fi
  timeout 300 python bench.py --workload usac_fd_imdct --steps 10 --warmup 3 --no-cpu-baseline --no-extra-stages --e2e-steps 1 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('variant [$v]', d['ms_per_step'], d['roofline']['frac'])"
done
