for v in z0; do
  export XAAC_B200_LIB=$PWD/build/var/libxaac_b200_$v.so
  timeout 300 python bench.py --workload qmf_synth_hq --steps 20 --warmup 5 --no-cpu-baseline --no-extra-stages 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('variant [$v]', d['ms_per_step'], d['roofline']['frac'])"
done
