for v in "" ps2; do
  if [ -n "$v" ]; then export XAAC_B200_LIB=$PWD/build/var/libxaac_b200_$v.so; fi
  timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-extra-stages --e2e-steps 1 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('variant [$v]', d['ms_per_step'], d['value'], d['kernels']['ps_frame_kernel']['launch_ms'])"
done
export XAAC_B200_LIB=$PWD/build/var/libxaac_b200_ps2.so
timeout 600 python -m pytest tests/test_sbrdec_gpu.py tests/test_chain_gpu.py -x -q -m gpu 2>&1 | tail -2
