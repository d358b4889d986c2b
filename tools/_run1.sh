for v in "" ls5x2 ls10 w8 w5x2; do
  if [ -n "$v" ]; then export XAAC_B200_LIB=$PWD/build/var/libxaac_b200_$v.so; fi
  timeout 300 python bench.py --workload qmf_synth_hq --steps 10 --warmup 3 --no-cpu-baseline --no-extra-stages 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('variant [$v]', d['ms_per_step'], d['roofline']['frac'])"
done
export XAAC_B200_LIB=$PWD/build/var/libxaac_b200_ls5x2.so
timeout 300 python -m pytest tests/test_qmf_synth_gpu.py -x -q -m gpu 2>&1 | tail -3
bash tools/ncu_quick.sh qmf_synth_hq_kernel gpurun_out/synth_v2_quick_c.csv -- python bench.py --workload qmf_synth_hq --steps 3 --warmup 3 --no-cpu-baseline --no-extra-stages > /dev/null 2>&1
grep '^"0"' gpurun_out/synth_v2_quick_c.csv | awk -F'","' '{print $(NF-2), $(NF)}' | grep -v "launch__\|barrier_per\|lg_thr\|mio\|branch"
