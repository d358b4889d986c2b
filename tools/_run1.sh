timeout 900 python -m pytest tests/test_qmf_synth_gpu.py tests/test_chain_gpu.py tests/test_sbrdec_gpu.py -x -q -m gpu 2>&1 | tail -4
for v in "" RING; do
  if [ -n "$v" ]; then export XAAC_B200_SYNTH_RING=1; fi
  timeout 300 python bench.py --workload qmf_synth_hq --steps 20 --warmup 5 --no-cpu-baseline --no-extra-stages 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('variant [$v]', d['ms_per_step'], d['roofline']['frac'])"
done
unset XAAC_B200_SYNTH_RING
bash tools/ncu_quick.sh qmf_synth_hq gpurun_out/r2_synth_g4_quick_a.csv -- python bench.py --workload qmf_synth_hq --steps 3 --warmup 3 --no-cpu-baseline --no-extra-stages > /dev/null 2>&1
grep '^"0"' gpurun_out/r2_synth_g4_quick_a.csv | awk -F'","' '{print $(NF-2), $(NF)}' | grep -v "launch__\|barrier_per\|lg_thr\|mio\|branch"
