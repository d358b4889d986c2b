timeout 2400 python -m pytest tests -x -q -m gpu 2>&1 | tail -3
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 600 python bench.py --workload aac_lc_stereo_output --no-extra-stages > gpurun_out/r2_bench_lcout_final3.json 2>/dev/null
python -c "
import json
d=json.loads(open('gpurun_out/r2_bench_lcout_final3.json').read().strip().splitlines()[-1]); print('lcout', d['value'], d['ms_per_step'], d['e2e']['value'], d['cpu_baseline']['value'])"
