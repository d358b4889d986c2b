timeout 900 python -m pytest tests/test_peaklim_gpu.py -x -q -m gpu 2>&1 | tail -2
timeout 300 python bench.py --workload aac_lc_stereo_output --steps 10 --warmup 3 --no-cpu-baseline --no-extra-stages > gpurun_out/lcout_split_e.json 2> gpurun_out/lcout_split_e.err
python - <<'P'
import json
d=json.loads(open("gpurun_out/lcout_split_e.json").read().strip().splitlines()[-1])
print(d["value"], d["ms_per_step"], d.get("e2e",{}).get("value"))
for k,v in d.get("kernels",{}).items(): print(k, round(v["launch_ms"],4), v.get("frac"))
P
