timeout 900 python -m pytest tests/test_peaklim_gpu.py -x -q -m gpu 2>&1 | tail -2
timeout 300 python bench.py --workload aac_lc_stereo_output --steps 10 --warmup 3 --no-cpu-baseline --no-extra-stages > gpurun_out/lcout_split_d.json 2> gpurun_out/lcout_split_d.err
python - <<'P'
import json
d=json.loads(open("gpurun_out/lcout_split_d.json").read().strip().splitlines()[-1])
print(d["value"], d["ms_per_step"], d.get("e2e",{}).get("value"))
for k,v in d.get("kernels",{}).items(): print(k, round(v["launch_ms"],4), v.get("frac"))
P

python - <<'P'
import csv
rows=[r for r in csv.reader(open("gpurun_out/peaklim_quick_d.csv")) if len(r)>10 and r[0].isdigit()]
by={}
for r in rows: by.setdefault((r[0],r[4].split("(")[0]),{})[r[-3]]=r[-1]
for k,v in by.items():
    print(k, {m.replace("smsp__average_warps_issue_stalled_","st_").replace("_per_issue_active.ratio",""):x for m,x in v.items() if not m.startswith("launch")})
P
