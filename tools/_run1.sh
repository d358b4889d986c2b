timeout 600 python bench.py --workload sbr_sideinfo --steps 20 --warmup 5 --no-extra-stages > gpurun_out/r2_bench_sideinfo_a.json 2> gpurun_out/r2_bench_sideinfo_a.err
tail -2 gpurun_out/r2_bench_sideinfo_a.err
python - <<'P'
import json
d=json.loads(open("gpurun_out/r2_bench_sideinfo_a.json").read().strip().splitlines()[-1])
print("value",d["value"],"ms",d["ms_per_step"],"e2e",d["e2e"]["value"],"roofline",d["roofline"]["frac"],d["roofline"]["achieved"],"cpu",d["cpu_baseline"]["value"],d["cpu_baseline"]["cores"])
P
