set -x
timeout 900 python bench.py > gpurun_out/r2_bench_default_final5.json 2> gpurun_out/r2_bench_default_final5.err; tail -2 gpurun_out/r2_bench_default_final5.err
timeout 600 python bench.py --impl reference > gpurun_out/r2_bench_default_ref_final5.json 2> gpurun_out/r2_bench_default_ref_final5.err
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_chain_launches_f.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extra-stages --e2e-steps 1 > /dev/null 2>&1
grep -c xb:: gpurun_out/r2_chain_launches_f.csv
python - <<'P'
import json
d=json.loads([l for l in open("gpurun_out/r2_bench_default_final5.json").read().strip().splitlines() if l.startswith("{")][-1])
print("value %.4g ms %.4g e2e %.4g launches %d"%(d["value"], d["ms_per_step"], d["e2e"]["value"], d["gpu_launches"]), d["clocks"])
print({n:(round(v['launch_ms'],4), round(v['frac'],3) if v.get('frac') else None) for n,v in d['kernels'].items()})
print(d["roofline"]["frac"], d["cpu_baseline"])
print({k:(v["value"], v["ms_per_step"]) for k,v in d["other_configs"].items()})
r=json.loads([l for l in open("gpurun_out/r2_bench_default_ref_final5.json").read().strip().splitlines() if l.startswith("{")][-1])
print("ref", r["value"], r["cpu_baseline"])
P
