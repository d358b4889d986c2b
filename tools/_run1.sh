M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum
ncu --metrics $M --clock-control none -s 33 -c 22 --csv --log-file gpurun_out/r2_chain_launches_e.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extra-stages --e2e-steps 1 > /dev/null 2>&1
timeout 900 python bench.py > gpurun_out/r2_bench_default_final4.json 2> gpurun_out/r2_bench_default_final4.err; tail -2 gpurun_out/r2_bench_default_final4.err
timeout 600 python bench.py --impl reference > gpurun_out/r2_bench_default_ref_final4.json 2>/dev/null
python - <<'P'
import json
for f in ("r2_bench_default_final4","r2_bench_default_ref_final4"):
    d=json.loads(open("gpurun_out/%s.json"%f).read().strip().splitlines()[-1])
    print(f, "value %.4g ms %.4g e2e %.4g"%(d["value"], d["ms_per_step"], d["e2e"]["value"]), "roof", (d.get("roofline") or {}).get("frac"), "dom", ((d.get("roofline") or {}).get("dominant_kernel") or {}))
    for k,v in (d.get("kernels") or {}).items(): print("   ", k, round(v["launch_ms"],3), round(v["share_of_step"],3), None if v.get("frac") is None else round(v["frac"],3))
P
