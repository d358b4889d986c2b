timeout 900 python bench.py > gpurun_out/r2_bench_default_final3.json 2> gpurun_out/r2_bench_default_final3.err; tail -2 gpurun_out/r2_bench_default_final3.err
timeout 600 python bench.py --impl reference > gpurun_out/r2_bench_default_ref_final3.json 2>/dev/null
python - <<'P'
import json
for f in ("r2_bench_default_final3","r2_bench_default_ref_final3"):
    d=json.loads(open("gpurun_out/%s.json"%f).read().strip().splitlines()[-1])
    print(f, "value %.4g ms %.4g e2e %.4g"%(d["value"], d["ms_per_step"], d["e2e"]["value"]), "roof", (d.get("roofline") or {}).get("frac"), "dom", ((d.get("roofline") or {}).get("dominant_kernel") or {}).get("kernel"))
    if "other_configs" in d:
        for k,v in d["other_configs"].items(): print("   ", k, "%.4g"%v["value"], "e2e %.4g"%v["e2e"]["value"])
    if "stage_rooflines" in d:
        for k,v in d["stage_rooflines"].items(): print("   stage", k, v["kernel"], round(v["frac"],3))
P
