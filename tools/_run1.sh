timeout 2400 python -m pytest tests -x -q -m gpu 2>&1 | tail -4
timeout 900 python bench.py > gpurun_out/r2_bench_default_final2.json 2> gpurun_out/r2_bench_default_final2.err; tail -2 gpurun_out/r2_bench_default_final2.err
timeout 600 python bench.py --impl reference > gpurun_out/r2_bench_default_ref_final2.json 2>/dev/null
timeout 600 python bench.py --workload aac_lc_stereo_output --no-extra-stages > gpurun_out/r2_bench_lcout_final2.json 2>/dev/null
timeout 600 python bench.py --workload sbr_sideinfo --no-extra-stages > gpurun_out/r2_bench_sideinfo_final2.json 2>/dev/null
python - <<'P'
import json
for f in ("r2_bench_default_final2","r2_bench_default_ref_final2","r2_bench_lcout_final2","r2_bench_sideinfo_final2"):
    try:
        d=json.loads(open("gpurun_out/%s.json"%f).read().strip().splitlines()[-1])
        print(f, "value %.4g ms %.4g e2e %.4g"%(d["value"], d["ms_per_step"], d["e2e"]["value"]), "roof", (d.get("roofline") or {}).get("frac"), "dom", ((d.get("roofline") or {}).get("dominant_kernel") or {}).get("kernel"), "cpu", (d.get("cpu_baseline") or {}).get("value"))
    except Exception as e: print(f,"ERR",e)
P
