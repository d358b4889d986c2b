#!/usr/bin/env python
"""Rebuild profiles/r2_traffic.json (per-unit DRAM bytes of every kernel) from ncu metric passes made with tools/ncu_quick.sh.
usage: tools/update_traffic.py <csv>:<units>:<description> [...]
Every entry records the sha256 of the kernel's source file; bench.py flags an entry as stale when the source has changed since."""
import collections
import csv
import re
import hashlib
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = {"imdct_ola_kernel": "imdct_kernels.cu", "pcm16_from_imdct_kernel": "sbr_glue_kernels.cu", "sbr_pre_kernel": "sbr_glue_kernels.cu",
       "sbr_scale_kernel": "sbr_glue_kernels.cu", "sbr_post_kernel": "sbr_glue_kernels.cu", "qmf_anal_hq_kernel": "qmf_anal_kernel.cu",
       "hf_generator_hq_kernel": "hfgen_kernel.cu", "calc_sbrenvelope_hq_kernel": "envcalc_kernel.cu", "ps_frame_kernel": "ps_kernel.cu",
       "qmf_synth_hq_kernel": "qmf_synth_kernel.cu", "usac_fd_kernel": "usac_fd_kernel.cu", "esbr_anal_kernel": "esbr_synth_kernel.cu",
       "esbr_hbe_kernel": "esbr_hbe_kernel.cu", "esbr_hfgen_kernel": "esbr_hfgen_kernel.cu", "esbr_envcalc_kernel": "esbr_envcalc_kernel.cu",
       "esbr_synth_kernel": "esbr_synth_kernel.cu", "esbr_ps_kernel": "esbr_ps_kernel.cu", "sbr_dec_lp_kernel": "sbr_lp_kernel.cu", "peak_limiter_kernel": "peaklim_kernel.cu",
       "peak_limiter_smooth_kernel": "peaklim_kernel.cu", "peak_limiter_finish_kernel": "peaklim_kernel.cu",
       "sbr_sideinfo_kernel": "sbr_sideinfo_kernel.cu", "ps_sideinfo_kernel": "sbr_sideinfo_kernel.cu",
       "sbr_front_hq_kernel": "qmf_anal_kernel.cu", "calc_sbrenvelope_hq_post_kernel": "envcalc_kernel.cu"}


def sha(path):
    return hashlib.sha256(open(path, "rb").read()).hexdigest()[:16]


def parse(path, units):
    lines = [ln for ln in open(path) if ln.startswith('"')]
    per = collections.OrderedDict()
    for row in csv.DictReader(lines):
        k = re.sub(r"<.*>$", "", row["Kernel Name"].split("(")[0].replace("void ", "").replace("xb::", "").strip())
        per.setdefault((row["ID"], k), {})[row["Metric Name"]] = (row["Metric Value"], row["Metric Unit"])
    acc = collections.defaultdict(list)
    for (_, k), m in per.items():
        def val(n):
            v, u = m[n]
            return float(v.replace(",", "")) * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1)
        acc[k].append((val("dram__bytes_read.sum") + val("dram__bytes_write.sum")) / units)
    # launches of the chunked host-buffer arm (a few thousand units each) are in the same list: keep the full-batch ones
    full = {k: [x for x in v if x >= 0.5 * max(v)] for k, v in acc.items()}
    return {k: sum(v) / len(v) for k, v in full.items()}


def main():
    out_path = os.path.join(ROOT, "profiles", "r2_traffic.json")
    out = json.load(open(out_path)) if os.path.exists(out_path) else {}
    out["_comment"] = ("DRAM bytes per unit (dram__bytes_read.sum + dram__bytes_write.sum) / units from ncu metric passes under gpurun "
                       "(tools/ncu_quick.sh, tools/update_traffic.py); bench.py multiplies by the units of a launch for roofline.traffic "
                       "and flags an entry whose kernel source has changed since (source_sha256)")
    for a in sys.argv[1:]:
        path, units, desc = a.split(":", 2)
        for k, t in parse(path, int(units)).items():
            if k not in SRC:
                continue
            src = os.path.join("libxaac_b200", "csrc", SRC[k])
            out[k] = {"bytes_per_unit": round(t), "source": f"{os.path.relpath(path, ROOT) if os.path.isabs(path) else path}, {desc}",
                      "kernel_source": src, "source_sha256": sha(os.path.join(ROOT, src))}
    json.dump(out, open(out_path, "w"), indent=1)
    print("wrote", out_path, sorted(k for k in out if not k.startswith("_")))


if __name__ == "__main__":
    main()
