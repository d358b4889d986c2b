#!/usr/bin/env python
"""bench.py — throughput of the libxaac decode-side DSP hot path on B200.

  python bench.py [--gpus N] [--steps K] [--warmup W]            our CUDA path
  python bench.py --impl reference [--gpus N] [--steps K] ...    the reference's own generic-C path on host cores

A "step" is one pass of the hot path over one batch of synthetic pre-parsed frames (one frame for every stream
of the batch, state carried from step to step like consecutive frames of a stream).  One JSON line is printed by
rank 0.  `value` is device-resident throughput, `e2e` goes through the host-buffer C-ABI call with pinned host
buffers (H2D + kernel + D2H inside the timed region).
"""
import argparse
import ctypes
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

IMDCT_BYTES_PER_UNIT = 12288  # 4096 spec in + 2048 overlap in + 2048 overlap out + 4096 WORD32 out
SYNTH_BYTES_PER_UNIT = 25600  # 16384 matrix + 2560 state in + 2560 state out + 4096 PCM16 out
# algorithmic HBM bytes per unit and launch of every kernel of the HE-AACv2 chain
CHAIN_KERNEL_BYTES = {
    "imdct_ola_kernel": 12288,            # 4096 spec + 2048 + 2048 overlap + 4096 WORD32 out
    "pcm16_from_imdct_kernel": 6144,      # 4096 in + 2048 out
    "sbr_pre_kernel": 6144,               # 3072 overlap slots in + 3072 matrix rows out
    "qmf_anal_hq_kernel": 11520,          # 2048 PCM16 + 2 x 640 state + 8192 matrix
    "sbr_scale_kernel": 29184,            # 38 x 32 bands x 8 B read + written, 32 x 64 x 4 B cleared, 2 KB LPC rows r/w
    "hf_generator_hq_kernel": 13824,      # hfgen_kernel.cu header
    "calc_sbrenvelope_hq_kernel": 19760,  # envcalc_kernel.cu header
    "sbr_post_kernel": 4672,              # 1536 overlap out + LPC rows 2 x 1024 + parameters
    # the stage as the driver runs it (glue inside the heavy kernels; the three above only with XAAC_B200_SBR_UNFUSED=1):
    # 4096 WORD32 core output + 2 x 640 ring + 3072 overlap slots in + 38 rows x 512 B matrix out (32 bands x 8 B + the cleared
    # upper half) + 2 x 1024 LPC rows r/w + 160 HF generator record
    "sbr_front_hq_kernel": 30112,
    "calc_sbrenvelope_hq_post_kernel": 24432,  # envelope adjuster 19760 + previous-frame / overlap save 4672
    "ps_frame_kernel": 60928,             # ps_kernel.cu header
    "qmf_synth_hq_kernel": SYNTH_BYTES_PER_UNIT,
    # fused low-power stage (sbr_lp_kernel.cu): 2048 PCM16 in + 4096 PCM16 out + 2 x 5536 channel state (analysis ring 644,
    # synthesis ring 2564, envelope state 464, overlap rows 1536, LPC rows 256, scale factors / misc / bw 72) + 1480 side info
    "sbr_dec_lp_kernel": 18696,
    # per channel unit (a stream = 2 units): 4096 WORD32 in + 2048 PCM16 out + 2 x 1.35 KB limiter state (220-sample window)
    "peak_limiter_kernel": 8850,
    # streams whose attack / release recursion is active finish in two follow-up kernels; their traffic is scratch (raw and
    # smoothed gains, 4 KB per stream each way) on top of the stage's algorithmic bytes, so no roofline figure is attached
    "peak_limiter_smooth_kernel": None,
    "peak_limiter_finish_kernel": None,
}
ESBR_ANAL_BYTES_PER_UNIT = 14848   # 4096 float in + 1280 + 1280 WORD32 ring + 8192 (32 x 32 complex float out)
ESBR_SYNTH_BYTES_PER_UNIT = 34816  # 16384 float matrix + 5120 + 5120 WORD32 state + 8192 float out
USAC_FD_BYTES_PER_UNIT = 16384  # 4096 coefficients + 4096 overlap in + 4096 overlap out + 4096 WORD32 out
WORKLOADS = {
    # name -> (BASELINE.json config index, stereo frames per GPU, description)
    "heaacv2_chain": (3, 131072, "HE-AACv2 (SBR+PS) stereo 44.1 kHz batch=131072: full IMDCT->QMF->SBR->PS "
                                 "hybrid/decorrelate chain (fixed-point path of the reference, -esbr:0)"),
    "aac_lc_stereo_imdct_ola": (1, 65536, "AAC-LC stereo 44.1 kHz batch=65536 frames, IMDCT+OLA only"),
    "heaacv1_stereo_chain": (2, 65536, "HE-AACv1 stereo 48 kHz batch=65536: IMDCT + 64-band QMF analysis/synthesis + LPP "
                                       "HF-gen + env_calc (fixed-point path of the reference, -esbr:0: low-power SBR)"),
    "aac_lc_stereo_output": (1, 65536, "AAC-LC stereo 44.1 kHz batch=65536 frames with the reference's default flags: IMDCT+OLA "
                                       "-> peak limiter -> PCM16"),
    "usac_fd_imdct": (4, 131072, "xHE-AAC/USAC stereo 32 kHz batch=131072: the fixed-point FD core transform of the chain "
                                 "(ixheaacd_fd_frm_dec: IMDCT 1024/128 + windowing + overlap); the float eSBR stage is not "
                                 "built yet"),
    "esbr_anal32": (4, 65536, "xHE-AAC/USAC eSBR stereo 32 kHz: the 32-band eSBR QMF analysis bank of the chain "
                              "(ixheaacd_esbr_analysis_filt_block), batch=65536 stereo frames (131072 core channels)"),
    "esbr_generate_hf": (4, 65536, "xHE-AAC/USAC eSBR stereo: the float HF generator of the chain (ixheaacd_generate_hf: 38-slot "
                                   "covariance, 2nd-order complex prediction, patching, HBE high band), batch=65536 stereo "
                                   "frames (131072 core channels)"),
    "xheaac_stereo_chain": (4, 131072, "xHE-AAC/USAC eSBR stereo 32 kHz batch=131072 stereo frames (262144 core channels): FD core IMDCT -> "
                                       "float eSBR stage with the QMF harmonic transposer (HBE: real synthesis bank, complex analysis "
                                       "bank, stretch-2 products + pitch cross products as the reference encoder signals them), HF "
                                       "generator, envelope adjuster, polyphase QMF synthesis -> stereo PCM16"),
    "heaacv2_esbr_chain": (3, 65536, "HE-AACv2 (mono core + SBR + PS) decoded with the reference's DEFAULT flags, batch=65536 stereo "
                                     "frames: the float eSBR stage of the mono + PS element (harmonic transposer forced on for "
                                     "legacy streams) with the float parametric stereo -> float stereo output; the AAC-LC core "
                                     "IMDCT of the chain is not part of this workload (aac_lc_stereo_imdct_ola measures it)"),
    "sbr_sideinfo": (3, 131072, "HE-AAC (SBR) batch=131072 elements: SBR side-info dequantisation of the chain (ixheaacd_dec_sbrdata, "
                                "fixed-point path): envelope / noise-floor delta decoding, concealment, dequantisation, coupling"),
    "aac_lc_spectral": (1, 65536, "AAC-LC stereo 44.1 kHz batch=65536 frames: the pre-IMDCT spectral stage of the chain "
                                  "(ixheaacd_channel_pair_process: M/S and intensity stereo, perceptual noise substitution, TNS) on "
                                  "channel pairs, in place"),
    "esbr_hbe": (4, 65536, "xHE-AAC/USAC eSBR stereo: the QMF harmonic transposer of the chain (ixheaacd_qmf_hbe_apply), "
                           "batch=65536 stereo frames (131072 core channels)"),
    "xheaac_plain_stereo_chain": (4, 65536, "xHE-AAC/USAC stereo 32 kHz with eSBR, default (LPP) patching instead of the harmonic "
                                            "transposer, batch=65536: FD core IMDCT -> float eSBR stage (QMF analysis, HF generator, "
                                            "envelope adjuster, QMF synthesis) -> stereo PCM16, per channel"),
    "esbr_env_calc": (4, 65536, "xHE-AAC/USAC eSBR stereo: the float envelope adjuster of the chain (ixheaacd_sbr_env_calc, ORIG_SBR: "
                                "energies, gains in double, limiter, smoothing, noise, sinusoids), batch=65536 stereo frames "
                                "(131072 core channels)"),
    "esbr_synth64": (4, 65536, "xHE-AAC/USAC eSBR stereo 32 kHz: the 64-band eSBR QMF synthesis bank of the chain (per-slot core of "
                               "ixheaacd_esbr_synthesis_filt_block), batch=65536 stereo frames (131072 output channels)"),
    "qmf_synth_hq": (3, 65536, "stand-alone fixed-point HQ 64-band QMF synthesis stage of the HE-AAC chain, "
                               "batch=65536 stereo frames (131072 output channels)"),
}


def ncu_traffic():
    """per-unit DRAM bytes of each kernel measured by ncu (profiles/r2_traffic.json, tools/update_traffic.py); bench.py does not run
    ncu.  An entry whose kernel source has changed since the capture (source_sha256) is flagged stale instead of silently reused."""
    if getattr(ncu_traffic, "_cache", None) is not None:
        return ncu_traffic._cache
    p = os.path.join(ROOT, "profiles", "r2_traffic.json")
    out = {}
    if os.path.exists(p):
        import hashlib
        for k, v in json.load(open(p)).items():
            if not isinstance(v, dict):
                continue
            v = dict(v)
            src, h = v.get("kernel_source"), v.get("source_sha256")
            if src and h and os.path.exists(os.path.join(ROOT, src)):
                if hashlib.sha256(open(os.path.join(ROOT, src), "rb").read()).hexdigest()[:16] != h:
                    v["stale"] = True
                    sys.stderr.write(f"[bench] warning: {src} changed since the ncu capture of {k}: its DRAM traffic figure is stale\n")
            out[k] = v
    ncu_traffic._cache = out
    return out


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


# ---------------------------------------------------------------------------------------------------------
# synthetic pre-parsed batch: per-unit magnitude 2^12..2^27, corner units, window-sequence walk
# ---------------------------------------------------------------------------------------------------------
def sequence_walk(n_units, n_steps, seed):
    """ics[step][unit] = (window_sequence, window_shape): legal AAC block-switching walk with ~90 % long->long,
    4 % start, 4 % stop, 2 % short in the stationary mix; both channels of a frame share the sequence."""
    rng = np.random.default_rng(seed)
    n_frames = n_units // 2
    prev = np.zeros(n_frames, np.uint8)
    out = np.zeros((n_steps, n_units, 2), np.uint8)
    for s in range(n_steps):
        r = rng.random(n_frames)
        nxt = np.zeros(n_frames, np.uint8)
        longish = (prev == 0) | (prev == 3)
        nxt[longish & (r < 0.045)] = 1                 # long -> start
        shortish = ~longish
        nxt[shortish] = np.where(r[shortish] < 0.33, 2, 3)  # start/short -> short | stop
        shape = (rng.random(n_frames) < 0.9).astype(np.uint8)  # the reference encoder signals KBD for most frames
        out[s, 0::2, 0] = nxt
        out[s, 1::2, 0] = nxt
        out[s, 0::2, 1] = shape
        out[s, 1::2, 1] = shape
        prev = nxt
    return out


def make_spec_torch(n_units, seed, device):
    import torch
    g = torch.Generator(device=device)
    g.manual_seed(seed)
    s = torch.randint(12, 28, (n_units, 1), generator=g, device=device, dtype=torch.int32)
    spec = torch.empty((n_units, 1024), dtype=torch.int32, device=device)
    chunk = 16384
    for i in range(0, n_units, chunk):
        r = torch.randint(-(2 ** 31), 2 ** 31 - 1, (min(chunk, n_units - i), 1024), generator=g, device=device,
                          dtype=torch.int64).to(torch.int32)
        spec[i:i + chunk] = r >> (31 - s[i:i + chunk])
    # corner units: silence, alternating full scale, impulse, DC
    spec[0] = 0
    spec[1] = torch.where(torch.arange(1024, device=device) % 2 == 0, 2 ** 31 - 1, -(2 ** 31)).to(torch.int32)
    spec[2] = 0
    spec[2, 17] = 2 ** 31 - 1
    spec[3] = 1 << 20
    return spec


def make_synth_inputs_torch(n_units, seed, device):
    """QMF matrices [n,32,128] with per-unit magnitude 2^14..2^26 and typical HE-AAC scale factors."""
    import torch
    g = torch.Generator(device=device)
    g.manual_seed(seed)
    s = torch.randint(14, 27, (n_units, 1, 1), generator=g, device=device, dtype=torch.int32)
    matrix = torch.empty((n_units, 32, 128), dtype=torch.int32, device=device)
    chunk = 8192
    for i in range(0, n_units, chunk):
        r = torch.randint(-(2 ** 31), 2 ** 31 - 1, (min(chunk, n_units - i), 32, 128), generator=g, device=device,
                          dtype=torch.int64).to(torch.int32)
        matrix[i:i + chunk] = r >> (31 - s[i:i + chunk])
    params = torch.zeros((n_units, 8), dtype=torch.int16, device=device)
    params[:, 0] = torch.randint(-10, -3, (n_units,), generator=g, device=device)  # ov_lb_scale
    params[:, 1] = torch.randint(-10, -3, (n_units,), generator=g, device=device)  # lb_scale
    params[:, 2] = torch.randint(-12, -4, (n_units,), generator=g, device=device)  # hb_scale
    params[:, 3] = -6
    lsb = torch.randint(16, 33, (n_units,), generator=g, device=device)
    params[:, 4] = lsb
    params[:, 5] = torch.clamp(lsb + torch.randint(8, 33, (n_units,), generator=g, device=device), max=64)
    params[:, 6] = 6
    matrix[0] = 0
    matrix[1] = 2 ** 31 - 1
    return matrix, params


def cpu_arm_synth(n_units, threads, seed, reps=1, min_seconds=0.0):
    """Time ixheaacd_cplx_synt_qmffilt (HQ) per unit on host threads. Returns (units_per_s, kind)."""
    from tests import oracle_util
    ref = oracle_util.Ref.try_load()
    matrix0, fs, pos, params = oracle_util.synth_qmf_units(n_units, seed)
    params[:, 0:3] = np.clip(params[:, 0:3], -12, -3)
    P = oracle_util.P
    pcm = np.zeros((n_units, 2048), np.int16)
    bounds = np.linspace(0, n_units, threads + 1).astype(int)
    sf = np.ascontiguousarray(params[:, 0:4], np.int32)
    off = np.ascontiguousarray(pos[:, 0], np.int32)
    fp = np.ascontiguousarray(pos[:, 1], np.int32)
    lsb = np.ascontiguousarray(params[:, 4], np.int32)
    usb = np.ascontiguousarray(params[:, 5], np.int32)
    if ref is not None:
        kind, fn, pre = "reference", ref.lib.ref_synt_qmffilt_hq_batch, []
    else:
        orc = oracle_util.Oracle()
        kind, fn, pre = "port", orc.lib.xo_synt_qmffilt_hq_batch, [P(orc.qrom)]

    def work(t, matrix):
        a, b = bounds[t], bounds[t + 1]
        if b > a:
            fn(*pre, P(matrix[a:b]), P(fs[a:b]), P(off[a:b]), P(fp[a:b]), P(sf[a:b]), P(lsb[a:b]), P(usb[a:b]),
               P(pcm[a:b]), int(b - a))

    def one_pass():
        matrix = matrix0.copy()  # the reference modifies the matrix in place
        t0 = time.perf_counter()
        th = [threading.Thread(target=work, args=(t, matrix)) for t in range(threads)]
        for x in th:
            x.start()
        for x in th:
            x.join()
        return time.perf_counter() - t0

    one_pass()
    dt, done = 0.0, 0
    while done < reps or dt < min_seconds:
        dt += one_pass()
        done += 1
    return n_units * done / dt, kind


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms while the timed region runs."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._pump, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for ln in self.proc.stdout:
            self.lines.append(ln.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
            except ValueError:
                continue
            for nm, v in zip(names, f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons),
                "samples": len(sm)}


# ---------------------------------------------------------------------------------------------------------
# CPU arm: the reference's own generic-C stage (oracle/_ref) or, if that was not built, our C port (oracle/)
# ---------------------------------------------------------------------------------------------------------
def cpu_arm(n_units, threads, seed, reps=1, min_seconds=0.0):
    """Time ixheaacd_imdct_process on `n_units` units spread over `threads` host threads (one private state per
    unit; the 1024-sample path touches no global scratch). Returns (units_per_s, kind)."""
    from tests import oracle_util
    ref = oracle_util.Ref.try_load()
    rng = np.random.default_rng(seed)
    s = rng.integers(12, 28, size=(n_units, 1))
    spec0 = ((rng.random((n_units, 1024)) * 2 - 1) * (2.0 ** s)).astype(np.int64).astype(np.int32)
    walk = sequence_walk(n_units, reps + 1, seed)
    P = oracle_util.P
    if ref is not None:
        kind, lib = "reference", ref.lib
        fn = lib.ref_imdct_process_batch
    else:
        kind = "port"
        orc = oracle_util.Oracle()
        lib, rom = orc.lib, orc.rom
        fn = lib.xo_imdct_process_batch
    ovl = np.zeros((n_units, 512), np.int32)
    pshape = np.zeros(n_units, np.int32)
    pseq = np.zeros(n_units, np.int32)
    out = np.zeros((n_units, 1024), np.int32)
    adj = np.zeros(n_units, np.int32)
    bounds = np.linspace(0, n_units, threads + 1).astype(int)

    def work(t, step, spec):
        a, b = bounds[t], bounds[t + 1]
        if b <= a:
            return
        ws = np.ascontiguousarray(walk[step, a:b, 0], np.int32)
        wh = np.ascontiguousarray(walk[step, a:b, 1], np.int32)
        args = [P(spec[a:b]), P(ovl[a:b]), P(pshape[a:b]), P(pseq[a:b]), P(ws), P(wh), P(out[a:b]), P(adj[a:b]),
                int(b - a)]
        if kind == "port":
            args = [P(rom)] + args
        fn(*args)

    def one_pass(step):
        spec = spec0.copy()  # the reference destroys its input spectrum
        t0 = time.perf_counter()
        th = [threading.Thread(target=work, args=(t, step, spec)) for t in range(threads)]
        for x in th:
            x.start()
        for x in th:
            x.join()
        return time.perf_counter() - t0

    one_pass(0)  # warm-up (page faults, caches)
    dt, done = 0.0, 0
    while done < reps or dt < min_seconds:
        dt += one_pass(1 + done % reps)
        done += 1
    return n_units * done / dt, kind


def load_chain_golden():
    """HE-AACv2 side info / state tapped from a real decode of the reference (tests/golden/sbrdec_tapped.npz, made by
    tools/make_golden.py): records 1..11 are consecutive frames of one 44.1 kHz mono+PS stream."""
    g = np.load(os.path.join(ROOT, "tests", "golden", "sbrdec_tapped.npz"))
    assert g["side"][1:12, 737].all()
    return g["side"][1:12].copy(), g["st_in"][1].copy(), g["ps_in"][1].copy()


def chain_inputs_np(n_units, seed):
    rng = np.random.default_rng(seed)
    s = rng.integers(12, 22, size=(n_units, 1))
    spec = ((rng.random((n_units, 1024)) * 2 - 1) * (2.0 ** s)).astype(np.int64).astype(np.int32)
    return spec


def cpu_arm_chain(n_units, threads, seed, reps=1, min_seconds=0.0):
    """Time the reference's own chain per stream-frame on host threads: ixheaacd_imdct_process -> WORD32->WORD16
    hand-over -> ixheaacd_sbr_dec (HQ + PS, 2 x ixheaacd_cplx_synt_qmffilt), persistent reference structs per stream.
    Returns (stream_frames_per_s * 2, kind): the factor 2 keeps the caller's units/2 = frames convention."""
    from tests import oracle_util
    ref = oracle_util.Ref.try_load()
    if ref is None:
        raise SystemExit("bench.py: the HE-AACv2 chain CPU baseline needs oracle/_ref/libxaac_ref.so (make ref)")
    side_frames, st0, ps0 = load_chain_golden()
    P = oracle_util.P
    side0 = np.ascontiguousarray(np.tile(side_frames[0], (n_units, 1)))
    st = np.ascontiguousarray(np.tile(st0, (n_units, 1)))
    ps = np.ascontiguousarray(np.tile(ps0, (n_units, 1)))
    ref.lib.ref_chain_create.restype = ctypes.c_void_p
    h = ctypes.c_void_p(ref.lib.ref_chain_create(n_units, P(side0), P(st), P(ps)))
    spec0 = chain_inputs_np(n_units, seed)
    walk = sequence_walk(n_units + (n_units & 1), reps + 1, seed)[:, :n_units]
    out = np.zeros((n_units, 2048, 2), np.int16)
    bounds = np.linspace(0, n_units, threads + 1).astype(int)
    sides = [np.ascontiguousarray(side_frames[(np.arange(n_units) + f) % 11]) for f in range(11)]

    def work(t, spec, ics, side):
        a, b = int(bounds[t]), int(bounds[t + 1])
        if b > a:
            ref.lib.ref_chain_step(h, a, b, P(spec), P(ics), P(side), P(out))

    def one_pass(step):
        spec = spec0.copy()
        ics = np.ascontiguousarray(walk[step])
        t0 = time.perf_counter()
        th = [threading.Thread(target=work, args=(t, spec, ics, sides[step % 11])) for t in range(threads)]
        for x in th:
            x.start()
        for x in th:
            x.join()
        return time.perf_counter() - t0

    one_pass(0)
    dt, done = 0.0, 0
    while done < reps or dt < min_seconds:
        dt += one_pass(1 + done % reps)
        done += 1
    ref.lib.ref_chain_destroy(h)
    return 2.0 * n_units * done / dt, "reference"


def load_chain_lp_golden():
    """HE-AACv1 stereo side info / state tapped from a real decode of the reference (tests/golden/sbrdec_lp_tapped.npz):
    records 2..25 are 12 consecutive frames of the two channels (even records L, odd records R)."""
    g = np.load(os.path.join(ROOT, "tests", "golden", "sbrdec_lp_tapped.npz"))
    side = np.stack([g["side"][2:26:2], g["side"][3:26:2]], axis=1)  # [frame 12][ch 2][1232]
    st0 = np.stack([g["st_in"][2], g["st_in"][3]])                   # [ch 2][3920]
    return side.copy(), st0.copy()


def chain_lp_side(side_frames, n_units, f):
    """side info of frame f for n_units channel units (unit u = channel u & 1 of stream u >> 1, phase (u >> 1) mod 12)"""
    u = np.arange(n_units)
    return np.ascontiguousarray(side_frames[((u >> 1) + f) % 12, u & 1])


def cpu_arm_chain_lp(n_units, threads, seed, reps=1, min_seconds=0.0):
    """Time the reference's own stereo HE-AACv1 chain per channel unit on host threads: ixheaacd_imdct_process ->
    WORD32->WORD16 hand-over -> ixheaacd_sbr_dec (low_pow_flag = 1).  Returns (units_per_s, kind)."""
    from tests import oracle_util
    ref = oracle_util.Ref.try_load()
    if ref is None:
        raise SystemExit("bench.py: the HE-AACv1 chain CPU baseline needs oracle/_ref/libxaac_ref.so (make ref)")
    n_units -= n_units & 1
    side_frames, st0 = load_chain_lp_golden()
    P = oracle_util.P
    st = np.ascontiguousarray(np.tile(st0, (n_units // 2, 1)))
    ref.lib.ref_chain_lp_create.restype = ctypes.c_void_p
    h = ctypes.c_void_p(ref.lib.ref_chain_lp_create(n_units, P(chain_lp_side(side_frames, n_units, 0)), P(st)))
    spec0 = chain_inputs_np(n_units, seed)
    walk = sequence_walk(n_units, reps + 1, seed)
    out = np.zeros((n_units // 2, 2048, 2), np.int16)
    bounds = (np.linspace(0, n_units // 2, threads + 1).astype(int)) * 2
    sides = [chain_lp_side(side_frames, n_units, f) for f in range(12)]

    def work(t, spec, ics, side):
        a, b = int(bounds[t]), int(bounds[t + 1])
        if b > a:
            ref.lib.ref_chain_lp_step(h, a, b, P(spec), P(ics), P(side), P(out))

    def one_pass(step):
        spec = spec0.copy()
        ics = np.ascontiguousarray(walk[step])
        t0 = time.perf_counter()
        th = [threading.Thread(target=work, args=(t, spec, ics, sides[step % 12])) for t in range(threads)]
        for x in th:
            x.start()
        for x in th:
            x.join()
        return time.perf_counter() - t0

    one_pass(0)
    dt, done = 0.0, 0
    while done < reps or dt < min_seconds:
        dt += one_pass(1 + done % reps)
        done += 1
    ref.lib.ref_chain_destroy(h)
    return n_units * done / dt, "reference"


def cpu_arm_lc_output(n_units, threads, seed, reps=1, min_seconds=0.0):
    """Time ixheaacd_imdct_process x 2 channels + ixheaacd_peak_limiter_process + round16 per stereo frame on host threads.
    Returns (units_per_s, kind) with units = channels."""
    from tests import oracle_util
    ref = oracle_util.Ref.try_load()
    if ref is None:
        raise SystemExit("bench.py: the AAC-LC output CPU baseline needs oracle/_ref/libxaac_ref.so (make ref)")
    P = oracle_util.P
    n_units -= n_units & 1
    nf = n_units // 2
    rng = np.random.default_rng(seed)
    s = rng.integers(12, 30, size=(n_units, 1))
    spec0 = ((rng.random((n_units, 1024)) * 2 - 1) * (2.0 ** s)).astype(np.int64).astype(np.int32)
    walk = sequence_walk(n_units, reps + 1, seed)
    ovl = np.zeros((n_units, 512), np.int32)
    pshape = np.zeros(n_units, np.int32)
    pseq = np.zeros(n_units, np.int32)
    out = np.zeros((n_units, 1024), np.int32)
    adj = np.zeros(n_units, np.int32)
    st = np.tile(ref.peak_limiter_init(2, 44100)[0], (nf, 1))
    inter = np.zeros((nf, 1024, 2), np.int32)
    pcm = np.zeros((nf, 1024, 2), np.int16)
    bounds = np.linspace(0, nf, threads + 1).astype(int) * 2

    def work(t, step, spec):
        a, b = int(bounds[t]), int(bounds[t + 1])
        if b <= a:
            return
        ws = np.ascontiguousarray(walk[step, a:b, 0], np.int32)
        wh = np.ascontiguousarray(walk[step, a:b, 1], np.int32)
        ref.lib.ref_imdct_process_batch(P(spec[a:b]), P(ovl[a:b]), P(pshape[a:b]), P(pseq[a:b]), P(ws), P(wh), P(out[a:b]),
                                        P(adj[a:b]), b - a)
        fa, fb = a // 2, b // 2
        inter[fa:fb] = out[a:b].reshape(fb - fa, 2, 1024).transpose(0, 2, 1)
        q = np.ascontiguousarray(adj[a:b].astype(np.int8).reshape(fb - fa, 2))
        ref.lib.ref_peak_limiter_batch(P(st[fa:fb]), P(inter[fa:fb]), P(q), P(pcm[fa:fb]), 2, fb - fa)

    def one_pass(step):
        spec = spec0.copy()
        t0 = time.perf_counter()
        th = [threading.Thread(target=work, args=(t, step, spec)) for t in range(threads)]
        for x in th:
            x.start()
        for x in th:
            x.join()
        return time.perf_counter() - t0

    one_pass(0)
    dt, done = 0.0, 0
    while done < reps or dt < min_seconds:
        dt += one_pass(1 + done % reps)
        done += 1
    return n_units * done / dt, "reference"


def cpu_arm_esbr_anal(n_units, threads, seed, reps=1, min_seconds=0.0):
    """Time ixheaacd_esbr_analysis_filt_block per unit on host threads (ref_esbr_anal32, oracle/ref_shim.c)."""
    from tests import oracle_util
    ref = oracle_util.Ref.try_load()
    P = oracle_util.P
    x, st, pos = oracle_util.synth_esbr_anal_units(n_units, seed)
    st[:] = 0
    pos[:] = 0
    qmf = np.zeros((n_units, 32, 128), np.float32)
    bounds = np.linspace(0, n_units, threads + 1).astype(int)
    if ref is not None:
        kind, fn, pre = "reference", ref.lib.ref_esbr_anal32_batch, []
    else:
        orc = oracle_util.Oracle()
        kind, fn, pre = "port", orc.lib.xo_esbr_anal32_batch, [P(orc.esrom)]

    def work(t):
        a, b = int(bounds[t]), int(bounds[t + 1])
        if b > a:
            fn(*pre, P(x[a:b]), P(st[a:b]), P(pos[a:b]), P(qmf[a:b]), b - a)

    def one_pass():
        t0 = time.perf_counter()
        th = [threading.Thread(target=work, args=(t,)) for t in range(threads)]
        for z in th:
            z.start()
        for z in th:
            z.join()
        return time.perf_counter() - t0

    one_pass()
    dt, done = 0.0, 0
    while done < reps or dt < min_seconds:
        dt += one_pass()
        done += 1
    return n_units * done / dt, kind


def esbr_hfgen_bytes(par):
    """Algorithmic HBM bytes per unit of ixheaacd_generate_hf from its parameters: every source cell once (40 rows x the low
    band below f_master_tbl[0] for the covariance / patch branch — the patches re-read cells the covariance already
    touched — or 40 rows x the high band of the phase-vocoder buffer for the HBE branch), every output cell once (slots x
    [sub_band_start, 64)), 8 bytes per complex cell, + the 384-byte parameter record."""
    from tests.oracle_util import EHF
    num_mf = par[:, EHF["NUM_MF"]]
    lsb = par[:, EHF["FMASTER"]]
    usb = par[np.arange(len(par)), EHF["FMASTER"] + num_mf]
    sbs = par[:, EHF["SB_START"]]
    slots = 2 * (par[:, EHF["BORDER_LAST"]] - par[:, EHF["BORDER_FIRST"]])
    lpc = (par[:, EHF["PATCHING_MODE"]] != 0) | (par[:, EHF["HBE_FLAG"]] == 0)
    cells = np.where(lpc, 40 * (lsb - 1), 40 * (usb - sbs)) + slots * (64 - sbs)
    return 8.0 * cells + 384


def cpu_arm_esbr_hfgen(n_units, threads, seed, reps=1, min_seconds=0.0):
    """Time ixheaacd_generate_hf per unit on host threads (ref_esbr_generate_hf_batch, oracle/ref_shim.c)."""
    from tests import oracle_util
    ref = oracle_util.Ref.try_load()
    P = oracle_util.P
    d = oracle_util.synth_esbr_hfgen_units(n_units, seed)
    d["par"][15::16, oracle_util.EHF["INVF_TBL"]:oracle_util.EHF["INVF_TBL"] + 5] = 64
    dr, di, bw = d["dst_re"].copy(), d["dst_im"].copy(), d["bw_prev"].copy()
    patch = np.zeros((n_units, 8), np.int32)
    err = np.zeros(n_units, np.int32)
    bounds = np.linspace(0, n_units, threads + 1).astype(int)
    if ref is not None:
        kind, fn = "reference", ref.lib.ref_esbr_generate_hf_batch
    else:
        kind, fn = "port", oracle_util.Oracle().lib.xo_esbr_generate_hf_batch

    def work(t):
        a, b = int(bounds[t]), int(bounds[t + 1])
        if b > a:
            fn(P(d["src_re"][a:b]), P(d["src_im"][a:b]), P(d["pv_re"][a:b]), P(d["pv_im"][a:b]), P(dr[a:b]), P(di[a:b]),
               P(d["par"][a:b]), P(bw[a:b]), P(patch[a:b]), P(err[a:b]), b - a)

    def one_pass():
        t0 = time.perf_counter()
        th = [threading.Thread(target=work, args=(t,)) for t in range(threads)]
        for z in th:
            z.start()
        for z in th:
            z.join()
        return time.perf_counter() - t0

    one_pass()
    dt, done = 0.0, 0
    while done < reps or dt < min_seconds:
        dt += one_pass()
        done += 1
    return n_units * done / dt, kind


def esbr_envcalc_bytes(ipar):
    """Algorithmic HBM bytes per unit of ixheaacd_sbr_env_calc: the adjusted cells read and written once (8 bytes per complex
    cell each way), the smoothing history in and out (5120), the parameter records (1152 + 1856)."""
    from tests.oracle_util import EEC
    nsub = ipar[:, EEC["SB_END"]] - ipar[:, EEC["SB_START"]]
    last = ipar[np.arange(len(ipar)), EEC["BORDER"] + ipar[:, EEC["NUM_ENV"]]]
    slots = 2 * (last - ipar[:, EEC["BORDER"]])
    return 16.0 * slots * nsub + 5120 + 1152 + 1856


def cpu_arm_esbr_envcalc(n_units, threads, seed, reps=1, min_seconds=0.0):
    """Time ixheaacd_sbr_env_calc per unit on host threads (ref_esbr_env_calc_batch, oracle/ref_shim.c)."""
    from tests import oracle_util
    ref = oracle_util.Ref.try_load()
    P = oracle_util.P
    d = oracle_util.synth_esbr_envcalc_units(n_units, seed)
    E = oracle_util.EEC
    d["ipar"][:, E["NUM_NOISE_ENV"]] = np.where(d["ipar"][:, E["NUM_ENV"]] == 1, 1, 2)
    re, im, ipar, state = d["re"].copy(), d["im"].copy(), d["ipar"].copy(), d["state"].copy()
    err = np.zeros(n_units, np.int32)
    bounds = np.linspace(0, n_units, threads + 1).astype(int)
    if ref is not None:
        kind, fn, pre = "reference", ref.lib.ref_esbr_env_calc_batch, []
    else:
        kind, fn, pre = "port", oracle_util.Oracle().lib.xo_esbr_env_calc_batch, [P(oracle_util.esbr_random_phase())]

    def work(t):
        a, b = int(bounds[t]), int(bounds[t + 1])
        if b > a:
            fn(*pre, P(re[a:b]), P(im[a:b]), P(ipar[a:b]), P(d["fpar"][a:b]), P(state[a:b]), P(err[a:b]), b - a)

    def one_pass():
        t0 = time.perf_counter()
        th = [threading.Thread(target=work, args=(t,)) for t in range(threads)]
        for z in th:
            z.start()
        for z in th:
            z.join()
        return time.perf_counter() - t0

    one_pass()
    dt, done = 0.0, 0
    while done < reps or dt < min_seconds:
        dt += one_pass()
        done += 1
    return n_units * done / dt, kind


def load_esbr_stage_golden():
    g = np.load(os.path.join(ROOT, "tests", "golden", "esbr_stage_tapped.npz"))
    return {k: g[k] for k in g.files}


def load_esbr_hbe_stage_golden():
    g = np.load(os.path.join(ROOT, "tests", "golden", "esbr_hbe_stage_tapped.npz"))
    return {k: g[k] for k in g.files}


HBE_FRAMES = 6  # consecutive frames in tests/golden/esbr_hbe_stage_tapped.npz


def esbr_hbe_stage_params(g, n_units, f):
    """(hbe_cfg, hf_par, ec_ipar, ec_fpar, rg_par) of frame f (0..5) of the tapped -harmonic_sbr:1 stream, unit u = channel u % 2"""
    j = (np.arange(n_units) % 2) + 2 * (f % HBE_FRAMES)
    h = g["head"][j]
    rg = np.stack([h[:, 7], h[:, 8], 2 * h[:, 9], 0 * h[:, 9]], 1).astype(np.int32)
    return (np.ascontiguousarray(g["hbe_cfg"][j]), np.ascontiguousarray(g["hf_par"][j]), np.ascontiguousarray(g["ec_ipar_in"][j]),
            np.ascontiguousarray(g["ec_fpar"][j]), np.ascontiguousarray(rg))


def esbr_stage_params(g, n_units, f):
    """parameter records of frame f (0..7) of the tapped stream for n_units channel units (unit u = channel u % 2)"""
    j = (np.arange(n_units) % 2) + 2 * (f % 8)
    h = g["head"][j]
    rg = np.stack([h[:, 7], h[:, 8], 2 * h[:, 9], 0 * h[:, 9]], 1).astype(np.int32)
    return (np.ascontiguousarray(g["hf_par"][j]), np.ascontiguousarray(g["ec_ipar_in"][j]), np.ascontiguousarray(g["ec_fpar"][j]),
            np.ascontiguousarray(rg))


def esbr_hbe_bytes(cfg):
    """Algorithmic HBM bytes per unit of ixheaacd_qmf_hbe_apply: the synthesis bank's input cells (32 columns x synth_size bands
    x 8 B), the written phase-vocoder cells (32 rows x (end - start) bands x 8 B) and the instance state read and written once
    (tail + both bank histories 38 x synth_size words, 12 analysis rows x 4 synth_size words, 10 carry rows x 2 (end - start))."""
    s, nb = cfg[:, 0].astype(np.int64), (cfg[:, 3] - cfg[:, 2]).astype(np.int64)
    state = 4 * (37 * s + 12 * 4 * s + 10 * 2 * nb)
    return 32 * s * 8 + 32 * nb * 8 + 2 * state


def cpu_arm_esbr_hbe(n_units, threads, seed, reps=1, min_seconds=0.0):
    """Time ixheaacd_qmf_hbe_apply per unit on host threads (ref_esbr_hbe_apply_batch, oracle/ref_shim_hbe.c)."""
    from tests import oracle_util
    ref = oracle_util.Ref.try_load()
    if ref is None:
        raise SystemExit("bench.py: the harmonic-transposer CPU baseline needs oracle/_ref/libxaac_ref.so (make ref)")
    P = oracle_util.P
    g = np.load(os.path.join(ROOT, "tests", "golden", "esbr_hbe_tapped.npz"))
    k = len(g["cfg"])
    idx = np.arange(n_units) % k
    cfg = np.ascontiguousarray(g["cfg"][idx])
    state = np.ascontiguousarray(g["state_in"][idx])
    qre, qim = np.ascontiguousarray(g["qmf_re"][idx]), np.ascontiguousarray(g["qmf_im"][idx])
    pr, pi = np.zeros_like(qre), np.zeros_like(qim)
    tbl = np.zeros((n_units, 128), np.int16)
    tbl[:, :6] = [1, 1, cfg[0, 2], cfg[0, 3], cfg[0, 2], cfg[0, 3]]
    err = np.zeros(n_units, np.int32)
    bounds = np.linspace(0, n_units, threads + 1).astype(int)

    def work(t):
        a, b = int(bounds[t]), int(bounds[t + 1])
        if b > a:
            ref.lib.ref_esbr_hbe_apply_batch(P(cfg[a:]), P(state[a:]), P(qre[a:]), P(qim[a:]), P(pr[a:]), P(pi[a:]), P(tbl[a:]),
                                             P(err[a:]), b - a)

    def one_pass():
        t0 = time.perf_counter()
        th = [threading.Thread(target=work, args=(t,)) for t in range(threads)]
        for x in th:
            x.start()
        for x in th:
            x.join()
        return time.perf_counter() - t0

    one_pass()
    dt, done = 0.0, 0
    while done < reps or dt < min_seconds:
        dt += one_pass()
        done += 1
    assert not err.any()
    return n_units * done / dt, "reference"


def cpu_arm_xheaac_hbe_chain(n_units, threads, seed, reps=1, min_seconds=0.0):
    """Time the reference's own xHE-AAC chain WITH the harmonic transposer per channel unit on host threads
    (ref_xheaac_hbe_chain_batch, oracle/ref_shim_hbe.c: ixheaacd_fd_frm_dec -> ixheaacd_esbr_analysis_filt_block ->
    ixheaacd_qmf_hbe_apply -> ixheaacd_generate_hf -> ixheaacd_sbr_env_calc -> synthesis -> ixheaacd_samples_sat)."""
    from tests import oracle_util
    ref = oracle_util.Ref.try_load()
    if ref is None:
        raise SystemExit("bench.py: the xHE-AAC chain CPU baseline needs oracle/_ref/libxaac_ref.so (make ref)")
    P = oracle_util.P
    n_units -= n_units & 1
    g = load_esbr_hbe_stage_golden()
    rng = np.random.default_rng(seed)
    sc = rng.integers(10, 20, size=(n_units, 1))
    coef0 = ((rng.random((n_units, 1024)) * 2 - 1) * (2.0 ** sc)).astype(np.int64).astype(np.int32)
    walk = usac_walk(n_units, reps + 1, seed)
    ov = np.zeros((n_units, 1024), np.int32)
    prev = np.zeros(n_units, np.int32)
    ch = np.arange(n_units) % 2
    q6 = np.ascontiguousarray(np.concatenate([g["in0_" + k][ch].reshape(n_units, -1) for k in
                                              ("qmf_re", "qmf_im", "out_re", "out_im", "pv_re", "pv_im")], 1))
    st = {k: np.ascontiguousarray(g["in0_" + k][ch]) for k in ("anal_states", "anal_pos", "synth_states", "synth_pos", "bw_prev",
                                                                "patch", "ec_state", "hbe_state")}
    params = [esbr_hbe_stage_params(g, n_units, f) for f in range(HBE_FRAMES)]
    # frequency tables that make the reference's in-call re-initialisation (bank size 20) reproduce the tapped configuration
    c0 = g["hbe_cfg"][0]
    tbl = np.zeros(128, np.int16)
    tbl[:6] = [1, 1, c0[2], c0[3], c0[2], c0[3]]
    pcm = np.zeros((n_units // 2, 2048, 2), np.int16)
    err = np.zeros(n_units, np.int32)
    bounds = (np.linspace(0, n_units // 2, threads + 1).astype(int)) * 2

    def work(t, coef, seq, shape, pr):
        a, b = int(bounds[t]), int(bounds[t + 1])
        if b > a:
            ref.lib.ref_xheaac_hbe_chain_batch(P(coef), P(ov), P(seq), P(shape), P(prev), P(q6), P(st["anal_states"]),
                                               P(st["anal_pos"]), P(st["synth_states"]), P(st["synth_pos"]), P(st["bw_prev"]),
                                               P(st["patch"]), P(st["ec_state"]), P(st["hbe_state"]), P(pr[0]), P(tbl), P(pr[1]),
                                               P(pr[2]), P(pr[3]), P(pr[4]), P(pcm), P(err), a, b)

    def one_pass(step):
        coef = coef0.copy()
        seq = np.ascontiguousarray(walk[step, :, 0], np.int32)
        shape = np.ascontiguousarray(walk[step, :, 1], np.int32)
        t0 = time.perf_counter()
        th = [threading.Thread(target=work, args=(t, coef, seq, shape, params[step % HBE_FRAMES])) for t in range(threads)]
        for x in th:
            x.start()
        for x in th:
            x.join()
        dt = time.perf_counter() - t0
        prev[:] = shape
        assert not err.any(), f"reference xHE-AAC (HBE) chain returned an error: {sorted(set(err.tolist()))}"
        return dt

    one_pass(0)
    dt, done = 0.0, 0
    while done < reps or dt < min_seconds:
        dt += one_pass(1 + done % reps)
        done += 1
    return n_units * done / dt, "reference"


def spectral_units(n, seed):
    """records for the spectral-stage workload: 2048 distinct seeded elements tiled over the batch — M/S masks on every pair,
    intensity bands on a quarter of the right channels' bands, TNS on about 12 % and PNS on about 25 % of the channels"""
    from tests import oracle_util as ou
    k = min(n, 2048)
    spec, rec = ou.synth_sps_units(k, seed, pns=True, tns_prob=0.12, pns_prob=0.25)
    idx = np.arange(n) % k
    return spec[idx], rec[idx]


def cpu_arm_aac_spectral(n_units, threads, seed, reps=1, min_seconds=0.0):
    """Time ixheaacd_channel_pair_process per element on host threads (ref_channel_pair_process_batch, oracle/ref_shim_sps.c).
    n_units counts elements (stream-frames); like the other arms with units_per_frame = 1 the return value is 2 x elements/s."""
    from tests import oracle_util as ou
    ref = ou.Ref.try_load()
    if ref is None:
        raise SystemExit("bench.py: the spectral-stage CPU baseline needs oracle/_ref/libxaac_ref.so (make ref)")
    P = ou.P
    n = max(threads, n_units)
    spec0, rec = spectral_units(n, seed)
    rec = np.ascontiguousarray(rec)
    seeds = np.arange(n, dtype=np.int32)
    err = np.zeros(n, np.int32)
    bounds = np.linspace(0, n, threads + 1).astype(int)
    ref.lib.ref_channel_pair_process_batch.argtypes = [ctypes.c_int64] + [ctypes.c_void_p] * 4

    def one_pass():
        spec = spec0.copy()

        def work(t):
            a, b = int(bounds[t]), int(bounds[t + 1])
            if b > a:
                ref.lib.ref_channel_pair_process_batch(b - a, P(spec[a:]), P(rec[a:]), P(seeds[a:]), P(err[a:]))

        t0 = time.perf_counter()
        th = [threading.Thread(target=work, args=(t,)) for t in range(threads)]
        for x in th:
            x.start()
        for x in th:
            x.join()
        return time.perf_counter() - t0

    one_pass()
    dt, done = 0.0, 0
    while done < reps or dt < min_seconds:
        dt += one_pass()
        done += 1
    assert not err.any()
    return 2.0 * n * done / dt, "reference"


def sideinfo_records(n, seed):
    """seeded XAAC_SD_* records (tests/oracle_util.synth_sbrdata_records): 4096 distinct elements tiled over the batch"""
    from tests import oracle_util as ou
    k = min(n, 4096)
    rec = ou.synth_sbrdata_records(k, seed)
    return rec[np.arange(n) % k]


def cpu_arm_sbr_sideinfo(n_units, threads, seed, reps=1, min_seconds=0.0):
    """Time ixheaacd_dec_sbrdata per element on host threads (ref_dec_sbrdata_batch, oracle/ref_shim_sd.c: record -> reference
    structs -> the compiled function -> record; the struct rebuild is part of the measured time and stated in `sample`).
    n_units counts elements; like the other arms with units_per_frame = 1 the return value is 2 x elements/s."""
    from tests import oracle_util as ou
    ref = ou.Ref.try_load()
    if ref is None:
        raise SystemExit("bench.py: the side-info CPU baseline needs oracle/_ref/libxaac_ref.so (make ref)")
    P = ou.P
    n = max(threads, n_units)
    rec0 = np.ascontiguousarray(sideinfo_records(n, seed))
    bounds = np.linspace(0, n, threads + 1).astype(int)
    ref.lib.ref_dec_sbrdata_batch.argtypes = [ctypes.c_int64, ctypes.c_void_p]

    def one_pass():
        rec = rec0.copy()

        def work(t):
            a, b = int(bounds[t]), int(bounds[t + 1])
            if b > a:
                ref.lib.ref_dec_sbrdata_batch(b - a, P(rec[a:]))

        t0 = time.perf_counter()
        th = [threading.Thread(target=work, args=(t,)) for t in range(threads)]
        for x in th:
            x.start()
        for x in th:
            x.join()
        return time.perf_counter() - t0

    one_pass()
    dt, done = 0.0, 0
    while done < reps or dt < min_seconds:
        dt += one_pass()
        done += 1
    return 2.0 * n * done / dt, "reference"


def cpu_arm_heaacv2_esbr_chain(n_units, threads, seed, reps=1, min_seconds=0.0):
    """Time the reference's eSBR stage of mono + PS elements (harmonic transposer + float PS, two synthesis banks) per element on
    host threads (ref_heaacv2_esbr_chain_batch, oracle/ref_shim_fps.c).  n_units counts elements (stream-frames) and, like the
    other chain arm with units_per_frame = 1, the return value is 2 x elements/s."""
    from tests import oracle_util as ou
    ref = ou.Ref.try_load()
    if ref is None:
        raise SystemExit("bench.py: the HE-AACv2 eSBR chain CPU baseline needs oracle/_ref/libxaac_ref.so (make ref)")
    n = max(threads, n_units)
    g = load_esbr_hbe_stage_golden()
    pg = np.load(os.path.join(ROOT, "tests", "golden", "esbr_ps_ref.npz"))
    st = ou.heaacv2_esbr_units(g, n)
    q6 = np.ascontiguousarray(np.concatenate([st[k].reshape(n, -1) for k in ou.ESP_KEYS[:6]], 1))
    c0 = g["hbe_cfg"][0]
    tbl = np.zeros(128, np.int16)
    tbl[:6] = [1, 1, c0[2], c0[3], c0[2], c0[3]]
    rng = np.random.default_rng(seed)
    time_in = (rng.standard_normal((n, 1024)) * 4000).astype(np.float32)
    frames = []
    for f in range(HBE_FRAMES):
        hc, hf, ip, fp, rg = esbr_hbe_stage_params(g, n, f)
        par = np.ascontiguousarray(pg["par"][f % 3][np.arange(n) % 4]).copy()
        par[:, 7] = ip[:, 1]
        frames.append((hc, hf, ip, fp, rg, par))
    bounds = np.linspace(0, n, threads + 1).astype(int)

    def one_pass(step):
        hc, hf, ip, fp, rg, par = frames[step % HBE_FRAMES]
        ipar = ip.copy()
        errs = [None] * threads

        def work(t):
            a, b = int(bounds[t]), int(bounds[t + 1])
            if b > a:
                errs[t] = ou.ref_heaacv2_esbr_chain(ref, st, time_in, hc, tbl, hf, ipar, fp, rg, par, a, b, q6=q6)[2][a:b]

        t0 = time.perf_counter()
        th = [threading.Thread(target=work, args=(t,)) for t in range(threads)]
        for x in th:
            x.start()
        for x in th:
            x.join()
        dt = time.perf_counter() - t0
        assert not any(e is not None and e.any() for e in errs), "reference HE-AACv2 eSBR chain returned an error"
        return dt

    one_pass(0)
    dt, done = 0.0, 0
    while done < reps or dt < min_seconds:
        dt += one_pass(1 + done)
        done += 1
    return 2.0 * n * done / dt, "reference"


def cpu_arm_xheaac_chain(n_units, threads, seed, reps=1, min_seconds=0.0):
    """Time the reference's own xHE-AAC chain per channel unit on host threads (ref_xheaac_chain_batch, oracle/ref_shim.c)."""
    from tests import oracle_util
    ref = oracle_util.Ref.try_load()
    if ref is None:
        raise SystemExit("bench.py: the xHE-AAC chain CPU baseline needs oracle/_ref/libxaac_ref.so (make ref)")
    P = oracle_util.P
    n_units -= n_units & 1
    g = load_esbr_stage_golden()
    rng = np.random.default_rng(seed)
    sc = rng.integers(10, 20, size=(n_units, 1))
    coef0 = ((rng.random((n_units, 1024)) * 2 - 1) * (2.0 ** sc)).astype(np.int64).astype(np.int32)
    walk = usac_walk(n_units, reps + 1, seed)
    ov = np.zeros((n_units, 1024), np.int32)
    prev = np.zeros(n_units, np.int32)
    ch = np.arange(n_units) % 2
    q4 = np.ascontiguousarray(np.stack([g["in0_" + k][ch] for k in ("qmf_re", "qmf_im", "out_re", "out_im")], 1))
    st = {k: np.ascontiguousarray(g["in0_" + k][ch]) for k in ("anal_states", "anal_pos", "synth_states", "synth_pos", "bw_prev",
                                                                "patch", "ec_state")}
    params = [esbr_stage_params(g, n_units, f) for f in range(8)]
    pcm = np.zeros((n_units // 2, 2048, 2), np.int16)
    err = np.zeros(n_units, np.int32)
    bounds = (np.linspace(0, n_units // 2, threads + 1).astype(int)) * 2

    def work(t, coef, seq, shape, pr):
        a, b = int(bounds[t]), int(bounds[t + 1])
        if b > a:
            ref.lib.ref_xheaac_chain_batch(P(coef), P(ov), P(seq), P(shape), P(prev), P(q4), P(st["anal_states"]), P(st["anal_pos"]),
                                           P(st["synth_states"]), P(st["synth_pos"]), P(st["bw_prev"]), P(st["patch"]),
                                           P(st["ec_state"]), P(pr[0]), P(pr[1]), P(pr[2]), P(pr[3]), P(pcm), P(err), a, b)

    def one_pass(step):
        coef = coef0.copy()
        seq = np.ascontiguousarray(walk[step, :, 0], np.int32)
        shape = np.ascontiguousarray(walk[step, :, 1], np.int32)
        t0 = time.perf_counter()
        th = [threading.Thread(target=work, args=(t, coef, seq, shape, params[step % 8])) for t in range(threads)]
        for x in th:
            x.start()
        for x in th:
            x.join()
        dt = time.perf_counter() - t0
        prev[:] = shape
        assert not err.any(), "reference xHE-AAC chain returned an error"
        return dt

    one_pass(0)
    dt, done = 0.0, 0
    while done < reps or dt < min_seconds:
        dt += one_pass(1 + done % reps)
        done += 1
    return n_units * done / dt, "reference"


def cpu_arm_esbr_synth(n_units, threads, seed, reps=1, min_seconds=0.0):
    """Time the reference's eSBR synthesis leaves per unit on host threads (ref_esbr_synth64, oracle/ref_shim.c)."""
    from tests import oracle_util
    ref = oracle_util.Ref.try_load()
    P = oracle_util.P
    qmf, fs, pos = oracle_util.synth_esbr_units(n_units, seed)
    fs[:] = 0
    pos[:] = 0
    out = np.zeros((n_units, 2048), np.float32)
    bounds = np.linspace(0, n_units, threads + 1).astype(int)
    if ref is not None:
        kind, fn, pre = "reference", ref.lib.ref_esbr_synth64_batch, []
    else:
        orc = oracle_util.Oracle()
        kind, fn, pre = "port", orc.lib.xo_esbr_synth64_batch, [P(orc.esrom)]

    def work(t):
        a, b = int(bounds[t]), int(bounds[t + 1])
        if b > a:
            fn(*pre, P(qmf[a:b]), P(fs[a:b]), P(pos[a:b]), P(out[a:b]), b - a)

    def one_pass():
        t0 = time.perf_counter()
        th = [threading.Thread(target=work, args=(t,)) for t in range(threads)]
        for x in th:
            x.start()
        for x in th:
            x.join()
        return time.perf_counter() - t0

    one_pass()
    dt, done = 0.0, 0
    while done < reps or dt < min_seconds:
        dt += one_pass()
        done += 1
    return n_units * done / dt, kind


def usac_walk(n_units, n_steps, seed):
    """ics[step][unit] = (window_sequence, window_shape) of a legal USAC FD walk, both channels of a frame alike:
    ~85 % ONLY_LONG, the rest start / short / stop / stop-start runs"""
    rng = np.random.default_rng(seed)
    nf = n_units // 2
    prev = np.zeros(nf, np.uint8)
    out = np.zeros((n_steps, n_units, 2), np.uint8)
    for s_ in range(n_steps):
        r = rng.random(nf)
        nxt = np.zeros(nf, np.uint8)
        longish = (prev == 0) | (prev == 3)
        nxt[longish & (r < 0.05)] = 1
        st = ~longish
        nxt[st] = np.where(r[st] < 0.35, 2, np.where(r[st] < 0.85, 3, 4))
        shape = (rng.random(nf) < 0.5).astype(np.uint8)
        out[s_, 0::2, 0] = nxt
        out[s_, 1::2, 0] = nxt
        out[s_, 0::2, 1] = shape
        out[s_, 1::2, 1] = shape
        prev = nxt
    return out


def cpu_arm_usac(n_units, threads, seed, reps=1, min_seconds=0.0):
    """Time ixheaacd_fd_frm_dec per unit on host threads (unmodified reference through oracle/_ref, or the C port)."""
    from tests import oracle_util
    ref = oracle_util.Ref.try_load()
    P = oracle_util.P
    rng = np.random.default_rng(seed)
    s = rng.integers(12, 28, size=(n_units, 1))
    coef0 = ((rng.random((n_units, 1024)) * 2 - 1) * (2.0 ** s)).astype(np.int64).astype(np.int32)
    walk = usac_walk(n_units, reps + 1, seed)
    ov = np.zeros((n_units, 1024), np.int32)
    out = np.zeros((n_units, 1024), np.int32)
    err = np.zeros(n_units, np.int32)
    prev = np.zeros(n_units, np.int32)
    bounds = np.linspace(0, n_units, threads + 1).astype(int)
    if ref is not None:
        kind, fn, pre = "reference", ref.lib.ref_usac_fd_frm_dec_batch, []
    else:
        orc = oracle_util.Oracle()
        kind, fn, pre = "port", orc.lib.xo_usac_fd_frm_dec_batch, [P(orc.urom)]

    def work(t, coef, seq, shape):
        a, b = int(bounds[t]), int(bounds[t + 1])
        if b > a:
            fn(*pre, P(coef[a:b]), P(ov[a:b]), P(seq[a:b]), P(shape[a:b]), P(prev[a:b]), P(out[a:b]), P(err[a:b]), b - a)

    def one_pass(step):
        coef = coef0.copy()
        seq = np.ascontiguousarray(walk[step, :, 0], np.int32)
        shape = np.ascontiguousarray(walk[step, :, 1], np.int32)
        t0 = time.perf_counter()
        th = [threading.Thread(target=work, args=(t, coef, seq, shape)) for t in range(threads)]
        for x in th:
            x.start()
        for x in th:
            x.join()
        dt = time.perf_counter() - t0
        prev[:] = shape
        return dt

    one_pass(0)
    dt, done = 0.0, 0
    while done < reps or dt < min_seconds:
        dt += one_pass(1 + done % reps)
        done += 1
    return n_units * done / dt, kind


def host_threads():
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


STAGES = {
    "heaacv2_chain": dict(kernel=None, bytes_per_unit=None, units_per_frame=1,
                          stage="IMDCT+OLA -> QMF analysis (WORD32 -> PCM16 hand-over in its load, block-FP bookkeeping in the same "
                                "kernel) -> HF generation -> envelope adjustment -> PS hybrid/decorrelation/rotation -> "
                                "2 x QMF synthesis (fixed-point, bit-exact; 7 launches)",
                          ref_stage="ixheaacd_imdct_process + ixheaacd_sbr_dec (HQ, PS)", cpu=cpu_arm_chain,
                          cpu_units_per_core=128, cpu_reps=12, realtime_fps=21.533, h2d=4096 + 2 + 2464, d2h=8192),
    "heaacv1_stereo_chain": dict(kernel=None, top_kernel="sbr_dec_lp_kernel", bytes_per_unit=None,
                                 stage="per channel: IMDCT+OLA -> fused low-power SBR stage (WORD32 -> PCM16 hand-over in its load, real QMF "
                                       "analysis, LP HF generation, envelope adjustment + alias reduction, real QMF "
                                       "synthesis; fixed-point, bit-exact)",
                                 ref_stage="ixheaacd_imdct_process + ixheaacd_sbr_dec (low power)", cpu=cpu_arm_chain_lp,
                                 cpu_units_per_core=256, cpu_reps=12, realtime_fps=23.4375, h2d=4096 + 2 + 2464, d2h=4096),
    "aac_lc_stereo_imdct_ola": dict(kernel="imdct_ola_kernel", bytes_per_unit=IMDCT_BYTES_PER_UNIT,
                                    stage="IMDCT + window/OLA (fixed-point WORD32, bit-exact)",
                                    ref_stage="ixheaacd_imdct_process", cpu=cpu_arm, cpu_units_per_core=4096,
                                    realtime_fps=43.066, h2d=4096 + 2, d2h=4096 + 1),
    "aac_lc_stereo_output": dict(kernel=None, top_kernel="imdct_ola_kernel", bytes_per_unit=None,
                                 stage="IMDCT + window/OLA (interleaved WORD32) -> peak limiter -> round16 (bit-exact)",
                                 ref_stage="ixheaacd_imdct_process x 2 + ixheaacd_peak_limiter_process + round16",
                                 cpu=cpu_arm_lc_output, cpu_units_per_core=2048, realtime_fps=43.066, h2d=4096 + 2, d2h=2048),
    "usac_fd_imdct": dict(kernel="usac_fd_kernel", bytes_per_unit=USAC_FD_BYTES_PER_UNIT,
                          stage="USAC FD core transform: IMDCT 1024 / 8 x 128 (saturating radix-4 FFT) + windowing + overlap "
                                "(fixed-point WORD32, bit-exact)",
                          ref_stage="ixheaacd_fd_frm_dec", cpu=cpu_arm_usac, cpu_units_per_core=2048, realtime_fps=31.25,
                          h2d=4096 + 2, d2h=4096),
    "esbr_anal32": dict(kernel="esbr_anal_kernel", bytes_per_unit=ESBR_ANAL_BYTES_PER_UNIT,
                        stage="eSBR 32-band QMF analysis: float -> WORD32, 5-tap window with WORD64 accumulation, forward "
                              "modulation (2 x 16-point FFT, 32-bit twiddles), t_cos rotation, -> float (bit-exact)",
                        ref_stage="ixheaacd_esbr_analysis_filt_block", cpu=cpu_arm_esbr_anal, cpu_units_per_core=2048,
                        realtime_fps=15.625, h2d=4096, d2h=8192),
    "esbr_generate_hf": dict(kernel="esbr_hfgen_kernel", bytes_per_unit=None,
                             stage="eSBR float HF generator: chirp factors, 38-slot complex covariance per low band, alpha "
                                   "solve, patch walk + 2nd-order prediction filter, HBE high band (bit-exact floats)",
                             ref_stage="ixheaacd_generate_hf", cpu=cpu_arm_esbr_hfgen, cpu_units_per_core=1024,
                             realtime_fps=15.625, h2d=4 * 10240 + 384, d2h=2 * 10240, dtype="f32"),
    "xheaac_stereo_chain": dict(kernel=None, top_kernel="esbr_synth_kernel", bytes_per_unit=None,
                                stage="USAC FD core transform -> (x 2^-15 in the load) eSBR analysis bank (32-slot codec delay) -> QMF "
                                      "harmonic transposer -> HF generator -> envelope adjuster -> (regrouping in the load) eSBR "
                                      "synthesis bank -> (samples_sat in the store) PCM16",
                                ref_stage="ixheaacd_fd_frm_dec -> eSBR branch of ixheaacd_sbr_dec with hbe_flag = 1 (ixheaacd_qmf_hbe_apply "
                                          "included) -> ixheaacd_samples_sat",
                                cpu=cpu_arm_xheaac_hbe_chain, cpu_units_per_core=128, cpu_reps=6, realtime_fps=15.625,
                                h2d=4096 + 2 + 64 + 384 + 1152 + 1856 + 16, d2h=4096, dtype="int32 + f32/f64"),
    "heaacv2_esbr_chain": dict(kernel=None, top_kernel="esbr_ps_kernel", bytes_per_unit=None, units_per_frame=1,
                               stage="eSBR analysis bank (32-slot codec delay) -> QMF harmonic transposer -> HF generator -> envelope "
                                     "adjuster -> (regrouping + look-ahead in the load) float parametric stereo: hybrid analysis, "
                                     "transient detector, all-pass decorrelator, rotation, hybrid synthesis -> eSBR synthesis bank "
                                     "x 2 (left through the element's bank, right through the second channel's)",
                               ref_stage="eSBR branch of ixheaacd_sbr_dec for a mono + PS element with hbe_flag = 1: "
                                         "ixheaacd_esbr_analysis_filt_block -> ixheaacd_qmf_hbe_apply -> ixheaacd_generate_hf -> "
                                         "ixheaacd_sbr_env_calc -> ixheaacd_esbr_apply_ps -> 2 x synthesis",
                               cpu=cpu_arm_heaacv2_esbr_chain, cpu_units_per_core=64, cpu_reps=6, realtime_fps=21.533,
                               h2d=4096 + 64 + 384 + 1152 + 1856 + 16 + 4096, d2h=2 * 8192, dtype="f32/f64"),
    "aac_lc_spectral": dict(kernel="aac_spectral_kernels", bytes_per_unit=2 * 2 * 4096 + 3712 + 8, units_per_frame=1,
                            stage="M/S + intensity stereo (warp per element), PNS (thread per element: one generator through both "
                                  "channels), TNS (thread per channel: all-pole recursion with saturating accumulation) — three "
                                  "launches timed as one unit",
                            ref_stage="ixheaacd_channel_pair_process", cpu=cpu_arm_aac_spectral, cpu_units_per_core=4096,
                            realtime_fps=43.066, h2d=2 * 4096 + 3712 + 4, d2h=2 * 4096 + 4, dtype="int32"),
    "sbr_sideinfo": dict(kernel="sbr_sideinfo_kernel", bytes_per_unit=2 * 2608, units_per_frame=1,
                         stage="warp per element: frequency-direction delta decoding as a warp prefix sum, time direction and "
                               "low -> high resolution mapping per band, concealment / timing compensation / range-check retry, "
                               "(mantissa | exponent) dequantisation, noise floor, coupled-pair conversion (bit-exact)",
                         ref_stage="ixheaacd_dec_sbrdata", cpu=cpu_arm_sbr_sideinfo, cpu_units_per_core=8192,
                         realtime_fps=21.533, h2d=2608, d2h=2608, dtype="int16"),
    "esbr_hbe": dict(kernel="esbr_hbe_kernel", bytes_per_unit=None,
                     stage="QMF harmonic transposer: critically sampled real synthesis bank, 2x complex analysis bank, stretch-2/3/4 "
                           "products with pitch cross products, phase rotation (bit-exact floats)",
                     ref_stage="ixheaacd_qmf_hbe_apply", cpu=cpu_arm_esbr_hbe, cpu_units_per_core=512,
                     realtime_fps=15.625, h2d=2 * 8192 + 64, d2h=2 * 8192, dtype="f32/f64"),
    "xheaac_plain_stereo_chain": dict(kernel=None, top_kernel="esbr_synth_kernel", bytes_per_unit=None,
                                stage="USAC FD core transform -> (x 2^-15 in the load) eSBR analysis bank -> HF generator -> envelope "
                                      "adjuster -> (regrouping in the load) eSBR synthesis bank -> (samples_sat in the store) PCM16",
                                ref_stage="ixheaacd_fd_frm_dec -> eSBR branch of ixheaacd_sbr_dec -> ixheaacd_samples_sat",
                                cpu=cpu_arm_xheaac_chain, cpu_units_per_core=256, cpu_reps=6, realtime_fps=15.625,
                                h2d=4096 + 2 + 384 + 1152 + 1856 + 16, d2h=4096, dtype="int32 + f32/f64"),
    "esbr_env_calc": dict(kernel="esbr_envcalc_kernel", bytes_per_unit=None,
                          stage="eSBR float envelope adjuster: per-envelope energies, gains / noise / sinusoid levels (double), "
                                "limiter + boost, 5-tap smoothing, noise and sinusoid insertion (bit-exact floats)",
                          ref_stage="ixheaacd_sbr_env_calc", cpu=cpu_arm_esbr_envcalc, cpu_units_per_core=1024,
                          realtime_fps=15.625, h2d=2 * 10240 + 1152 + 1856, d2h=2 * 10240, dtype="f32/f64"),
    "esbr_synth64": dict(kernel="esbr_synth_kernel", bytes_per_unit=ESBR_SYNTH_BYTES_PER_UNIT,
                         stage="eSBR 64-band QMF synthesis: float -> WORD32, inverse modulation (2 x 32-point FFT, 32-bit "
                               "twiddles), 10-tap window with WORD64 accumulation, -> float (bit-exact)",
                         ref_stage="ixheaacd_esbr_inv_modulation + ixheaacd_shiftrountine_with_rnd_hq + "
                                   "ixheaacd_esbr_qmfsyn64_winadd x 32 slots", cpu=cpu_arm_esbr_synth,
                         cpu_units_per_core=1024, realtime_fps=15.625, h2d=16384, d2h=8192),
    "qmf_synth_hq": dict(kernel="qmf_synth_hq_kernel", bytes_per_unit=SYNTH_BYTES_PER_UNIT,
                         stage="complex 64-band QMF synthesis (fixed-point, bit-exact)",
                         ref_stage="ixheaacd_cplx_synt_qmffilt", cpu=cpu_arm_synth, cpu_units_per_core=1024,
                         realtime_fps=21.533, h2d=16384 + 16, d2h=4096),
}


def run_reference_arm(args, rank, world):
    if rank != 0:
        return
    cores = host_threads()
    cfg_idx, frames, desc = WORKLOADS[args.workload]
    stg = STAGES[args.workload]
    sample_units = min(2 * frames, stg["cpu_units_per_core"] * cores)
    sample_units -= sample_units % 2
    t0 = time.perf_counter()
    ups, kind = stg["cpu"](sample_units, cores, 0xAAC0 + cfg_idx, reps=max(1, args.steps), min_seconds=10.0)
    unit_name = "units (frame x channel)"
    if stg.get("units_per_frame", 2) == 1:
        sample_units *= 2  # the chain's CPU arm counts stream-frames and returns 2 x frames/s
        unit_name = "half stream-frames (i.e. %d stream-frames)" % (sample_units // 2)
    fps = ups / 2.0
    line = {
        "impl": "reference", "metric": "decoded_stereo_frames_per_sec", "value": fps, "unit": "frames/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * (sample_units / 2) / fps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": stg.get("dtype", "int32"), "data": "synthetic",
        "config": {"workload": args.workload, "baseline_config": desc, "stage": stg["ref_stage"],
                   "step": f"bounded sample: {sample_units // 2} stereo frames per step on host cores"},
        "cpu_baseline": {"value": fps, "unit": "frames/s", "cores": cores, "kind": kind,
                         "sample": f"{sample_units} {unit_name} per pass, at least {max(1, args.steps)} passes and 10 s "
                                   f"of timed CPU work (+1 warm-up pass), {cores} threads, private state per unit; the flat-record -> reference-struct "
                                   f"refresh of oracle/ref_shim*.c (a few KB of copies per frame: < 1 % of a frame's time for the DSP stages, "
                                   f"a comparable share for sbr_sideinfo, whose function runs ~1 us per element) is inside the timed region"},
        "e2e": {"value": fps, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "wall_s": time.perf_counter() - t0,
    }
    print(json.dumps(line), flush=True)


class ImdctWork:
    def __init__(self, xb, ctx, n_units, steps_total, seed, dev):
        import torch
        self.xb, self.ctx, self.n = xb, ctx, n_units
        self.spec = make_spec_torch(n_units, seed, dev)
        self.walk = torch.from_numpy(sequence_walk(n_units, steps_total, seed)).to(dev)
        self.state = xb.ImdctBatch(n_units, device=dev)
        self.out = torch.empty((n_units, 1024), dtype=torch.int32, device=dev)
        self.adj = torch.empty((n_units,), dtype=torch.int8, device=dev)
        self.nw = steps_total

    def step(self, i, stream):
        self.xb.imdct_process(self.ctx, self.state, self.spec, self.walk[i % self.nw], self.out, self.adj, stream=stream)

    def host_setup(self):
        import torch
        self.h_spec = torch.empty((self.n, 1024), dtype=torch.int32).pin_memory()
        self.h_spec.copy_(self.spec)
        self.h_out = torch.empty((self.n, 1024), dtype=torch.int32).pin_memory()
        self.h_adj = torch.empty((self.n,), dtype=torch.int8).pin_memory()
        self.h_walk = self.walk.cpu().pin_memory()
        self.hstate = self.xb.ImdctHostState(self.ctx, self.n)

    def host_step(self, i):
        self.xb.imdct_process_host(self.ctx, self.hstate, self.h_spec, self.h_walk[i % self.nw], self.h_out, self.h_adj)

    def host_close(self):
        self.hstate.close()


class SynthWork:
    def __init__(self, xb, ctx, n_units, steps_total, seed, dev):
        import torch
        self.xb, self.ctx, self.n = xb, ctx, n_units
        self.matrix, self.params = make_synth_inputs_torch(n_units, seed, dev)
        self.state = xb.QmfSynthBatch(n_units, device=dev)
        self.pcm = torch.empty((n_units, 2048), dtype=torch.int16, device=dev)

    def step(self, i, stream):
        self.xb.cplx_synt_qmffilt(self.ctx, self.state, self.matrix, self.params, self.pcm, stream=stream)

    def host_setup(self):
        import torch
        self.h_matrix = torch.empty((self.n, 32, 128), dtype=torch.int32).pin_memory()
        self.h_matrix.copy_(self.matrix)
        self.h_params = self.params.cpu().pin_memory()
        self.h_pcm = torch.empty((self.n, 2048), dtype=torch.int16).pin_memory()
        self.hstate = self.xb.QmfSynthHostState(self.ctx, self.n)

    def host_step(self, i):
        self.xb.cplx_synt_qmffilt_host(self.ctx, self.hstate, self.h_matrix, self.h_params, self.h_pcm)

    def host_close(self):
        self.hstate.close()


class ChainWork:
    """HE-AACv2 stream-frames: mono core IMDCT -> PCM16 -> SBR stage with PS -> stereo PCM16.  Side info / initial state
    are tiled from a tapped real stream (11 consecutive frames, unit u runs them with phase u mod 11); the core
    spectra are seeded noise with per-unit magnitude and a block-switching walk."""

    def __init__(self, xb, ctx, n_units, steps_total, seed, dev):
        import torch
        self.xb, self.ctx, self.n = xb, ctx, n_units
        side_frames, st0, ps0 = load_chain_golden()
        self.spec = torch.from_numpy(chain_inputs_np(n_units, seed)).to(dev)
        walk = sequence_walk(n_units + (n_units & 1), steps_total, seed)[:, :n_units]
        self.walk = torch.from_numpy(np.ascontiguousarray(walk)).to(dev)
        self.nw = steps_total
        sf = torch.from_numpy(side_frames).to(dev)
        phase = torch.arange(n_units, device=dev)
        self.side = [sf[(phase + f) % 11].contiguous() for f in range(11)]
        self.imdct_state = xb.ImdctBatch(n_units, device=dev)
        self.state = xb.SbrState(ctx, n_units, with_ps=True)
        self.st0, self.ps0 = st0, ps0
        self.state.upload(np.tile(st0, (n_units, 1)), np.tile(ps0, (n_units, 1)))
        self.w32 = torch.empty((n_units, 1024), dtype=torch.int32, device=dev)
        self.adj = torch.empty((n_units,), dtype=torch.int8, device=dev)
        self.pcm = torch.empty((n_units, 2048, 2), dtype=torch.int16, device=dev)
        self.err = torch.zeros((n_units,), dtype=torch.int32, device=dev)

    def step(self, i, stream):
        xb, ctx = self.xb, self.ctx
        xb.imdct_process(ctx, self.imdct_state, self.spec, self.walk[i % self.nw], self.w32, self.adj, stream=stream)
        xb.sbr_dec_w32(ctx, self.state, self.side[i % 11], self.w32, self.adj, self.pcm, self.err, stream=stream)

    def check(self):
        assert int(self.err.abs().max().item()) == 0, "the SBR stage reported an error for some unit"

    def host_setup(self):
        import torch
        self.h_spec = torch.empty((self.n, 1024), dtype=torch.int32).pin_memory()
        self.h_spec.copy_(self.spec)
        self.h_walk = self.walk.cpu().pin_memory()
        self.h_side = [x.cpu().pin_memory() for x in self.side]
        self.h_pcm = torch.empty((self.n, 2048, 2), dtype=torch.int16).pin_memory()
        self.h_imdct = self.xb.ImdctHostState(self.ctx, self.n)
        self.h_state = self.xb.SbrState(self.ctx, self.n, with_ps=True)
        self.h_state.upload(np.tile(self.st0, (self.n, 1)), np.tile(self.ps0, (self.n, 1)))

    def host_step(self, i):
        self.xb.heaac_frame_host(self.ctx, self.h_imdct, self.h_state, self.h_spec, self.h_walk[i % self.nw],
                                 self.h_side[i % 11], self.h_pcm)

    def host_close(self):
        self.h_imdct.close()
        self.h_state.close()
        self.state.close()


class EsbrAnalWork:
    def __init__(self, xb, ctx, n_units, steps_total, seed, dev):
        import torch
        self.xb, self.ctx, self.n = xb, ctx, n_units
        g = torch.Generator(device=dev)
        g.manual_seed(seed)
        amp = torch.exp2(torch.rand((n_units, 1), generator=g, device=dev) * 12 - 12)
        self.x = (torch.rand((n_units, 1024), generator=g, device=dev) * 2 - 1) * amp
        self.state = xb.EsbrAnalBatch(n_units, device=dev)
        self.qmf = torch.zeros((n_units, 32, 128), dtype=torch.float32, device=dev)
        self.err = torch.zeros((n_units,), dtype=torch.int32, device=dev)

    def step(self, i, stream):
        self.xb.esbr_analysis_filt_block(self.ctx, self.state, self.x, self.qmf, self.err, stream=stream)

    def check(self):
        assert int(self.err.abs().max().item()) == 0

    def host_setup(self):
        import torch
        self.h_x = torch.empty((self.n, 1024), dtype=torch.float32).pin_memory()
        self.h_x.copy_(self.x)
        self.h_q = torch.empty((self.n, 32, 128), dtype=torch.float32).pin_memory()
        self.d_x = torch.empty_like(self.x)

    def host_step(self, i):
        import torch
        self.d_x.copy_(self.h_x, non_blocking=True)
        self.xb.esbr_analysis_filt_block(self.ctx, self.state, self.d_x, self.qmf, self.err)
        self.h_q.copy_(self.qmf, non_blocking=True)
        torch.cuda.current_stream().synchronize()

    def host_close(self):
        pass


class EsbrHfgenWork:
    """131 072 units tiled from 1024 distinct seeded ones (tables, borders, modes, QMF history); the destination and the
    chirp-factor state are carried on the device from step to step."""

    def __init__(self, xb, ctx, n_units, steps_total, seed, dev):
        import torch
        from tests import oracle_util
        self.xb, self.ctx, self.n = xb, ctx, n_units
        base = min(1024, n_units)
        d = oracle_util.synth_esbr_hfgen_units(base, seed)
        d["par"][15::16, oracle_util.EHF["INVF_TBL"]:oracle_util.EHF["INVF_TBL"] + 5] = 64
        # keep the units the stage accepts (random tables also produce the reference's own -1 returns, e.g. a 7th patch)
        t0 = {k: torch.from_numpy(np.ascontiguousarray(v)).to(dev) for k, v in d.items()}
        _, e0 = xb.esbr_generate_hf(ctx, t0["src_re"], t0["src_im"], t0["dst_re"].clone(), t0["dst_im"].clone(), t0["par"],
                                    t0["bw_prev"].clone(), pv_re=t0["pv_re"], pv_im=t0["pv_im"])
        keep = np.flatnonzero(e0.cpu().numpy() == 0)
        d = {k: v[keep] for k, v in d.items()}
        base = len(keep)
        reps = (n_units + base - 1) // base
        tile = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev).repeat((reps,) + (1,) * (a.ndim - 1))[:n_units].contiguous()
        self.t = {k: tile(v) for k, v in d.items()}
        self.patch = torch.zeros((n_units, 8), dtype=torch.int32, device=dev)
        self.err = torch.zeros((n_units,), dtype=torch.int32, device=dev)
        self.bytes_per_unit = float(esbr_hfgen_bytes(d["par"]).mean())

    def step(self, i, stream):
        t = self.t
        self.xb.esbr_generate_hf(self.ctx, t["src_re"], t["src_im"], t["dst_re"], t["dst_im"], t["par"], t["bw_prev"],
                                 pv_re=t["pv_re"], pv_im=t["pv_im"], patch_out=self.patch, err=self.err, stream=stream)

    def check(self):
        assert int(self.err.abs().max().item()) == 0

    def host_setup(self):
        import torch
        keys = ("src_re", "src_im", "pv_re", "pv_im", "par")
        self.h_in = {k: torch.empty(self.t[k].shape, dtype=self.t[k].dtype).pin_memory() for k in keys}
        for k in keys:
            self.h_in[k].copy_(self.t[k])
        self.h_out = [torch.empty(self.t[k].shape, dtype=torch.float32).pin_memory() for k in ("dst_re", "dst_im")]

    def host_step(self, i):
        import torch
        for k, h in self.h_in.items():
            self.t[k].copy_(h, non_blocking=True)
        self.step(i, None)
        self.h_out[0].copy_(self.t["dst_re"], non_blocking=True)
        self.h_out[1].copy_(self.t["dst_im"], non_blocking=True)
        torch.cuda.current_stream().synchronize()

    def host_close(self):
        pass


class EsbrEnvcalcWork:
    """131 072 units tiled from 1024 distinct seeded ones; the QMF cells are adjusted in place step after step (the adjuster
    normalises the band energies to the transmitted envelope, so the values stay bounded), history and indices carried."""

    def __init__(self, xb, ctx, n_units, steps_total, seed, dev):
        import torch
        from tests import oracle_util
        self.xb, self.ctx, self.n = xb, ctx, n_units
        base = min(1024, n_units)
        d = oracle_util.synth_esbr_envcalc_units(base, seed)
        E = oracle_util.EEC
        d["ipar"][:, E["NUM_NOISE_ENV"]] = np.where(d["ipar"][:, E["NUM_ENV"]] == 1, 1, 2)
        reps = (n_units + base - 1) // base
        tile = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev).repeat((reps,) + (1,) * (a.ndim - 1))[:n_units].contiguous()
        self.t = {k: tile(v) for k, v in d.items()}
        self.err = torch.zeros((n_units,), dtype=torch.int32, device=dev)
        self.bytes_per_unit = float(esbr_envcalc_bytes(d["ipar"]).mean())

    def step(self, i, stream):
        t = self.t
        self.xb.esbr_env_calc(self.ctx, t["re"], t["im"], t["ipar"], t["fpar"], t["state"], err=self.err, stream=stream)

    def check(self):
        assert int(self.err.abs().max().item()) == 0

    def host_setup(self):
        import torch
        keys = ("re", "im", "ipar", "fpar")
        self.h_in = {k: torch.empty(self.t[k].shape, dtype=self.t[k].dtype).pin_memory() for k in keys}
        for k in keys:
            self.h_in[k].copy_(self.t[k])
        self.h_out = [torch.empty(self.t[k].shape, dtype=torch.float32).pin_memory() for k in ("re", "im")]

    def host_step(self, i):
        import torch
        t = self.t
        for k, h in self.h_in.items():
            t[k].copy_(h, non_blocking=True)
        self.xb.esbr_env_calc(self.ctx, t["re"], t["im"], t["ipar"], t["fpar"], t["state"], err=self.err)
        self.h_out[0].copy_(t["re"], non_blocking=True)
        self.h_out[1].copy_(t["im"], non_blocking=True)
        torch.cuda.current_stream().synchronize()

    def host_close(self):
        pass


class XheaacChainWork:
    """Stereo xHE-AAC frames: unit = one core channel (units 2k / 2k+1 = L / R of stream k): USAC FD core transform -> float
    eSBR stage -> interleaved stereo PCM16, 5 launches per step.  eSBR parameters and initial state are tiled from a tapped
    real stream (8 consecutive frames); core spectra are seeded noise with a window-sequence walk."""

    def __init__(self, xb, ctx, n_units, steps_total, seed, dev):
        import torch
        self.xb, self.ctx, self.n = xb, ctx, n_units
        g = load_esbr_stage_golden()
        gen = torch.Generator(device=dev)
        gen.manual_seed(seed)
        sc = torch.randint(10, 20, (n_units, 1), generator=gen, device=dev).to(torch.float32)
        self.coef = ((torch.rand((n_units, 1024), generator=gen, device=dev) * 2 - 1) * torch.exp2(sc)).to(torch.int32)
        self.walk = torch.from_numpy(usac_walk(n_units, steps_total, seed)).to(dev)
        self.nw = steps_total
        self.core_state = xb.UsacFdBatch(n_units, device=dev)
        self.core = torch.empty((n_units, 1024), dtype=torch.int32, device=dev)
        self.state = xb.EsbrDecBatch(n_units, device=dev)
        ch = np.arange(n_units) % 2
        for k in xb.EsbrDecBatch.SHAPES:
            getattr(self.state, k).copy_(torch.from_numpy(np.ascontiguousarray(g["in0_" + k][ch])))
        self.params = [[torch.from_numpy(a).to(dev) for a in esbr_stage_params(g, n_units, f)] for f in range(8)]
        self.pcm = torch.zeros((n_units // 2, 2048, 2), dtype=torch.int16, device=dev)
        self.err = torch.zeros((4, n_units), dtype=torch.int32, device=dev)
        hf = np.concatenate([esbr_stage_params(g, 2, f)[0] for f in range(8)])
        ec = np.concatenate([esbr_stage_params(g, 2, f)[1] for f in range(8)])
        CHAIN_KERNEL_BYTES.update({
            "usac_fd_kernel": USAC_FD_BYTES_PER_UNIT,
            # stage mode: 4096 core in + 2 x 1280 ring + 2 x 8 rows x 256 B history r/w + 32 rows x 32 bands x 8 B written
            "esbr_anal_kernel": 4096 + 2560 + 8192 + 8192,
            "esbr_hfgen_kernel": float(esbr_hfgen_bytes(hf).mean()) + 8192,      # + the history rows of sbr_qmf_out r/w
            "esbr_envcalc_kernel": float(esbr_envcalc_bytes(ec).mean()),
            # stage mode: 32 rows x 64 bands x 8 B regrouped in + 2 x 5120 state + 4096 PCM16 out
            "esbr_synth_kernel": 16384 + 10240 + 4096,
        })

    def step(self, i, stream):
        xb = self.xb
        xb.usac_fd_frm_dec(self.ctx, self.core_state, self.coef, self.walk[i % self.nw], self.core, stream=stream)
        hf, ip, fp, rg = self.params[i % 8]
        xb.esbr_dec(self.ctx, self.state, self.core, hf, ip, fp, rg, pcm16=self.pcm, ch_fac=2, err=self.err, stream=stream,
                    want_float=False)

    def check(self):
        assert int(self.err.abs().max().item()) == 0

    def host_setup(self):
        import torch
        self.h_coef = torch.empty((self.n, 1024), dtype=torch.int32).pin_memory()
        self.h_coef.copy_(self.coef)
        self.h_walk = self.walk.cpu().pin_memory()
        self.h_params = [[a.cpu().pin_memory() for a in p] for p in self.params]
        self.h_pcm = torch.empty((self.n // 2, 2048, 2), dtype=torch.int16).pin_memory()
        self.d_coef = torch.empty_like(self.coef)
        self.d_ics = torch.empty((self.n, 2), dtype=torch.uint8, device=self.coef.device)
        self.d_params = [torch.empty_like(a) for a in self.params[0]]

    def host_step(self, i):
        import torch
        xb = self.xb
        self.d_coef.copy_(self.h_coef, non_blocking=True)
        self.d_ics.copy_(self.h_walk[i % self.nw], non_blocking=True)
        for d, h in zip(self.d_params, self.h_params[i % 8]):
            d.copy_(h, non_blocking=True)
        xb.usac_fd_frm_dec(self.ctx, self.core_state, self.d_coef, self.d_ics, self.core)
        hf, ip, fp, rg = self.d_params
        xb.esbr_dec(self.ctx, self.state, self.core, hf, ip, fp, rg, pcm16=self.pcm, ch_fac=2, err=self.err, want_float=False)
        self.h_pcm.copy_(self.pcm, non_blocking=True)
        torch.cuda.current_stream().synchronize()

    def host_close(self):
        pass


class XheaacHbeChainWork(XheaacChainWork):
    """BASELINE configs[4]: as XheaacChainWork with the harmonic transposer (hbe_flag = 1), 6 launches per step; parameters and
    initial state tiled from 6 consecutive frames of a tapped -harmonic_sbr:1 stream (bank size 20, stretch 2, pitch 24 / 96)."""

    def __init__(self, xb, ctx, n_units, steps_total, seed, dev):
        import torch
        self.xb, self.ctx, self.n = xb, ctx, n_units
        g = load_esbr_hbe_stage_golden()
        gen = torch.Generator(device=dev)
        gen.manual_seed(seed)
        sc = torch.randint(10, 20, (n_units, 1), generator=gen, device=dev).to(torch.float32)
        self.coef = ((torch.rand((n_units, 1024), generator=gen, device=dev) * 2 - 1) * torch.exp2(sc)).to(torch.int32)
        self.walk = torch.from_numpy(usac_walk(n_units, steps_total, seed)).to(dev)
        self.nw = steps_total
        self.core_state = xb.UsacFdBatch(n_units, device=dev)
        self.core = torch.empty((n_units, 1024), dtype=torch.int32, device=dev)
        self.state = xb.EsbrDecHbeBatch(n_units, device=dev)
        for k in xb.EsbrDecHbeBatch.SHAPES:
            v = torch.from_numpy(np.ascontiguousarray(g["in0_" + k])).to(dev)
            getattr(self.state, k).view((n_units // 2, 2) + tuple(v.shape[1:])).copy_(v.unsqueeze(0).expand((n_units // 2,) + tuple(v.shape)))
        self.params = [[torch.from_numpy(a).to(dev) for a in esbr_hbe_stage_params(g, n_units, f)] for f in range(HBE_FRAMES)]
        self.pcm = torch.zeros((n_units // 2, 2048, 2), dtype=torch.int16, device=dev)
        self.err = torch.zeros((5, n_units), dtype=torch.int32, device=dev)
        hf = np.concatenate([esbr_hbe_stage_params(g, 2, f)[1] for f in range(HBE_FRAMES)])
        ec = np.concatenate([esbr_hbe_stage_params(g, 2, f)[2] for f in range(HBE_FRAMES)])
        CHAIN_KERNEL_BYTES.update({
            "usac_fd_kernel": USAC_FD_BYTES_PER_UNIT,
            # stage mode with the 32-slot delay: 4096 core in + 2 x 1280 ring + 2 x 40 rows x 256 B history r/w (low 32 bands count:
            # 2 x 40 x 128 B x 2) + 32 rows x 32 bands x 8 B written
            "esbr_anal_kernel": 4096 + 2560 + 2 * 40 * 128 * 2 + 8192,
            "esbr_hbe_kernel": float(esbr_hbe_bytes(g["hbe_cfg"]).mean()) + 2 * 8 * 512,   # + ph_vocod history rows r/w
            "esbr_hfgen_kernel": float(esbr_hfgen_bytes(hf).mean()) + 8192,
            "esbr_envcalc_kernel": float(esbr_envcalc_bytes(ec).mean()),
            "esbr_synth_kernel": 16384 + 10240 + 4096,
        })

    def step(self, i, stream):
        xb = self.xb
        xb.usac_fd_frm_dec(self.ctx, self.core_state, self.coef, self.walk[i % self.nw], self.core, stream=stream)
        hc, hf, ip, fp, rg = self.params[i % HBE_FRAMES]
        xb.esbr_dec_hbe(self.ctx, self.state, self.core, hc, hf, ip, fp, rg, pcm16=self.pcm, ch_fac=2, err=self.err, stream=stream,
                        want_float=False)

    def host_setup(self):
        import torch
        self.h_coef = torch.empty((self.n, 1024), dtype=torch.int32).pin_memory()
        self.h_coef.copy_(self.coef)
        self.h_walk = self.walk.cpu().pin_memory()
        self.h_params = [[a.cpu().pin_memory() for a in p] for p in self.params]
        self.h_pcm = torch.empty((self.n // 2, 2048, 2), dtype=torch.int16).pin_memory()
        self.d_coef = torch.empty_like(self.coef)
        self.d_ics = torch.empty((self.n, 2), dtype=torch.uint8, device=self.coef.device)
        self.d_params = [torch.empty_like(a) for a in self.params[0]]

    def host_step(self, i):
        import torch
        xb = self.xb
        self.d_coef.copy_(self.h_coef, non_blocking=True)
        self.d_ics.copy_(self.h_walk[i % self.nw], non_blocking=True)
        for d, h in zip(self.d_params, self.h_params[i % HBE_FRAMES]):
            d.copy_(h, non_blocking=True)
        xb.usac_fd_frm_dec(self.ctx, self.core_state, self.d_coef, self.d_ics, self.core)
        hc, hf, ip, fp, rg = self.d_params
        xb.esbr_dec_hbe(self.ctx, self.state, self.core, hc, hf, ip, fp, rg, pcm16=self.pcm, ch_fac=2, err=self.err, want_float=False)
        self.h_pcm.copy_(self.pcm, non_blocking=True)
        torch.cuda.current_stream().synchronize()


class SpectralWork:
    """AAC-LC channel pairs through the pre-IMDCT spectral stage, in place, the PNS generator state resident."""

    def __init__(self, xb, ctx, n_units, steps_total, seed, dev):
        import torch
        self.xb, self.ctx, self.n = xb, ctx, n_units
        spec, rec = spectral_units(n_units, seed)
        self.spec0 = torch.from_numpy(np.ascontiguousarray(spec)).to(dev)
        self.spec = self.spec0.clone()
        self.rec = torch.from_numpy(np.ascontiguousarray(rec)).to(dev)
        self.seed = torch.arange(n_units, dtype=torch.int32, device=dev)
        self.err = torch.zeros((n_units,), dtype=torch.int32, device=dev)

    def step(self, i, stream):
        self.xb.aac_channel_pair_process(self.ctx, self.spec, self.rec, pns_seed=self.seed, err=self.err, stream=stream)

    def check(self):
        assert int(self.err.abs().max().item()) == 0

    def host_setup(self):
        import torch
        self.h_spec = torch.empty((self.n, 2, 1024), dtype=torch.int32).pin_memory()
        self.h_spec.copy_(self.spec0)
        self.h_rec = self.rec.cpu().pin_memory()
        self.h_seed = self.seed.cpu().pin_memory()
        self.h_out = torch.empty((self.n, 2, 1024), dtype=torch.int32).pin_memory()

    def host_step(self, i):
        import torch
        self.spec.copy_(self.h_spec, non_blocking=True)
        self.rec.copy_(self.h_rec, non_blocking=True)
        self.seed.copy_(self.h_seed, non_blocking=True)
        self.xb.aac_channel_pair_process(self.ctx, self.spec, self.rec, pns_seed=self.seed, err=self.err)
        self.h_out.copy_(self.spec, non_blocking=True)
        self.h_seed.copy_(self.seed, non_blocking=True)
        torch.cuda.current_stream().synchronize()

    def host_close(self):
        pass


class SideinfoWork:
    """SBR elements through the side-info dequantisation; every step starts from the same parsed records (the stage rewrites
    them in place), restored by a device copy outside the kernel's own time but inside the step."""

    def __init__(self, xb, ctx, n_units, steps_total, seed, dev):
        import torch
        self.xb, self.ctx, self.n = xb, ctx, n_units
        self.rec0 = torch.from_numpy(np.ascontiguousarray(sideinfo_records(n_units, seed))).to(dev)
        self.rec = self.rec0.clone()

    def step(self, i, stream):
        self.xb.dec_sbrdata(self.ctx, self.rec, stream=stream)

    def check(self):
        assert int((self.rec[:, 2] != 0).sum().item()) < self.n // 4

    def host_setup(self):
        self.h_rec = self.rec0.cpu().pin_memory()
        self.h_out = self.h_rec.clone().pin_memory()

    def host_step(self, i):
        import torch
        self.rec.copy_(self.h_rec, non_blocking=True)
        self.xb.dec_sbrdata(self.ctx, self.rec)
        self.h_out.copy_(self.rec, non_blocking=True)
        torch.cuda.current_stream().synchronize()

    def host_close(self):
        pass


class Heaacv2EsbrChainWork:
    """Mono + PS elements (unit = one element = one stereo stream-frame): the float eSBR stage with the harmonic transposer and the
    float parametric stereo, 7 launches per step.  eSBR parameters and initial state are tiled from the tapped -harmonic_sbr:1
    stream, PS side records from the golden records of the compiled reference (tests/golden/esbr_ps_ref.npz); the PS state and
    the second synthesis bank start fresh and stay resident."""

    def __init__(self, xb, ctx, n_units, steps_total, seed, dev):
        import torch
        self.xb, self.ctx, self.n = xb, ctx, n_units
        g = load_esbr_hbe_stage_golden()
        pg = np.load(os.path.join(ROOT, "tests", "golden", "esbr_ps_ref.npz"))
        gen = torch.Generator(device=dev)
        gen.manual_seed(seed)
        self.time_in = torch.randn((n_units, 1024), generator=gen, device=dev) * 4000.0
        self.state = xb.EsbrDecPsBatch(n_units, device=dev)
        for k in xb.EsbrDecHbeBatch.SHAPES:
            v = torch.from_numpy(np.ascontiguousarray(g["in0_" + k])).to(dev)
            getattr(self.state, k).view((n_units // 2, 2) + tuple(v.shape[1:])).copy_(v.unsqueeze(0).expand((n_units // 2,) + tuple(v.shape)))
        self.params = []
        for f in range(HBE_FRAMES):
            hc, hf, ip, fp, rg = esbr_hbe_stage_params(g, n_units, f)
            side = np.ascontiguousarray(pg["side"][f % 3][np.arange(n_units) % 4]).copy()
            side.view(np.int32)[:, 7] = ip[:, 1]  # usb = sub_band_end of the element
            self.params.append([torch.from_numpy(a).to(dev) for a in (hc, hf, ip, fp, rg, side)])
        self.out_l = torch.empty((n_units, 2048), dtype=torch.float32, device=dev)
        self.out_r = torch.empty((n_units, 2048), dtype=torch.float32, device=dev)
        self.err = torch.zeros((6, n_units), dtype=torch.int32, device=dev)
        hf = np.concatenate([esbr_hbe_stage_params(g, 2, f)[1] for f in range(HBE_FRAMES)])
        ec = np.concatenate([esbr_hbe_stage_params(g, 2, f)[2] for f in range(HBE_FRAMES)])
        CHAIN_KERNEL_BYTES.update({
            "esbr_anal_kernel": 4096 + 2560 + 2 * 40 * 128 * 2 + 8192,
            "esbr_hbe_kernel": float(esbr_hbe_bytes(g["hbe_cfg"]).mean()) + 2 * 8 * 512,
            "esbr_hfgen_kernel": float(esbr_hfgen_bytes(hf).mean()) + 8192,
            "esbr_envcalc_kernel": float(esbr_envcalc_bytes(ec).mean()),
            # 32 slots x 64 bands x 8 B regrouped in + 6 x 5 look-ahead cells + 4096 side + 2 x 17472 state + 2 x 16384 matrices out
            "esbr_ps_kernel": 16384 + 240 + 4096 + 2 * 17472 + 2 * 16384,
            # plain mode, per launch (two per step): 16384 matrix in + 2 x 5120 state + 8192 float out
            "esbr_synth_kernel": 16384 + 10240 + 8192,
        })

    def step(self, i, stream):
        hc, hf, ip, fp, rg, side = self.params[i % HBE_FRAMES]
        self.xb.esbr_dec_ps(self.ctx, self.state, self.time_in, hc, hf, ip, fp, rg, side, out_l=self.out_l, out_r=self.out_r,
                            err=self.err, stream=stream)

    def check(self):
        assert int(self.err.abs().max().item()) == 0

    def host_setup(self):
        import torch
        self.h_time = torch.empty((self.n, 1024), dtype=torch.float32).pin_memory()
        self.h_time.copy_(self.time_in)
        self.h_params = [[a.cpu().pin_memory() for a in p] for p in self.params]
        self.h_l = torch.empty((self.n, 2048), dtype=torch.float32).pin_memory()
        self.h_r = torch.empty((self.n, 2048), dtype=torch.float32).pin_memory()
        self.d_time = torch.empty_like(self.time_in)
        self.d_params = [torch.empty_like(a) for a in self.params[0]]

    def host_step(self, i):
        import torch
        self.d_time.copy_(self.h_time, non_blocking=True)
        for d, h in zip(self.d_params, self.h_params[i % HBE_FRAMES]):
            d.copy_(h, non_blocking=True)
        hc, hf, ip, fp, rg, side = self.d_params
        self.xb.esbr_dec_ps(self.ctx, self.state, self.d_time, hc, hf, ip, fp, rg, side, out_l=self.out_l, out_r=self.out_r, err=self.err)
        self.h_l.copy_(self.out_l, non_blocking=True)
        self.h_r.copy_(self.out_r, non_blocking=True)
        torch.cuda.current_stream().synchronize()

    def host_close(self):
        pass


class EsbrHbeWork:
    """The harmonic transposer on its own: 16 tapped calls tiled over the batch, state carried from step to step."""

    def __init__(self, xb, ctx, n_units, steps_total, seed, dev):
        import torch
        self.xb, self.ctx, self.n = xb, ctx, n_units
        g = np.load(os.path.join(ROOT, "tests", "golden", "esbr_hbe_tapped.npz"))
        k = len(g["cfg"])
        idx = torch.arange(n_units, device=dev) % k
        self.cfg = torch.from_numpy(g["cfg"]).to(dev)[idx].contiguous()
        self.hb = xb.EsbrHbeBatch(n_units, device=dev)
        self.hb.state.copy_(torch.from_numpy(g["state_in"]).to(dev)[idx])
        gen = torch.Generator(device=dev)
        gen.manual_seed(seed)
        amp = torch.exp2(torch.rand((n_units, 1, 1), generator=gen, device=dev) * 8 + 2)
        self.qre = (torch.randn((n_units, 32, 64), generator=gen, device=dev) * amp)
        self.qim = (torch.randn((n_units, 32, 64), generator=gen, device=dev) * amp)
        self.qre[:, :, 32:] = 0   # the 32-band analysis bank leaves the upper half empty
        self.qim[:, :, 32:] = 0
        self.pr = torch.zeros((n_units, 32, 64), dtype=torch.float32, device=dev)
        self.pi = torch.zeros_like(self.pr)
        self.err = torch.zeros((n_units,), dtype=torch.int32, device=dev)
        self.bytes_per_unit = float(esbr_hbe_bytes(g["cfg"]).mean())

    def step(self, i, stream):
        self.xb.esbr_qmf_hbe_apply(self.ctx, self.hb, self.qre, self.qim, self.pr, self.pi, self.cfg, err=self.err, stream=stream)

    def check(self):
        assert int(self.err.abs().max().item()) == 0

    def host_setup(self):
        import torch
        self.h_qre = torch.empty((self.n, 32, 64), dtype=torch.float32).pin_memory()
        self.h_qim = torch.empty((self.n, 32, 64), dtype=torch.float32).pin_memory()
        self.h_qre.copy_(self.qre)
        self.h_qim.copy_(self.qim)
        self.h_cfg = self.cfg.cpu().pin_memory()
        self.h_pr = torch.empty((self.n, 32, 64), dtype=torch.float32).pin_memory()
        self.h_pi = torch.empty((self.n, 32, 64), dtype=torch.float32).pin_memory()
        self.d_qre, self.d_qim, self.d_cfg = torch.empty_like(self.qre), torch.empty_like(self.qim), torch.empty_like(self.cfg)

    def host_step(self, i):
        import torch
        self.d_qre.copy_(self.h_qre, non_blocking=True)
        self.d_qim.copy_(self.h_qim, non_blocking=True)
        self.d_cfg.copy_(self.h_cfg, non_blocking=True)
        self.xb.esbr_qmf_hbe_apply(self.ctx, self.hb, self.d_qre, self.d_qim, self.pr, self.pi, self.d_cfg, err=self.err)
        self.h_pr.copy_(self.pr, non_blocking=True)
        self.h_pi.copy_(self.pi, non_blocking=True)
        torch.cuda.current_stream().synchronize()

    def host_close(self):
        pass


class EsbrSynthWork:
    def __init__(self, xb, ctx, n_units, steps_total, seed, dev):
        import torch
        self.xb, self.ctx, self.n = xb, ctx, n_units
        g = torch.Generator(device=dev)
        g.manual_seed(seed)
        mag = torch.exp2(torch.rand((n_units, 1, 1), generator=g, device=dev) * 20 - 4)
        self.qmf = (torch.rand((n_units, 32, 128), generator=g, device=dev) * 2 - 1) * mag
        self.state = xb.EsbrSynthBatch(n_units, device=dev)
        self.out = torch.empty((n_units, 2048), dtype=torch.float32, device=dev)
        self.err = torch.zeros((n_units,), dtype=torch.int32, device=dev)

    def step(self, i, stream):
        self.xb.esbr_synthesis_filt(self.ctx, self.state, self.qmf, self.out, self.err, stream=stream)

    def check(self):
        assert int(self.err.abs().max().item()) == 0

    def host_setup(self):
        import torch
        self.h_qmf = torch.empty((self.n, 32, 128), dtype=torch.float32).pin_memory()
        self.h_qmf.copy_(self.qmf)
        self.h_out = torch.empty((self.n, 2048), dtype=torch.float32).pin_memory()
        self.d_qmf = torch.empty_like(self.qmf)

    def host_step(self, i):
        import torch
        self.d_qmf.copy_(self.h_qmf, non_blocking=True)
        self.xb.esbr_synthesis_filt(self.ctx, self.state, self.d_qmf, self.out, self.err)
        self.h_out.copy_(self.out, non_blocking=True)
        torch.cuda.current_stream().synchronize()

    def host_close(self):
        pass


class LcOutputWork:
    """AAC-LC stereo frames with the reference's default flags: units 2k / 2k+1 = L / R of stream k; IMDCT writes the
    interleaved WORD32 time buffer (ch_fac = 2), the limiter (one unit per stream) turns it into PCM16."""

    def __init__(self, xb, ctx, n_units, steps_total, seed, dev):
        import torch
        self.xb, self.ctx, self.n = xb, ctx, n_units
        g = torch.Generator(device=dev)
        g.manual_seed(seed)
        self.spec = make_spec_torch(n_units, seed, dev)
        # a quarter of the streams is loud enough to clip after the qshift scaling: the limiter engages there
        loud = (torch.arange(n_units, device=dev) // 2) % int(os.environ.get("XAAC_LC_LOUD_EVERY", "4")) == 0
        self.spec[loud] = self.spec[loud] << 3
        self.walk = torch.from_numpy(sequence_walk(n_units, steps_total, seed)).to(dev)
        self.nw = steps_total
        self.state = xb.ImdctBatch(n_units, device=dev)
        self.lim = xb.PeakLimiterBatch(n_units // 2, 2, 44100, device=dev)
        self.w32 = torch.empty((n_units // 2, 1024, 2), dtype=torch.int32, device=dev)
        self.adj = torch.empty((n_units,), dtype=torch.int8, device=dev)
        self.pcm = torch.empty((n_units // 2, 1024, 2), dtype=torch.int16, device=dev)
        self.err = torch.zeros((n_units // 2,), dtype=torch.int32, device=dev)

    def step(self, i, stream):
        xb, ctx = self.xb, self.ctx
        xb.imdct_process(ctx, self.state, self.spec, self.walk[i % self.nw], self.w32, self.adj, ch_fac=2, stream=stream)
        xb.peak_limiter_process(ctx, self.lim, self.w32, self.adj.view(self.n // 2, 2), self.pcm, err=self.err, stream=stream)

    def check(self):
        assert int(self.err.abs().max().item()) == 0

    def host_setup(self):
        import torch
        self.h_spec = torch.empty((self.n, 1024), dtype=torch.int32).pin_memory()
        self.h_spec.copy_(self.spec)
        self.h_pcm = torch.empty((self.n // 2, 1024, 2), dtype=torch.int16).pin_memory()
        self.h_walk = self.walk.cpu().pin_memory()
        self.d_spec = torch.empty_like(self.spec)
        self.d_ics = torch.empty((self.n, 2), dtype=torch.uint8, device=self.spec.device)

    def host_step(self, i):
        import torch
        s = torch.cuda.current_stream()
        self.d_spec.copy_(self.h_spec, non_blocking=True)
        self.d_ics.copy_(self.h_walk[i % self.nw], non_blocking=True)
        self.xb.imdct_process(self.ctx, self.state, self.d_spec, self.d_ics, self.w32, self.adj, ch_fac=2, stream=s)
        self.xb.peak_limiter_process(self.ctx, self.lim, self.w32, self.adj.view(self.n // 2, 2), self.pcm, err=self.err, stream=s)
        self.h_pcm.copy_(self.pcm, non_blocking=True)
        s.synchronize()

    def host_close(self):
        pass


class UsacFdWork:
    def __init__(self, xb, ctx, n_units, steps_total, seed, dev):
        import torch
        self.xb, self.ctx, self.n = xb, ctx, n_units
        self.coef = make_spec_torch(n_units, seed, dev)
        self.walk = torch.from_numpy(usac_walk(n_units, steps_total, seed)).to(dev)
        self.state = xb.UsacFdBatch(n_units, device=dev)
        self.out = torch.empty((n_units, 1024), dtype=torch.int32, device=dev)
        self.nw = steps_total

    def step(self, i, stream):
        self.xb.usac_fd_frm_dec(self.ctx, self.state, self.coef, self.walk[i % self.nw], self.out, stream=stream)

    def host_setup(self):
        """no dedicated host-buffer entry point for this stage yet: the end-to-end arm copies through pinned buffers
        around the device call (H2D coefficients + ics, D2H output inside the timed region)"""
        import torch
        self.h_coef = torch.empty((self.n, 1024), dtype=torch.int32).pin_memory()
        self.h_coef.copy_(self.coef)
        self.h_out = torch.empty((self.n, 1024), dtype=torch.int32).pin_memory()
        self.h_walk = self.walk.cpu().pin_memory()
        self.d_coef = torch.empty_like(self.coef)
        self.d_ics = torch.empty((self.n, 2), dtype=torch.uint8, device=self.coef.device)

    def host_step(self, i):
        import torch
        self.d_coef.copy_(self.h_coef, non_blocking=True)
        self.d_ics.copy_(self.h_walk[i % self.nw], non_blocking=True)
        self.xb.usac_fd_frm_dec(self.ctx, self.state, self.d_coef, self.d_ics, self.out)
        self.h_out.copy_(self.out, non_blocking=True)
        torch.cuda.current_stream().synchronize()

    def host_close(self):
        pass


class ChainLpWork:
    """Stereo HE-AACv1 frames on the reference's fixed-point low-power path: unit = one core channel (units 2k / 2k+1 =
    L / R of stream k): IMDCT -> PCM16 -> fused LP SBR stage -> interleaved stereo PCM16.  Side info / initial state are
    tiled from a tapped real stream (12 consecutive frames, stream k runs them with phase k mod 12)."""

    def __init__(self, xb, ctx, n_units, steps_total, seed, dev):
        import torch
        self.xb, self.ctx, self.n = xb, ctx, n_units
        side_frames, st0 = load_chain_lp_golden()
        self.spec = torch.from_numpy(chain_inputs_np(n_units, seed)).to(dev)
        self.walk = torch.from_numpy(sequence_walk(n_units, steps_total, seed)).to(dev)
        self.nw = steps_total
        self.side = [torch.from_numpy(chain_lp_side(side_frames, n_units, f)).to(dev) for f in range(12)]
        self.imdct_state = xb.ImdctBatch(n_units, device=dev)
        self.state = xb.SbrState(ctx, n_units, low_power=True)
        self.st0 = st0
        self.state.upload(np.tile(st0, (n_units // 2, 1)), None)
        self.w32 = torch.empty((n_units, 1024), dtype=torch.int32, device=dev)
        self.adj = torch.empty((n_units,), dtype=torch.int8, device=dev)
        self.p16 = torch.empty((n_units, 1024), dtype=torch.int16, device=dev)
        self.pcm = torch.empty((n_units // 2, 2048, 2), dtype=torch.int16, device=dev)
        self.err = torch.zeros((n_units,), dtype=torch.int32, device=dev)

    def step(self, i, stream):
        xb, ctx = self.xb, self.ctx
        xb.imdct_process(ctx, self.imdct_state, self.spec, self.walk[i % self.nw], self.w32, self.adj, stream=stream)
        xb.sbr_dec_lp_w32(ctx, self.state, self.side[i % 12], self.w32, self.adj, self.pcm, 2, self.err, stream=stream)

    def check(self):
        assert int(self.err.abs().max().item()) == 0, "the LP SBR stage reported an error for some unit"

    def host_setup(self):
        import torch
        self.h_spec = torch.empty((self.n, 1024), dtype=torch.int32).pin_memory()
        self.h_spec.copy_(self.spec)
        self.h_walk = self.walk.cpu().pin_memory()
        self.h_side = [x.cpu().pin_memory() for x in self.side]
        self.h_pcm = torch.empty((self.n // 2, 2048, 2), dtype=torch.int16).pin_memory()
        self.h_imdct = self.xb.ImdctHostState(self.ctx, self.n)
        self.h_state = self.xb.SbrState(self.ctx, self.n, low_power=True)
        self.h_state.upload(np.tile(self.st0, (self.n // 2, 1)), None)

    def host_step(self, i):
        self.xb.heaac_lp_frame_host(self.ctx, self.h_imdct, self.h_state, self.h_spec, self.h_walk[i % self.nw],
                                    self.h_side[i % 12], self.h_pcm, 2)

    def host_close(self):
        self.h_imdct.close()
        self.h_state.close()
        self.state.close()


def numa_bind(local_rank):
    """Bind this process to the CPUs of its GPU's NUMA node (best effort: the intersection with the container's cpuset must
    not be empty), so that the pinned staging buffers of the e2e arm are first-touched on that node.  Returns a description."""
    try:
        import torch
        bus = torch.cuda.get_device_properties(local_rank).pci_bus_id
        dom = getattr(torch.cuda.get_device_properties(local_rank), "pci_domain_id", 0)
        dev_id = torch.cuda.get_device_properties(local_rank).pci_device_id
        path = "/sys/bus/pci/devices/%04x:%02x:%02x.0/numa_node" % (dom, bus, dev_id)
        node = int(open(path).read().strip())
        if node < 0:
            return {"node": node, "bound": False, "why": "no NUMA affinity reported"}
        cl = open("/sys/devices/system/node/node%d/cpulist" % node).read().strip()
        cpus = set()
        for part in cl.split(","):
            lo, _, hi = part.partition("-")
            cpus.update(range(int(lo), int(hi or lo) + 1))
        allowed = os.sched_getaffinity(0)
        use = cpus & allowed
        if not use:
            return {"node": node, "bound": False, "why": "node CPUs %s outside the cpuset (%d CPUs allowed)" % (cl, len(allowed))}
        os.sched_setaffinity(0, use)
        return {"node": node, "bound": True, "cpus": len(use)}
    except Exception as e:  # noqa: BLE001 - purely advisory
        return {"node": None, "bound": False, "why": repr(e)[:120]}


def sharded_io_arm(xb, ctx, dev, rank, world, total_frames, K, seed, barrier):
    """configs[3] "sharded": rank 0 owns the pre-parsed per-frame buffers of ALL streams on its device (spectral coefficients,
    window info, SBR / PS side info) and the PCM must end up there.  Every step: grouped isend / irecv scatter of the inputs
    over NCCL (NVLink 5 / NVSwitch), every rank decodes its contiguous block of streams with its own resident state, grouped
    send / recv gather of the stereo PCM to rank 0 (libxaac_b200/shard.py).  The shard is cut into chunks so that
    the transfer of chunk c + 1 and the return of chunk c - 1 overlap the kernels of chunk c.  Timed with CUDA events on the
    launching stream, max over ranks; also timed: the exchange alone (NVLink rate out of / into rank 0)."""
    import torch
    import torch.distributed as dist
    from libxaac_b200.shard import stream_range
    N, C = total_frames, 4
    a, b = stream_range(N, rank, world)
    n = b - a
    cuts = [stream_range(n, c, C) for c in range(C)]
    works = [ChainWork(xb, ctx, hi - lo, K + 4, seed + 17 * c, dev) for c, (lo, hi) in enumerate(cuts)]
    full = None
    if rank == 0:
        w0 = ChainWork(xb, ctx, 1024, K + 4, seed, dev)  # generator of realistic inputs, tiled over all streams
        rep = (N + 1023) // 1024
        full = {"spec": w0.spec.repeat(rep, 1)[:N].contiguous(),
                "ics": [w0.walk[i].repeat(rep, 1)[:N].contiguous() for i in range(4)],
                "side": [w0.side[i].repeat(rep, 1)[:N].contiguous() for i in range(4)],
                "pcm": torch.empty((N, 2048, 2), dtype=torch.int16, device=dev)}
        w0.state.close()
        del w0
    stream = torch.cuda.current_stream(dev)

    def scatter(c, i):
        ops, copies = [], []
        if rank == 0:
            for r in range(world):
                ra, rb = stream_range(N, r, world)
                lo, hi = stream_range(rb - ra, c, C)
                srcs = (full["spec"][ra + lo:ra + hi], full["ics"][i % 4][ra + lo:ra + hi], full["side"][i % 4][ra + lo:ra + hi])
                if r == 0:
                    copies = srcs
                else:
                    ops += [dist.P2POp(dist.isend, t.view(torch.uint8), r) for t in srcs]  # NCCL moves bytes
            w = works[c]
            w.spec.copy_(copies[0]); w.ics_in = copies[1]; w.side_in = copies[2]
        else:
            w = works[c]
            if not hasattr(w, "ics_buf"):
                w.ics_buf = torch.empty_like(w.walk[0])
                w.side_buf = torch.empty_like(w.side[0])
            w.ics_in, w.side_in = w.ics_buf, w.side_buf
            ops = [dist.P2POp(dist.irecv, t.view(torch.uint8), 0) for t in (w.spec, w.ics_buf, w.side_buf)]
        return dist.batch_isend_irecv(ops) if ops else []

    def compute(c):
        w = works[c]
        xb.imdct_process(ctx, w.imdct_state, w.spec, w.ics_in, w.w32, w.adj, stream=stream)
        xb.sbr_dec_w32(ctx, w.state, w.side_in, w.w32, w.adj, w.pcm, w.err, stream=stream)

    def gather(c):
        ops = []
        if rank == 0:
            for r in range(world):
                ra, rb = stream_range(N, r, world)
                lo, hi = stream_range(rb - ra, c, C)
                if r == 0:
                    full["pcm"][ra + lo:ra + hi].copy_(works[c].pcm)
                else:
                    ops.append(dist.P2POp(dist.irecv, full["pcm"][ra + lo:ra + hi].view(torch.uint8), r))
        else:
            ops.append(dist.P2POp(dist.isend, works[c].pcm.view(torch.uint8), 0))
        return dist.batch_isend_irecv(ops) if ops else []

    def step(i, do_compute=True, overlap=True):
        pend = []
        rq = scatter(0, i)
        for c in range(C):
            nxt = scatter(c + 1, i) if (overlap and c + 1 < C) else None
            for q in rq:
                q.wait()
            if do_compute:
                compute(c)
            pend += gather(c)
            if not overlap and c + 1 < C:
                for q in pend:
                    q.wait()
                pend = []
                nxt = scatter(c + 1, i)
            rq = nxt or []
        for q in pend:
            q.wait()

    def timed_steps(k, **kw):
        for i in range(2):
            step(i, **kw)
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for i in range(k):
            step(2 + i, **kw)
        e1.record(stream)
        barrier()
        t = torch.tensor([e0.elapsed_time(e1) / k], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    ms_overlap = timed_steps(K, overlap=True)
    ms_serial = timed_steps(K, overlap=False)
    ms_comm = timed_steps(K, do_compute=False, overlap=True)
    for w in works:
        assert int(w.err.abs().max().item()) == 0
    ref_pcm = works[0].pcm[:64].clone()

    # ---- the same job over PEER MEMORY instead of NCCL: rank 0 exports its buffers (CUDA IPC), the other ranks map them and
    # (i) their IMDCT / SBR kernels load the spectral coefficients, window info and side info straight from rank 0's HBM over
    # NVLink (the scatter is fused into the kernels' loads, tile by tile), (ii) each chunk's PCM goes back into rank 0's buffer
    # by a copy-engine peer copy on a side stream while the next chunk's kernels run (no SM is spent on the exchange).
    peer_ms = peer_err = None
    cs = torch.cuda.Stream(device=dev)
    pt = None
    try:  # set-up: any rank may fail here; the outcome is agreed on collectively before any timed collective
        shapes = [((N, 1024), torch.int32)] + [((N, 2), torch.uint8)] * 4 + [((N, 1232), torch.int16)] * 4 + [((N, 2048, 2), torch.int16)]
        nbytes = [int(np.prod(sh)) * torch.empty((), dtype=dt).element_size() for sh, dt in shapes]
        objs = [None]
        if rank == 0:  # exportable copies of the buffers (xaac_b200_dev_alloc -> xaac_b200_ipc_export)
            srcs = [full["spec"]] + full["ics"] + full["side"] + [full["pcm"]]
            pt = [ctx.dev_tensor(nb, dt, sh) for (sh, dt), nb in zip(shapes, nbytes)]
            for d_, s_ in zip(pt, srcs):
                d_.copy_(s_)
            torch.cuda.synchronize(dev)
            full["pcm"] = pt[9]
            objs = [[ctx.ipc_export(t) for t in pt]]
        dist.broadcast_object_list(objs, src=0)
        if rank != 0:  # xaac_b200_ipc_import: mapped on THIS rank's device with peer access over NVLink
            pt = [ctx.ipc_import(h, nb, dt, sh) for h, (sh, dt), nb in zip(objs[0], shapes, nbytes)]
    except Exception as e:  # noqa: BLE001
        peer_err = repr(e)[:300]
        sys.stderr.write("[bench] rank %d: peer-memory set-up failed: %s\n" % (rank, peer_err))
    okf = torch.tensor([0 if peer_err else 1], device=dev)
    dist.all_reduce(okf, op=dist.ReduceOp.MIN)
    if int(okf.item()) == 1:
        p_spec, p_ics, p_side, p_pcm = pt[0], pt[1:5], pt[5:9], pt[9]

        cs_in = torch.cuda.Stream(device=dev)
        for w in works:
            w.ics_pull = torch.empty_like(w.walk[0])
            w.side_pull = torch.empty_like(w.side[0])

        def peer_step(i):
            # the small, pointer-chased records (window info, SBR / PS side info: 2.4 KB per stream-frame, read by six kernels)
            # are pulled by the copy engine one chunk ahead; the bulk input (4 KB of spectral coefficients per stream-frame)
            # is loaded by the IMDCT kernel itself from rank 0's HBM
            pulls = []
            with torch.cuda.stream(cs_in):
                for c in range(C):
                    lo, hi = cuts[c]
                    works[c].ics_pull.copy_(p_ics[i % 4][a + lo:a + hi], non_blocking=True)
                    works[c].side_pull.copy_(p_side[i % 4][a + lo:a + hi], non_blocking=True)
                    ev = torch.cuda.Event()
                    ev.record(cs_in)
                    pulls.append(ev)
            for c in range(C):
                lo, hi = cuts[c]
                w = works[c]
                stream.wait_event(pulls[c])
                xb.imdct_process(ctx, w.imdct_state, p_spec[a + lo:a + hi], w.ics_pull, w.w32, w.adj, stream=stream)
                xb.sbr_dec_w32(ctx, w.state, w.side_pull, w.w32, w.adj, w.pcm, w.err, stream=stream)
                ev = torch.cuda.Event()
                ev.record(stream)
                cs.wait_event(ev)
                with torch.cuda.stream(cs):
                    p_pcm[a + lo:a + hi].copy_(w.pcm, non_blocking=True)
            e2 = torch.cuda.Event()
            e2.record(cs)
            stream.wait_event(e2)
            e3 = torch.cuda.Event()
            e3.record(stream)
            cs_in.wait_event(e3)  # the next step's pulls must not overwrite records still in use

        for i in range(2):
            peer_step(i)
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for i in range(K):
            peer_step(2 + i)
        e1.record(stream)
        barrier()
        t = torch.tensor([e0.elapsed_time(e1) / K], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        peer_ms = float(t.item())
        for w in works:
            assert int(w.err.abs().max().item()) == 0
        barrier()
        if rank == 0:  # the last rank's last stream arrived in rank 0's buffer
            assert bool((full["pcm"][N - 64:].view(torch.int32) != 0).any())
    elif peer_err is None:
        peer_err = "set-up failed on another rank"
    del ref_pcm
    per_unit_in, per_unit_out = 4096 + 2 + 2 * 1232, 8192
    remote = N - stream_range(N, 0, world)[1]
    out = {"pattern": "rank 0 owns all pre-parsed buffers: grouped isend/irecv scatter (spectral coefficients, window info, SBR/PS "
                      "side info) -> every rank decodes its contiguous block of streams (state resident) -> grouped send/recv gather "
                      "of the stereo PCM to rank 0; %d chunks per shard, transfers of neighbouring chunks overlap the kernels" % C,
           "collective": "ncclSend/ncclRecv groups (torch.distributed.batch_isend_irecv), no reduction",
           "stereo_frames_total": N, "frames_per_rank": n,
           "scatter_bytes_per_step": remote * per_unit_in, "gather_bytes_per_step": remote * per_unit_out,
           "ms_per_step_overlapped": ms_overlap, "ms_per_step_serial": ms_serial, "ms_per_step_exchange_only": ms_comm,
           "frames_per_s": N / (ms_overlap * 1e-3),
           "nvlink_out_of_rank0_gbs": remote * per_unit_in / (ms_comm * 1e-3) / 1e9,
           "nvlink_into_rank0_gbs": remote * per_unit_out / (ms_comm * 1e-3) / 1e9,
           "nvlink_reference": "measured peer copy 770 GB/s per direction (B200_PROFILING.md), nominal 900",
           "peer_memory": {"pattern": "no NCCL on the data path: rank 0 exports its buffers (xaac_b200_ipc_export), the other ranks map them "
                                      "on their own device (xaac_b200_ipc_import); their IMDCT / SBR "
                                      "IMDCT kernels load the spectral coefficients straight from rank 0's HBM over NVLink (the scatter of the "
                                      "bulk input is fused into the kernel's loads); the small side-info records are pulled and each "
                                      "chunk's PCM is pushed back by copy-engine peer copies that overlap the neighbouring chunks' kernels "
                                      "(no SM is spent on the exchange)",
                           "ms_per_step": peer_ms, "frames_per_s": None if not peer_ms else N / (peer_ms * 1e-3),
                           "error": peer_err}}
    for w in works:
        w.state.close()
    return out


WORK = {"aac_lc_stereo_imdct_ola": ImdctWork, "qmf_synth_hq": SynthWork, "heaacv2_chain": ChainWork,
        "heaacv1_stereo_chain": ChainLpWork, "usac_fd_imdct": UsacFdWork,
        "aac_lc_stereo_output": LcOutputWork, "esbr_synth64": EsbrSynthWork,
        "esbr_anal32": EsbrAnalWork, "esbr_generate_hf": EsbrHfgenWork,
        "esbr_env_calc": EsbrEnvcalcWork, "xheaac_stereo_chain": XheaacHbeChainWork, "esbr_hbe": EsbrHbeWork,
        "xheaac_plain_stereo_chain": XheaacChainWork, "heaacv2_esbr_chain": Heaacv2EsbrChainWork, "aac_lc_spectral": SpectralWork, "sbr_sideinfo": SideinfoWork}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="heaacv2_chain", choices=sorted(WORKLOADS))
    ap.add_argument("--frames", type=int, default=0, help="stereo frames per GPU (default: the config's batch)")
    ap.add_argument("--e2e-steps", type=int, default=0, help="steps for the host-buffer arm (default min(steps,5))")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"],
                    help="N > 1: weak = every rank decodes the config's batch (default, the driver's scaling run); strong = the "
                         "config's batch is split over the ranks by stream (libxaac_b200.shard.stream_range)")
    ap.add_argument("--no-sharded-io", action="store_true",
                    help="N > 1: skip the arm in which rank 0 owns the pre-parsed buffers and scatters / gathers them over NCCL")
    ap.add_argument("--no-extra-stages", action="store_true", help="skip the short extra per-stage roofline runs")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))

    if args.impl == "reference":
        run_reference_arm(args, rank, world)
        return

    import torch
    import torch.distributed as dist
    import libxaac_b200 as xb

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the product path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        import datetime
        dist.init_process_group("nccl", device_id=dev, timeout=datetime.timedelta(seconds=180))

    cfg_idx, frames, desc = WORKLOADS[args.workload]
    stg = STAGES[args.workload]
    if args.frames:
        frames = args.frames
    upf = stg.get("units_per_frame", 2)
    total_frames = frames * world if args.scaling == "weak" else frames
    if args.scaling == "strong" and world > 1:
        from libxaac_b200.shard import stream_range
        a_, b_ = stream_range(frames, rank, world)  # never split a stream: the partition is by stream-frame
        frames = b_ - a_
    n_units = upf * frames  # units (frame x core channel) of this rank; every rank owns its own streams and their state
    numa = numa_bind(local_rank)  # before any pinned allocation: first touch puts the staging buffers on the GPU's node
    K, W = args.steps, args.warmup
    seed = 0xAAC0 + cfg_idx + 1000 * rank
    ctx = xb.Context(local_rank)
    stream = torch.cuda.current_stream(dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def timed(work, k, w, first=0):
        """w warm-up + k timed steps on the launching stream; returns (per-step ms list, total ms, launches)."""
        for s_ in range(w):
            work.step(first + s_, stream)
        barrier()
        l0 = ctx.launch_count
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(k + 1)]
        barrier()
        ev[0].record(stream)
        for s_ in range(k):
            work.step(first + w + s_, stream)
            ev[s_ + 1].record(stream)
        barrier()
        return [ev[i].elapsed_time(ev[i + 1]) for i in range(k)], ev[0].elapsed_time(ev[k]), ctx.launch_count - l0

    # ---- device-resident arm --------------------------------------------------------------------------
    work = WORK[args.workload](xb, ctx, n_units, W + K, seed, dev)
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    step_ms, total_ms, gpu_launches = timed(work, K, W)
    t = torch.tensor([total_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms_max = float(t.item())
    value = total_frames * K / (total_ms_max * 1e-3)
    if getattr(work, "bytes_per_unit", None):
        stg = dict(stg, bytes_per_unit=work.bytes_per_unit)
    kernel_ms = float(np.mean(step_ms))  # single-kernel workloads: one launch per step, step time == launch duration
    if hasattr(work, "check"):
        work.check()
    kernel_table = None
    if stg["kernel"] is None:
        # multi-kernel step: a second pass of K steps with every launch bracketed by CUDA events on the launching
        # stream (xaac_b200_kernel_timing) attributes the step time to the kernels
        ctx.kernel_timing(True)
        for s_ in range(K):
            work.step(W + K + s_, stream)
        kt = ctx.kernel_times()
        ctx.kernel_timing(False)
        tot = sum(v[0] for v in kt.values())
        kernel_table = {}
        for name, (ms, cnt) in kt.items():
            per = ms / cnt
            b = CHAIN_KERNEL_BYTES.get(name)
            kernel_table[name] = {"launch_ms": per, "launches_per_step": cnt / K, "share_of_step": ms / tot,
                                  "bytes_per_unit": b, "units_per_launch": n_units,
                                  "achieved": None if b is None else b * n_units / (per * 1e-3) / 1e9}

    # ---- end-to-end arm: host buffers through the C-ABI ---------------------------------------------------
    Ke = args.e2e_steps or min(K, 20)
    work.host_setup()
    work.host_step(0)  # warm-up (staging allocation, first-touch of the pinned buffers)
    work.host_step(1)
    work.host_step(2)
    barrier()
    t0 = time.perf_counter()
    for s_ in range(Ke):
        work.host_step(3 + s_)
    torch.cuda.synchronize(dev)
    e2e_s = time.perf_counter() - t0
    te = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_value = total_frames * Ke / float(te.item())
    e2e_rank_gbs = n_units * (stg["h2d"] + stg["d2h"]) * Ke / e2e_s / 1e9  # this rank's own host<->device rate
    if world > 1:
        gl = [None] * world
        dist.all_gather_object(gl, {"rank": rank, "h2d_d2h_gbs": e2e_rank_gbs, "numa": numa})
    else:
        gl = [{"rank": 0, "h2d_d2h_gbs": e2e_rank_gbs, "numa": numa}]
    sharded = None
    if world > 1 and args.workload == "heaacv2_chain" and not args.no_sharded_io:
        work.host_close()
        del work
        work = None
        torch.cuda.empty_cache()
        sharded = sharded_io_arm(xb, ctx, dev, rank, world, WORKLOADS[args.workload][1], min(K, 10), seed, barrier)
    clocks = sampler.stop() if rank == 0 else None
    if work is not None:
        work.host_close()
        del work
    torch.cuda.empty_cache()

    if rank == 0:
        peak, peak_src = peaks()
        if kernel_table is not None:
            # the BASELINE metric kernel is the QMF synthesis; it is also the largest share of the chain
            top = stg.get("top_kernel", "qmf_synth_hq_kernel")
            stg = dict(stg, kernel=top, bytes_per_unit=CHAIN_KERNEL_BYTES[top])
            kernel_ms = kernel_table[top]["launch_ms"]
        achieved = stg["bytes_per_unit"] * n_units / (kernel_ms * 1e-3) / 1e9
        line = {
            "metric": "decoded_stereo_frames_per_sec", "value": value, "unit": "frames/s", "n_gpus": world,
            "steps": K, "warmup": W, "ms_per_step": total_ms_max / K, "higher_is_better": True, "scaling": args.scaling,
            "vs_baseline": None, "dtype": stg.get("dtype", "int32"), "data": "synthetic",
            "config": {"workload": args.workload, "baseline_config": desc, "stereo_frames_per_gpu": frames,
                       "units_per_gpu": n_units, "stage": stg["stage"],
                       "l2_policy": "per-step working set > 1 GiB >> 126 MB L2 (no flush needed)",
                       "stereo_frames_total": total_frames,
                       "partition": "by stream, contiguous blocks (libxaac_b200.shard.stream_range); no data-path collective",
                       "realtime_x_per_stream": value / total_frames / stg["realtime_fps"]},
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": "frames/s", "h2d_bytes_per_step": n_units * stg["h2d"],
                    "d2h_bytes_per_step": n_units * stg["d2h"], "steps": Ke,
                    "timer": "host wall clock around the synchronous host-buffer C-ABI call"},
            "gpu_launches": int(gpu_launches),
            "host_io": {"per_rank": gl, "note": "each rank's own pinned-host <-> device rate inside the e2e arm; ranks are bound to "
                                                 "the CPUs of their GPU's NUMA node when the container's cpuset allows it"},
            "roofline": {"bound": "hbm", "kernel": stg["kernel"], "achieved": achieved, "peak": peak,
                         "unit": "GB/s", "frac": achieved / peak,
                         "traffic": (ncu_traffic().get(stg["kernel"], {}).get("bytes_per_unit", 0) * n_units) or None,
                         "traffic_source": ncu_traffic().get(stg["kernel"], {}).get("source"),
                         "traffic_standalone": (ncu_traffic().get(stg["kernel"], {}).get("standalone_bytes_per_unit", 0) * n_units) or None,
                         "traffic_stale": bool(ncu_traffic().get(stg["kernel"], {}).get("stale", False)), "peak_source": peak_src,
                         "bytes_per_unit": stg["bytes_per_unit"], "units_per_launch": n_units,
                         "launch_ms": kernel_ms},
        }
        if kernel_table is not None:
            for v in kernel_table.values():
                v["peak"] = peak
                v["frac"] = None if v["achieved"] is None else v["achieved"] / peak
            line["kernels"] = kernel_table
            line["roofline"]["note"] = ("per-launch duration from CUDA events around every launch of a second pass of "
                                        "the same steps" + ("; two launches per step (left, right)"
                                                            if top == "qmf_synth_hq_kernel" else ""))
            # the kernel that takes the largest share of the step, with its own roofline figures (the `kernel` above is the
            # one BASELINE.json's metric names, which need not be the dominant one)
            dom = max(kernel_table, key=lambda k_: kernel_table[k_]["share_of_step"])
            dv = kernel_table[dom]
            line["roofline"]["dominant_kernel"] = {
                "kernel": dom, "share_of_step": dv["share_of_step"], "launch_ms": dv["launch_ms"], "achieved": dv["achieved"],
                "frac": dv["frac"], "bytes_per_unit": dv["bytes_per_unit"],
                "traffic": (ncu_traffic().get(dom, {}).get("bytes_per_unit", 0) * n_units) or None}
        if sharded is not None:
            line["sharded_io"] = sharded
        if args.workload == "aac_lc_stereo_imdct_ola":
            line["config"]["window_sequence_mix"] = "walk: ~90% long, 4% start, 4% stop, 2% short"
        # short extra runs of the other stage kernels so every hot kernel has a live roofline number
        if not args.no_extra_stages and world == 1:
            extra = {}
            for name in WORK:
                if name == args.workload:
                    continue
                if STAGES[name]["kernel"] is None:
                    continue
                w2 = WORK[name](xb, ctx, 131072, 8, seed, dev)
                ms, _, _ = timed(w2, 5, 3)
                bpu = getattr(w2, "bytes_per_unit", None) or STAGES[name]["bytes_per_unit"]
                ach = bpu * 131072 / (float(np.mean(ms)) * 1e-3) / 1e9
                extra[name] = {"kernel": STAGES[name]["kernel"], "launch_ms": float(np.mean(ms)),
                               "units_per_launch": 131072, "bytes_per_unit": bpu,
                               "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak,
                               "stereo_frames_per_sec": 65536 / (float(np.mean(ms)) * 1e-3)}
                del w2
                torch.cuda.empty_cache()
            line["stage_rooflines"] = extra
            # the other BASELINE configs at their own batch sizes, device-resident and through host buffers (short runs), so that the
            # driver-run line carries them too and not only the builder's own captures
            other = {}
            for name in ("aac_lc_stereo_imdct_ola", "heaacv1_stereo_chain", "xheaac_stereo_chain", "heaacv2_esbr_chain"):
                if name == args.workload:
                    continue
                ci, fr, _ = WORKLOADS[name]
                upf2 = STAGES[name].get("units_per_frame", 2)
                w2 = WORK[name](xb, ctx, upf2 * fr, 8, 0xAAC0 + ci, dev)
                ms, tot, nl = timed(w2, 5, 3)
                if hasattr(w2, "check"):
                    w2.check()
                w2.host_setup()
                for s_ in range(2):
                    w2.host_step(s_)
                torch.cuda.synchronize(dev)
                t0 = time.perf_counter()
                for s_ in range(3):
                    w2.host_step(2 + s_)
                torch.cuda.synchronize(dev)
                dt = time.perf_counter() - t0
                other[name] = {"baseline_config_index": ci, "stereo_frames": fr, "value": fr * 5 / (tot * 1e-3), "unit": "frames/s",
                               "ms_per_step": tot / 5, "steps": 5, "warmup": 3, "gpu_launches": int(nl),
                               "e2e": {"value": fr * 3 / dt, "unit": "frames/s", "steps": 3,
                                       "h2d_bytes_per_step": upf2 * fr * STAGES[name]["h2d"],
                                       "d2h_bytes_per_step": upf2 * fr * STAGES[name]["d2h"]},
                               "realtime_x_per_stream": (fr * 5 / (tot * 1e-3)) / fr / STAGES[name]["realtime_fps"]}
                w2.host_close()
                del w2
                torch.cuda.empty_cache()
            line["other_configs"] = other
        if not args.no_cpu_baseline and world == 1:
            cores = host_threads()
            stg = STAGES[args.workload]
            creps = stg.get("cpu_reps", 2)
            ups1, kind = stg["cpu"](stg["cpu_units_per_core"], 1, 0xAAC0 + cfg_idx, reps=max(1, creps // 4))
            upsN, kind = stg["cpu"](stg["cpu_units_per_core"] * cores, cores, 0xAAC0 + cfg_idx, reps=creps, min_seconds=10.0)
            line["cpu_baseline"] = {"value": upsN / 2.0, "unit": "frames/s", "cores": cores, "kind": kind,
                                    "value_1core": ups1 / 2.0,
                                    "sample": f"{stg['cpu_units_per_core'] * cores} units per pass, >= {creps} passes and >= 10 s of "
                                              f"timed CPU work on {cores} threads "
                                              f"(1-core figure: {stg['cpu_units_per_core']} units), "
                                              f"{stg['ref_stage']} per unit"}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
