#!/usr/bin/env python
"""Generate the golden fixtures under tests/golden/ from the UNMODIFIED compiled reference.

Pipeline (all binaries come from `make ref`, i.e. gcc on the sources where they lie in /root/reference):
  synthetic WAV (numpy)  ->  oracle/_ref/xaacenc (reference encoder)  ->  bitstream
  bitstream              ->  oracle/_ref/xaacdec_tap (reference decoder + ld --wrap stage taps, oracle/ref_taps.c)
                         ->  per-call records (inputs, state-before, outputs, state-after)
  records                ->  tests/golden/*.npz (a small, window-sequence-balanced selection)

The reference ships no golden vectors for this path; these fixtures are what pins the oracle and
the CUDA kernels on the GPU box, where /root/reference does not exist.  Run: python tools/make_golden.py
"""
import os
import subprocess
import sys
import tempfile
import wave

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REFDIR = os.path.join(ROOT, "oracle", "_ref")
GOLD = os.path.join(ROOT, "tests", "golden")


def synth(fs, seconds, channels, seed):
    """Tonal + noise + click bursts + silence + full-scale segment: exercises long, start/stop and short blocks."""
    rng = np.random.default_rng(seed)
    n = int(fs * seconds)
    t = np.arange(n) / fs
    x = 0.35 * np.sin(2 * np.pi * 440 * t) + 0.2 * np.sin(2 * np.pi * 3520 * t + 1.0)
    x += 0.05 * rng.standard_normal(n)
    # castanet-like clicks every ~0.37 s force block switching
    for k in range(int(seconds / 0.37)):
        p = int((0.2 + 0.37 * k) * fs)
        ln = min(300, n - p)
        if ln > 0:
            x[p:p + ln] += 0.9 * rng.standard_normal(ln) * np.exp(-np.arange(ln) / 40.0)
    x[int(0.45 * n):int(0.5 * n)] = 0.0  # silence
    seg = slice(int(0.8 * n), int(0.85 * n))
    x[seg] = np.sign(np.sin(2 * np.pi * 1000 * t[seg]))  # full-scale square
    x = np.clip(x, -1.0, 1.0)
    chans = [x]
    if channels == 2:
        y = np.roll(x, 37) * 0.8 + 0.1 * np.sin(2 * np.pi * 997 * t)
        chans.append(np.clip(y, -1, 1))
    pcm = (np.stack(chans, axis=1) * 32767.0).astype(np.int16)
    return pcm


def write_wav(path, pcm, fs):
    with wave.open(path, "wb") as w:
        w.setnchannels(pcm.shape[1])
        w.setsampwidth(2)
        w.setframerate(fs)
        w.writeframes(pcm.tobytes())


def run(cmd, env=None):
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, env=env)
    if r.returncode != 0:
        sys.stderr.write(r.stdout.decode(errors="replace")[-2000:])
        raise RuntimeError("command failed: " + " ".join(cmd))
    return r.stdout.decode(errors="replace")


def encode(wav, out, aot, br, extra=()):
    run([os.path.join(REFDIR, "xaacenc"), f"-ifile:{wav}", f"-ofile:{out}", f"-aot:{aot}", "-adts:1", f"-br:{br}", *extra])


def decode_tap(bitstream, wav_out, tapfile, dec_args=(), tap_max=100000, stages="imd"):
    env = dict(os.environ, XAAC_TAP_FILE=tapfile, XAAC_TAP_MAX=str(tap_max), XAAC_TAP_STAGES=stages)
    run([os.path.join(REFDIR, "xaacdec_tap"), f"-ifile:{bitstream}", f"-ofile:{wav_out}", *dec_args], env=env)


def read_imdct_records(path):
    recs = []
    rec_words = 9 + 1024 + 512 + 1024 + 512
    raw = np.fromfile(path, dtype=np.int32)
    assert raw.size % rec_words == 0, (raw.size, rec_words)
    raw = raw.reshape(-1, rec_words)
    for r in raw:
        assert r[0] == 0x31444D49
        recs.append(dict(hdr=r[1:9].copy(), spec=r[9:1033].copy(), ovl_in=r[1033:1545].copy(),
                         out=r[1545:2569].copy(), ovl_out=r[2569:3081].copy()))
    return recs


def select_balanced(recs, per_class):
    """keep up to per_class records for every (prev_seq, win_seq) transition that occurs."""
    buckets = {}
    for i, r in enumerate(recs):
        key = (int(r["hdr"][4]), int(r["hdr"][5]))
        buckets.setdefault(key, []).append(i)
    keep = []
    for key, idx in sorted(buckets.items()):
        step = max(1, len(idx) // per_class)
        keep += idx[::step][:per_class]
    return sorted(keep), {k: len(v) for k, v in buckets.items()}


def make_imdct_golden(tmp):
    out = {}
    all_recs = []
    for fs, ch, seed, br in ((48000, 1, 11, 64000), (44100, 2, 12, 128000)):
        wav = os.path.join(tmp, f"in_{fs}_{ch}.wav")
        write_wav(wav, synth(fs, 6.0, ch, seed), fs)
        aac = os.path.join(tmp, f"lc_{fs}_{ch}.aac")
        encode(wav, aac, 2, br)
        tap = os.path.join(tmp, f"lc_{fs}_{ch}.tap")
        decode_tap(aac, os.path.join(tmp, "o.wav"), tap, ["-peak_limiter_off:1"])
        recs = read_imdct_records(tap)
        print(f"AAC-LC {fs} Hz {ch}ch: {len(recs)} imdct_process calls tapped")
        all_recs += recs
    keep, hist = select_balanced(all_recs, 6)
    print("transition histogram (prev_seq, win_seq):", hist)
    sel = [all_recs[i] for i in keep]
    # config 1 of BASELINE.json ("AAC-LC mono 48 kHz long-block: 1024-pt IMDCT + sine-window OLA, 1 frame on the
    # reference CPU path").  The reference encoder always signals KBD for long blocks, so the sine-window case is
    # produced by re-running the compiled reference stage (ref_imdct_process -> ixheaacd_imdct_process) on a tapped
    # mono 48 kHz long->long frame's own spectrum and overlap with window_shape forced to sine.
    import ctypes
    ref = ctypes.CDLL(os.path.join(REFDIR, "libxaac_ref.so"))
    src = next(r for r in all_recs if r["hdr"][2] == 1 and r["hdr"][4] == 0 and r["hdr"][5] == 0
               and np.abs(r["spec"].astype(np.int64)).max() > 1 << 20)
    spec = src["spec"].copy()
    ovl = src["ovl_in"].copy()
    o = np.zeros(1024, np.int32)
    ps, pq = ctypes.c_int32(0), ctypes.c_int32(0)
    P = lambda a: a.ctypes.data_as(ctypes.c_void_p)
    adj = ref.ref_imdct_process(P(spec), P(ovl), ctypes.byref(ps), ctypes.byref(pq), 0, 0, P(o), 1)
    first = dict(hdr=np.array([1024, 2, 1, 0, 0, 0, 0, adj], np.int32), spec=src["spec"].copy(),
                 ovl_in=src["ovl_in"].copy(), out=o, ovl_out=ovl)
    sel = [first] + sel
    # hdr columns: frame_length, aot, ch_fac, prev_shape, prev_seq, win_seq, win_shape, qshift_adj
    out["hdr"] = np.stack([r["hdr"] for r in sel]).astype(np.int32)
    for k in ("spec", "ovl_in", "out", "ovl_out"):
        out[k] = np.stack([r[k] for r in sel]).astype(np.int32)
    path = os.path.join(GOLD, "imdct_tapped.npz")
    np.savez_compressed(path, **out)
    print(f"wrote {path}: {len(sel)} records, {os.path.getsize(path)} bytes")


HFG_REC_BYTES = 4 + 160 + 24 + 1024 + 2 * 38 * 128 * 4 + 24 + 4


def read_hfg_records(path):
    """records written by __wrap_ixheaacd_hf_generator (oracle/ref_taps.c)"""
    raw = np.fromfile(path, dtype=np.uint8)
    assert raw.size % HFG_REC_BYTES == 0, (raw.size, HFG_REC_BYTES)
    out = []
    for r in raw.reshape(-1, HFG_REC_BYTES):
        assert r[0:4].view(np.int32)[0] == 0x31474648
        o = 188 + 1024
        out.append(dict(prm=r[4:164].view(np.int16).copy(), bw_in=r[164:188].view(np.int32).copy(),
                        lpc=r[188:o].view(np.int32).copy().reshape(2, 128),
                        m_in=r[o:o + 19456].view(np.int32).copy().reshape(38, 128),
                        m_out=r[o + 19456:o + 38912].view(np.int32).copy().reshape(38, 128),
                        bw_out=r[o + 38912:o + 38936].view(np.int32).copy(),
                        hb=int(r[o + 38936:o + 38940].view(np.int32)[0])))
    return out


def make_hfgen_golden(tmp):
    """HE-AACv2 (mono + PS, 44.1 kHz, 32 kb/s) and HE-AACv1 mono (48 kHz) streams decoded with -esbr:0 take the complex
    HQ SBR path (decoder/ixheaacd_sbrdecoder.c:408-419); every ixheaacd_hf_generator call is tapped."""
    recs = []
    for fs, ch, aot, br, seed in ((44100, 2, 29, 32000, 13), (48000, 1, 5, 32000, 14)):
        wav = os.path.join(tmp, f"in_{fs}_{ch}_{aot}.wav")
        write_wav(wav, synth(fs, 5.0, ch, seed), fs)
        aac = os.path.join(tmp, f"he_{aot}.aac")
        encode(wav, aac, aot, br)
        tap = os.path.join(tmp, f"he_{aot}.tap")
        decode_tap(aac, os.path.join(tmp, "o.wav"), tap, ["-esbr:0"], stages="hfg")
        r = read_hfg_records(tap) if os.path.exists(tap) and os.path.getsize(tap) else []
        print(f"aot {aot} {fs} Hz {ch}ch: {len(r)} hf_generator calls tapped")
        recs += r
    # keep a spread: first frames (start-up state), every ~9th afterwards, all distinct parameter rows
    keep, seen = [], set()
    for i, r in enumerate(recs):
        key = r["prm"].tobytes()
        if i < 4 or i % 9 == 0 or key not in seen:
            keep.append(i)
        seen.add(key)
    keep = keep[:36]
    sel = [recs[i] for i in keep]
    out = {k: np.stack([r[k] for r in sel]) for k in ("prm", "bw_in", "lpc", "m_in", "m_out", "bw_out")}
    out["hb"] = np.array([r["hb"] for r in sel], np.int32)
    path = os.path.join(GOLD, "hfgen_tapped.npz")
    np.savez_compressed(path, **out)
    print(f"wrote {path}: {len(sel)} records, {os.path.getsize(path)} bytes")


ENV_REC_BYTES = 4 + 656 * 2 + 16 + 464 + 2 * 38 * 128 * 4 + 16 + 464 + 4


def read_env_records(path):
    """records written by __wrap_ixheaacd_calc_sbrenvelope (oracle/ref_taps.c)"""
    raw = np.fromfile(path, dtype=np.uint8)
    assert raw.size % ENV_REC_BYTES == 0, (raw.size, ENV_REC_BYTES)
    out = []
    for r in raw.reshape(-1, ENV_REC_BYTES):
        assert r[0:4].view(np.int32)[0] == 0x31564E45
        o = 4
        d = {}
        for name, nbytes, dt, shape in (("prm", 1312, np.int16, (656,)), ("sf_in", 16, np.int16, (8,)),
                                        ("st_in", 464, np.int16, (232,)), ("m_in", 19456, np.int32, (38, 128)),
                                        ("m_out", 19456, np.int32, (38, 128)), ("sf_out", 16, np.int16, (8,)),
                                        ("st_out", 464, np.int16, (232,)), ("err", 4, np.int32, (1,))):
            d[name] = r[o:o + nbytes].view(dt).copy().reshape(shape)
            o += nbytes
        out.append(d)
    return out


def he_streams(tmp, stages, reader):
    """HE-AACv2 (mono + PS, 44.1 kHz, 32 kb/s) and HE-AACv1 mono (48 kHz) decoded with -esbr:0: both take the complex
    HQ SBR path (decoder/ixheaacd_sbrdecoder.c:408-419)."""
    recs = []
    for fs, ch, aot, br, seed in ((44100, 2, 29, 32000, 13), (48000, 1, 5, 32000, 14)):
        wav = os.path.join(tmp, f"in_{fs}_{ch}_{aot}.wav")
        write_wav(wav, synth(fs, 5.0, ch, seed), fs)
        aac = os.path.join(tmp, f"he_{aot}.aac")
        encode(wav, aac, aot, br)
        tap = os.path.join(tmp, f"he_{aot}_{stages}.tap")
        decode_tap(aac, os.path.join(tmp, "o.wav"), tap, ["-esbr:0"], stages=stages)
        r = reader(tap) if os.path.exists(tap) and os.path.getsize(tap) else []
        print(f"aot {aot} {fs} Hz {ch}ch: {len(r)} {stages} calls tapped")
        recs += r
    return recs


def make_envcalc_golden(tmp):
    """every ixheaacd_calc_sbrenvelope call of the two HQ streams is tapped; keep the start-up frames, every frame
    with a transient / more than one envelope / added harmonics, and a thin spread of the rest."""
    recs = he_streams(tmp, "env", read_env_records)
    keep = []
    for i, r in enumerate(recs):
        p = r["prm"]
        special = p[12] > 1 or p[13] >= 0 or p[151:207].any() or r["st_in"][169] != 0
        if special or i % 12 == 0:
            keep.append(i)
    keep = keep[:40]
    sel = [recs[i] for i in keep]
    out = {k: np.stack([r[k] for r in sel]) for k in sel[0]}
    path = os.path.join(GOLD, "envcalc_tapped.npz")
    np.savez_compressed(path, **out)
    print(f"wrote {path}: {len(sel)} of {len(recs)} records, {os.path.getsize(path)} bytes")


SBR_SIDE, SBR_ST, SBR_PS = 1232, 3920, 3888
SBR_REC_BYTES = 4 + 28 + 2 * (SBR_SIDE + SBR_ST + SBR_PS + 1024 + SBR_ST + SBR_PS + 2048 + 2048)


def read_sbr_records(path):
    """records written by __wrap_ixheaacd_sbr_dec (oracle/ref_taps.c)"""
    raw = np.fromfile(path, dtype=np.uint8)
    assert raw.size % SBR_REC_BYTES == 0, (raw.size, SBR_REC_BYTES)
    out = []
    for r in raw.reshape(-1, SBR_REC_BYTES):
        assert r[0:4].view(np.int32)[0] == 0x31524253
        d = {"hdr": r[4:32].view(np.int32).copy()}
        o = 32
        for name, n in (("side", SBR_SIDE), ("st_in", SBR_ST), ("ps_in", SBR_PS), ("tin", 1024), ("st_out", SBR_ST),
                        ("ps_out", SBR_PS), ("out_l", 2048), ("out_r", 2048)):
            d[name] = r[o:o + 2 * n].view(np.int16).copy()
            o += 2 * n
        out.append(d)
    return out


def make_sbrdec_golden(tmp):
    """whole-stage ixheaacd_sbr_dec records: a run of 12 CONSECUTIVE frames from the start of each stream (so the tests
    can also carry the state from frame to frame) plus a spread of later frames."""
    recs = he_streams(tmp, "sbr", read_sbr_records)
    n_ps = sum(1 for r in recs if r["side"][737])
    keep = list(range(0, 12)) + list(range(20, n_ps, 16)) + list(range(n_ps, n_ps + 12)) + \
        list(range(n_ps + 20, len(recs), 18))
    sel = [recs[i] for i in keep]
    out = {k: np.stack([r[k] for r in sel]) for k in sel[0]}
    out["index"] = np.array(keep, np.int32)
    path = os.path.join(GOLD, "sbrdec_tapped.npz")
    np.savez_compressed(path, **out)
    print(f"wrote {path}: {len(sel)} of {len(recs)} records ({n_ps} with PS), {os.path.getsize(path)} bytes")


def make_sbrdec_lp_golden(tmp):
    """whole-stage records of the low-power (real-valued) path: HE-AACv1 STEREO 48 kHz decoded with -esbr:0 (two
    ixheaacd_sbr_dec calls per frame, one per channel; records alternate between the channels)."""
    fs, ch, aot, br, seed = 48000, 2, 5, 64000, 15
    wav = os.path.join(tmp, "in_v1st.wav")
    write_wav(wav, synth(fs, 5.0, ch, seed), fs)
    aac = os.path.join(tmp, "he_v1st.aac")
    encode(wav, aac, aot, br)
    tap = os.path.join(tmp, "he_v1st.tap")
    decode_tap(aac, os.path.join(tmp, "o.wav"), tap, ["-esbr:0"], stages="slp")
    recs = read_sbr_records(tap)
    assert all(r["hdr"][5] == 1 for r in recs)
    keep = list(range(0, 26)) + list(range(30, len(recs), 14))
    sel = [recs[i] for i in keep]
    out = {k: np.stack([r[k] for r in sel]) for k in ("hdr", "side", "st_in", "tin", "st_out", "out_l")}
    out["index"] = np.array(keep, np.int32)
    path = os.path.join(GOLD, "sbrdec_lp_tapped.npz")
    np.savez_compressed(path, **out)
    print(f"wrote {path}: {len(sel)} of {len(recs)} records, {os.path.getsize(path)} bytes")


def make_usac_fd_golden(tmp):
    """USAC (xHE-AAC, aot 42, ccfl 1024) stereo 32 kHz: every ixheaacd_fd_frm_dec call of a real decode is tapped
    (oracle/ref_taps_usac.c).  Records of pure FD frames (previous frame FD, no FAC data, no concealment) are kept:
    a window-sequence-balanced selection plus a run of consecutive frames of both channels."""
    fs, ch, br, seed = 32000, 2, 64000, 16
    wav = os.path.join(tmp, "in_usac.wav")
    write_wav(wav, synth(fs, 8.0, ch, seed), fs)
    mp4 = os.path.join(tmp, "usac.mp4")
    run([os.path.join(REFDIR, "xaacenc"), f"-ifile:{wav}", f"-ofile:{mp4}", "-aot:42", f"-br:{br}", "-ccfl_idx:3"])
    meta = os.path.join(tmp, "usac.txt")
    tap = os.path.join(tmp, "usac.tap")
    decode_tap(mp4, os.path.join(tmp, "o.wav"), tap, [f"-imeta:{meta}", "-mp4:1"], stages="ufd")
    rec_words = 9 + 4 * 1024
    raw = np.fromfile(tap + ".ufd", dtype=np.int32)
    assert raw.size and raw.size % rec_words == 0, (raw.size, rec_words)
    raw = raw.reshape(-1, rec_words)
    assert (raw[:, 0] == 0x31444655).all()
    hdr = raw[:, 1:9]
    fd = (hdr[:, 4] == 0) & (hdr[:, 5] == 0) & (hdr[:, 6] == 0) & (hdr[:, 7] == 0)
    print(f"USAC {fs} Hz {ch}ch: {len(raw)} fd_frm_dec calls tapped, {int(fd.sum())} pure FD; "
          f"window sequences: {np.bincount(hdr[:, 1], minlength=5).tolist()}")
    idx = np.nonzero(fd)[0]
    keep = set()
    for seq in range(5):
        cand = idx[hdr[idx, 1] == seq]
        keep.update(cand[:: max(1, len(cand) // 8)][:8].tolist())
    # a run of 24 consecutive calls (12 frames x 2 channels) that are all pure FD, for the state-carry test
    run0 = next((i for i in range(40, len(raw) - 24) if fd[i:i + 24].all()), None)
    if run0 is not None:
        keep.update(range(run0, run0 + 24))
    keep = sorted(keep)
    sel = raw[keep]
    out = dict(hdr=sel[:, 1:9].copy(), coef=sel[:, 9:1033].copy(), ov_in=sel[:, 1033:2057].copy(),
               out=sel[:, 2057:3081].copy(), ov_out=sel[:, 3081:4105].copy(), index=np.array(keep, np.int32),
               run_start=np.array([-1 if run0 is None else keep.index(run0)], np.int32))
    path = os.path.join(GOLD, "usac_fd_tapped.npz")
    np.savez_compressed(path, **out)
    print(f"wrote {path}: {len(keep)} records, {os.path.getsize(path)} bytes")


def make_esbr_golden(tmp):
    """USAC (xHE-AAC, aot 42, ccfl 1024) stereo 32 kHz with eSBR, once with the default patching and once with the harmonic
    transposer (-harmonic_sbr:1): every ixheaacd_generate_hf / ixheaacd_sbr_env_calc call of the real decodes is tapped
    (oracle/ref_taps_esbr.c) in the flat XO_EHF_* / XO_EEC_* layouts.  A selection of records within the supported subset
    is kept (plus the counts of everything seen)."""
    fs, ch, br = 32000, 2, 64000
    ehf_words = 4 + 96 + 6 + 6 + 8 + 8 + 8 * 2560
    eec_words = 3 + 288 + 288 + 464 + 640 + 640 + 4 * 2560
    ehf_sel, eec_sel, stats = [], [], []
    for tag, extra, seed in (("plain", [], 21), ("hbe", ["-harmonic_sbr:1"], 22)):
        wav = os.path.join(tmp, f"in_esbr_{tag}.wav")
        write_wav(wav, synth(fs, 6.0, ch, seed), fs)
        mp4 = os.path.join(tmp, f"esbr_{tag}.mp4")
        run([os.path.join(REFDIR, "xaacenc"), f"-ifile:{wav}", f"-ofile:{mp4}", "-aot:42", f"-br:{br}", "-ccfl_idx:3"] + extra)
        meta = os.path.join(tmp, f"esbr_{tag}.txt")
        tap = os.path.join(tmp, f"esbr_{tag}.tap")
        decode_tap(mp4, os.path.join(tmp, "o.wav"), tap, [f"-imeta:{meta}", "-mp4:1"], stages="ehf,eec")
        a = np.fromfile(tap + ".ehf", dtype=np.int32)
        b = np.fromfile(tap + ".eec", dtype=np.int32)
        assert a.size and a.size % ehf_words == 0 and b.size and b.size % eec_words == 0, (a.size, b.size)
        a, b = a.reshape(-1, ehf_words), b.reshape(-1, eec_words)
        assert (a[:, 0] == 0x31464845).all() and (b[:, 0] == 0x31434545).all()
        par = a[:, 4:100]
        ok_a = (a[:, 1] == 0) & (a[:, 3] == 0) & (par[:, 8] == 0) & (par[:, 9] == 0)
        ip = b[:, 3:291]
        ok_b = (b[:, 1] == 0) & (b[:, 2] == 0) & (ip[:, 16] == 0) & (ip[:, 17] == 1) & (ip[:, 18] == 0) & (ip[:, 19] == 0) \
            & (ip[:, 44:52] == 0).all(1)
        stats.append(f"{tag}: generate_hf {len(a)} calls ({int(ok_a.sum())} in subset; hbe_flag {int((par[:, 5] != 0).sum())}, "
                     f"patching {int((par[:, 6] != 0).sum())}, pre_proc {int((par[:, 8] != 0).sum())}, has_pv {int((a[:, 2] != 0).sum())}); "
                     f"env_calc {len(b)} calls ({int(ok_b.sum())} in subset; reset {int((ip[:, 16] != 0).sum())}, "
                     f"patching changed {int((ip[:, 19] != 0).sum())}, inter-TES {int((ip[:, 44:52] != 0).any(1).sum())}, "
                     f"num_env {np.bincount(ip[:, 2], minlength=6).tolist()})")
        ia = np.flatnonzero(ok_a)
        ib = np.flatnonzero(ok_b)
        ehf_sel.append(a[ia[len(ia) // 5:: max(1, len(ia) // 6)][:6]])
        # a run of 8 consecutive calls (4 frames x 2 channels) for the state carry + a spread of others
        run0 = next((i for i in range(20, len(b) - 8) if ok_b[i:i + 8].all()), None)
        keep = set(ib[len(ib) // 7:: max(1, len(ib) // 5)][:5].tolist())
        if run0 is not None and tag == "plain":
            keep.update(range(run0, run0 + 8))
        eec_sel.append(b[sorted(keep)])
    for s_ in stats:
        print(s_)
    a, b = np.concatenate(ehf_sel), np.concatenate(eec_sel)
    f32 = lambda x: np.ascontiguousarray(x).view(np.float32)
    q = a[:, 128:].reshape(len(a), 8, 40, 64)
    path = os.path.join(GOLD, "esbr_hfgen_tapped.npz")
    np.savez_compressed(path, ret=a[:, 1].copy(), has_pv=a[:, 2].copy(), par=a[:, 4:100].copy(), bw_in=f32(a[:, 100:106]),
                        bw_out=f32(a[:, 106:112]), patch=a[:, 112:120].copy(), patch_in=a[:, 120:128].copy(), src_re=f32(q[:, 0]), src_im=f32(q[:, 1]),
                        pv_re=f32(q[:, 2]), pv_im=f32(q[:, 3]), dst_in_re=f32(q[:, 4]), dst_in_im=f32(q[:, 5]),
                        dst_out_re=f32(q[:, 6]), dst_out_im=f32(q[:, 7]), stats=np.array(stats))
    print(f"wrote {path}: {len(a)} records, {os.path.getsize(path)} bytes")
    o = 3
    ipi, ipo = b[:, o:o + 288], b[:, o + 288:o + 576]
    fp_ = f32(b[:, o + 576:o + 1040])
    sti, sto = f32(b[:, o + 1040:o + 1680]), f32(b[:, o + 1680:o + 2320])
    q = b[:, o + 2320:].reshape(len(b), 4, 40, 64)
    path = os.path.join(GOLD, "esbr_envcalc_tapped.npz")
    np.savez_compressed(path, ret=b[:, 1].copy(), ipar_in=ipi.copy(), ipar_out=ipo.copy(), fpar=fp_, state_in=sti, state_out=sto,
                        re_in=f32(q[:, 0]), im_in=f32(q[:, 1]), re_out=f32(q[:, 2]), im_out=f32(q[:, 3]), stats=np.array(stats))
    print(f"wrote {path}: {len(b)} records, {os.path.getsize(path)} bytes")


def make_esbr_stage_golden(tmp):
    """Whole float eSBR stage (ixheaacd_sbr_dec, eSBR branch) of a real USAC stereo decode: 8 consecutive frames of both
    channels, tapped around the unmodified stage call (oracle/ref_taps_esbr.c: esbr_stage_tap_pre / _post)."""
    fs, ch, br = 32000, 2, 64000
    wav = os.path.join(tmp, "in_esd.wav")
    write_wav(wav, synth(fs, 6.0, ch, 21), fs)
    mp4 = os.path.join(tmp, "esd.mp4")
    run([os.path.join(REFDIR, "xaacenc"), f"-ifile:{wav}", f"-ofile:{mp4}", "-aot:42", f"-br:{br}", "-ccfl_idx:3"])
    tap = os.path.join(tmp, "esd.tap")
    decode_tap(mp4, os.path.join(tmp, "o.wav"), tap, [f"-imeta:{os.path.join(tmp, 'esd.txt')}", "-mp4:1"], stages="esd")
    S = 4 * 2560 + 320 + 2 + 1280 + 2 + 6 + 8 + 640
    W = 16 + 1024 + S + 96 + 288 + 464 + 2048 + S + 288
    raw = np.fromfile(tap + ".esd", dtype=np.int32)
    assert raw.size and raw.size % W == 0, raw.size
    r = raw.reshape(-1, W)
    assert (r[:, 0] == 0x31445345).all()
    h = r[:, 1:16]
    ok = (h[:, 0] == 0) & (h[:, 1] == 1) & (h[:, 2] == 0) & (h[:, 3] == 0) & (h[:, 4] <= 0) & (h[:, 5] == 0) & (h[:, 6] == 1) \
        & (h[:, 13] == 1) & (h[:, 14] == 1)
    print(f"eSBR stage: {len(r)} calls tapped, {int(ok.sum())} in the supported subset; qmf_sb_prev {sorted(set(h[:, 7].tolist()))}, "
          f"sub_band_start {sorted(set(h[:, 8].tolist()))}, border0 {sorted(set(h[:, 9].tolist()))}")
    n_rec = 16
    r0 = next(i for i in range(20, len(r) - n_rec) if ok[i:i + n_rec].all() and h[i, 12] == 0)
    sel = r[r0:r0 + n_rec]
    f32 = lambda x: np.ascontiguousarray(x).view(np.float32)
    o = 16
    tin = f32(sel[:, o:o + 1024]); o += 1024
    st_in = sel[:, o:o + S]; o += S
    hf_par = sel[:, o:o + 96]; o += 96
    ipar_in = sel[:, o:o + 288]; o += 288
    fpar = f32(sel[:, o:o + 464]); o += 464
    tout = f32(sel[:, o:o + 2048]); o += 2048
    st_out = sel[:, o:o + S]; o += S
    ipar_out = sel[:, o:o + 288]

    def split(st):
        q = f32(st[:, :10240]).reshape(-1, 4, 40, 64)
        p = 10240
        d = dict(qmf_re=q[:, 0], qmf_im=q[:, 1], out_re=q[:, 2], out_im=q[:, 3])
        for k, n_, fl in (("anal_states", 320, 0), ("anal_pos", 2, 0), ("synth_states", 1280, 0), ("synth_pos", 2, 0),
                          ("bw_prev", 6, 1), ("patch", 8, 0), ("ec_state", 640, 1)):
            d[k] = f32(st[:, p:p + n_]) if fl else st[:, p:p + n_].copy()
            p += n_
        return d
    si, so = split(st_in), split(st_out)
    out = dict(head=sel[:, 1:16].copy(), time_in=tin, hf_par=hf_par.copy(), ec_ipar_in=ipar_in.copy(), ec_fpar=fpar,
               time_out=tout, ec_ipar_out=ipar_out.copy())
    for k, v in si.items():     # full state before the first frame of each channel
        out["in0_" + k] = v[:2].copy()
    for k, v in so.items():     # small states after every frame, the QMF arrays after the last frame of each channel
        out["out_" + k] = v[-2:].copy() if k in ("qmf_re", "qmf_im", "out_re", "out_im") else v.copy()
    path = os.path.join(GOLD, "esbr_stage_tapped.npz")
    np.savez_compressed(path, **out)
    print(f"wrote {path}: {n_rec} records, {os.path.getsize(path)} bytes")


HBE_WORDS = 2 + 16 + 3616 + 4 * 2048 + 3616


def read_hbe_records(path):
    a = np.fromfile(path, dtype=np.int32)
    assert a.size and a.size % HBE_WORDS == 0, a.size
    a = a.reshape(-1, HBE_WORDS)
    assert (a[:, 0] == 0x31454248).all()
    return a


def make_hbe_golden(tmp):
    """QMF harmonic transposer ixheaacd_qmf_hbe_apply of real USAC decodes (-harmonic_sbr:1, stereo 32 kHz): every call is
    tapped (oracle/ref_shim_hbe.c) in the flat XO_HBE_* layout.  Kept: 8 consecutive calls (4 frames x 2 channels) for the
    state carry plus a spread over the pitch values the encoder signals."""
    fs, ch, br = 32000, 2, 64000
    wav = os.path.join(tmp, "in_hbe.wav")
    write_wav(wav, synth(fs, 6.0, ch, 22), fs)
    mp4 = os.path.join(tmp, "hbe.mp4")
    run([os.path.join(REFDIR, "xaacenc"), f"-ifile:{wav}", f"-ofile:{mp4}", "-aot:42", f"-br:{br}", "-ccfl_idx:3", "-harmonic_sbr:1"])
    tap = os.path.join(tmp, "hbe.tap")
    decode_tap(mp4, os.path.join(tmp, "o.wav"), tap, [f"-imeta:{os.path.join(tmp, 'hbe.txt')}", "-mp4:1"], stages="hbe")
    a = read_hbe_records(tap + ".hbe")
    cfg = a[:, 2:18]
    ok = a[:, 1] == 0
    pitches = sorted(set(cfg[:, 5].tolist()))
    stats = (f"qmf_hbe_apply: {len(a)} calls, {int(ok.sum())} ok; synth_size {sorted(set(cfg[:, 0].tolist()))}, k_start "
             f"{sorted(set(cfg[:, 1].tolist()))}, bands {sorted(set(map(tuple, cfg[:, 2:4].tolist())))}, max_stretch "
             f"{sorted(set(cfg[:, 4].tolist()))}, pitch_in_bins {pitches}, reinit inside the call {int(cfg[:, 7].sum())}")
    print(stats)
    run0 = next(i for i in range(40, len(a) - 8) if ok[i:i + 8].all())
    keep = set(range(run0, run0 + 8))
    for pv in pitches:
        idx = np.flatnonzero(ok & (cfg[:, 5] == pv))
        keep.update(idx[len(idx) // 2: len(idx) // 2 + 2].tolist())
    sel = a[sorted(keep)]
    f32 = lambda x: np.ascontiguousarray(x).view(np.float32)
    o = 18
    st_in = f32(sel[:, o:o + 3616]); o += 3616
    q = f32(sel[:, o:o + 4 * 2048]).reshape(len(sel), 4, 32, 64); o += 4 * 2048
    st_out = f32(sel[:, o:o + 3616])
    path = os.path.join(GOLD, "esbr_hbe_tapped.npz")
    np.savez_compressed(path, ret=sel[:, 1].copy(), cfg=sel[:, 2:18].copy(), state_in=st_in, state_out=st_out, qmf_re=q[:, 0],
                        qmf_im=q[:, 1], pv_re=q[:, 2], pv_im=q[:, 3], run=np.array([sorted(keep).index(run0), 8]),
                        stats=np.array([stats]))
    print(f"wrote {path}: {len(sel)} records, {os.path.getsize(path)} bytes")


def make_esbr_hbe_stage_golden(tmp):
    """Whole float eSBR stage WITH the harmonic transposer (hbe_flag = 1) of a real USAC stereo decode (-harmonic_sbr:1):
    6 consecutive frames of both channels tapped around the unmodified ixheaacd_sbr_dec (oracle/ref_taps_esbr.c, '.esh')."""
    fs, ch, br = 32000, 2, 64000
    wav = os.path.join(tmp, "in_esh.wav")
    write_wav(wav, synth(fs, 6.0, ch, 22), fs)
    mp4 = os.path.join(tmp, "esh.mp4")
    run([os.path.join(REFDIR, "xaacenc"), f"-ifile:{wav}", f"-ofile:{mp4}", "-aot:42", f"-br:{br}", "-ccfl_idx:3", "-harmonic_sbr:1"])
    tap = os.path.join(tmp, "esh.tap")
    decode_tap(mp4, os.path.join(tmp, "o.wav"), tap, [f"-imeta:{os.path.join(tmp, 'esh.txt')}", "-mp4:1"], stages="esh")
    S = 2 * 4608 + 4 * 2560 + 320 + 2 + 1280 + 2 + 6 + 8 + 640 + 3616
    W = 16 + 1024 + S + 96 + 288 + 464 + 16 + 2048 + S + 288
    raw = np.fromfile(tap + ".esh", dtype=np.int32)
    assert raw.size and raw.size % W == 0, raw.size
    r = raw.reshape(-1, W)
    assert (r[:, 0] == 0x31485345).all()
    h = r[:, 1:16]
    ok = (h[:, 0] == 0) & (h[:, 1] == 1) & (h[:, 2] == 1) & (h[:, 3] == 0) & (h[:, 4] <= 0) & (h[:, 5] == 0) & (h[:, 6] == 1) \
        & (h[:, 13] == 1) & (h[:, 14] == 1)
    o = 16 + 1024 + S
    hfp = r[:, o:o + 96]
    ipi = r[:, o + 96:o + 96 + 288]
    # a change of sbr_patching_mode / a reset frame makes ixheaacd_sbr_env_calc rebuild its limiter tables (control plane: the
    # host recomputes them before the call); the tap records the tables as they were BEFORE the call
    ok &= (ipi[:, 16] == 0) & (ipi[:, 19] == 0)
    cfg = r[:, o + 96 + 288 + 464:o + 96 + 288 + 464 + 16]
    print(f"eSBR+HBE stage: {len(r)} calls tapped, {int(ok.sum())} in the supported subset; patching mode {sorted(set(hfp[:, 6].tolist()))}, "
          f"pitch {sorted(set(cfg[:, 5].tolist()))}, qmf_sb_prev {sorted(set(h[:, 7].tolist()))}, sub_band_start {sorted(set(h[:, 8].tolist()))}")
    n_rec = 12
    # a run that contains both plain (pitch 0) and cross-product frames if there is one
    cands = [i for i in range(20, len(r) - n_rec) if ok[i:i + n_rec].all() and h[i, 12] == 0]
    r0 = next((i for i in cands if len(set(cfg[i:i + n_rec, 5].tolist())) > 1), cands[0])
    sel = r[r0:r0 + n_rec]
    f32 = lambda x: np.ascontiguousarray(x).view(np.float32)
    o = 16
    tin = f32(sel[:, o:o + 1024]); o += 1024
    st_in = sel[:, o:o + S]; o += S
    hf_par = sel[:, o:o + 96]; o += 96
    ipar_in = sel[:, o:o + 288]; o += 288
    fpar = f32(sel[:, o:o + 464]); o += 464
    hcfg = sel[:, o:o + 16]; o += 16
    tout = f32(sel[:, o:o + 2048]); o += 2048
    st_out = sel[:, o:o + S]; o += S
    ipar_out = sel[:, o:o + 288]

    def split(st):
        d = {}
        p = 0
        for k, shp, fl in (("qmf_re", (72, 64), 1), ("qmf_im", (72, 64), 1), ("out_re", (40, 64), 1), ("out_im", (40, 64), 1),
                           ("pv_re", (40, 64), 1), ("pv_im", (40, 64), 1), ("anal_states", (320,), 0), ("anal_pos", (2,), 0),
                           ("synth_states", (1280,), 0), ("synth_pos", (2,), 0), ("bw_prev", (6,), 1), ("patch", (8,), 0),
                           ("ec_state", (640,), 1), ("hbe_state", (3616,), 1)):
            n_ = int(np.prod(shp))
            a = st[:, p:p + n_]
            d[k] = (f32(a) if fl else a.copy()).reshape((len(st),) + shp)
            p += n_
        assert p == S
        return d
    si, so = split(st_in), split(st_out)
    out = dict(head=sel[:, 1:16].copy(), time_in=tin, hf_par=hf_par.copy(), ec_ipar_in=ipar_in.copy(), ec_fpar=fpar,
               hbe_cfg=hcfg.copy(), time_out=tout, ec_ipar_out=ipar_out.copy())
    big = ("qmf_re", "qmf_im", "out_re", "out_im", "pv_re", "pv_im")
    for k, v in si.items():
        out["in0_" + k] = v[:2].copy()
    for k, v in so.items():
        out["out_" + k] = v[-2:].copy() if k in big else v.copy()
    path = os.path.join(GOLD, "esbr_hbe_stage_tapped.npz")
    np.savez_compressed(path, **out)
    print(f"wrote {path}: {n_rec} records, {os.path.getsize(path)} bytes")


def make_fps_golden(tmp):
    """Float parametric stereo ixheaacd_esbr_apply_ps: outputs of the COMPILED reference function (oracle/ref_shim_fps.c) for 4
    seeded channels over 3 consecutive frames (state carried), inputs quantised to int16 x 2^k so they store compactly.  Also
    holds the side records / smoothing history the drop-in's host code derives (checked against the reference by
    tests/test_ps_flt_host.py where oracle/_ref is present)."""
    sys.path.insert(0, ROOT)
    from tests import oracle_util as ou
    ref = ou.Ref.try_load()
    rng = np.random.default_rng(20261017)
    n, frames = 4, 3
    st, hst = ou.fps_fresh_state(n)
    rec = {k: [] for k in ("low_re_i16", "low_im_i16", "par", "side", "left", "right", "state_out", "hst_out")}
    shift = np.array([-4, 0, -8, -2], np.int32)
    for f in range(frames):
        lr, li, par = ou.synth_fps_frame(n, rng)
        qr = np.clip(np.round(lr / np.abs(lr).max((1, 2), keepdims=True) * 30000), -32768, 32767).astype(np.int16)
        qi = np.clip(np.round(li / np.abs(li).max((1, 2), keepdims=True) * 30000), -32768, 32767).astype(np.int16)
        lr = np.ldexp(qr.astype(np.float32), shift[:, None, None]).astype(np.float32)
        li = np.ldexp(qi.astype(np.float32), shift[:, None, None]).astype(np.float32)
        r = ou.ref_fps_batch(ref, lr, li, par, st, hst)
        assert r["rc"] == 0
        for k, v in (("low_re_i16", qr), ("low_im_i16", qi), ("par", par), ("side", r["side"]), ("left", r["left"]),
                     ("right", r["right"]), ("state_out", r["state"]), ("hst_out", r["hst"])):
            rec[k].append(v)
        st, hst = r["state"], r["hst"]
    path = os.path.join(GOLD, "esbr_ps_ref.npz")
    np.savez_compressed(path, shift=shift, **{k: np.stack(v) for k, v in rec.items()})
    print(f"wrote {path}: {frames} frames x {n} channels, {os.path.getsize(path)} bytes")


def make_sps_golden(tmp):
    """AAC pre-IMDCT spectral stage ixheaacd_channel_pair_process (AAC-LC): outputs of the COMPILED reference function
    (oracle/ref_shim_sps.c) for 20 seeded elements covering M/S, intensity, PNS (with generator state) and TNS."""
    sys.path.insert(0, ROOT)
    from tests import oracle_util as ou
    ref = ou.Ref.try_load()
    spec, rec = ou.synth_sps_units(20, 20261018, pns=True)
    seed = (np.arange(20, dtype=np.int64) * 2654435761 % (1 << 32)).astype(np.uint32).view(np.int32)
    out, err, seed_out = ou.ref_channel_pair_process(ref, spec, rec, seed)
    assert (err == 0).all()
    path = os.path.join(GOLD, "aac_spectral_ref.npz")
    np.savez_compressed(path, spec_in=spec, rec=rec, seed_in=seed, spec_out=out, seed_out=seed_out)
    print(f"wrote {path}: 20 elements, {os.path.getsize(path)} bytes; changed cells {int((out != spec).sum())}, "
          f"seeds advanced {int((seed_out != seed).sum())}")


def main():
    os.makedirs(GOLD, exist_ok=True)
    which = sys.argv[1:] or ["imdct", "hfgen", "envcalc", "sbrdec", "sbrdec_lp", "usac_fd", "esbr", "esbr_stage", "hbe", "hbe_stage", "fps", "sps"]
    with tempfile.TemporaryDirectory() as tmp:
        if "imdct" in which:
            make_imdct_golden(tmp)
        if "hfgen" in which:
            make_hfgen_golden(tmp)
        if "envcalc" in which:
            make_envcalc_golden(tmp)
        if "sbrdec" in which:
            make_sbrdec_golden(tmp)
        if "usac_fd" in which:
            make_usac_fd_golden(tmp)
        if "esbr" in which:
            make_esbr_golden(tmp)
        if "esbr_stage" in which:
            make_esbr_stage_golden(tmp)
        if "sbrdec_lp" in which:
            make_sbrdec_lp_golden(tmp)
        if "hbe" in which:
            make_hbe_golden(tmp)
        if "hbe_stage" in which:
            make_esbr_hbe_stage_golden(tmp)
        if "fps" in which:
            make_fps_golden(tmp)
        if "sps" in which:
            make_sps_golden(tmp)


if __name__ == "__main__":
    main()
