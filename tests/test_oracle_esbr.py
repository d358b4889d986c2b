"""CPU tests: the eSBR 64-band synthesis bank (first piece of SURVEY.md 8a-E) — our C restatement against the compiled
reference's own leaf functions driven in the order of ixheaacd_esbr_synthesis_filt_block.  Integer inside, so the float
output is compared bit for bit."""
import numpy as np

from tests import oracle_util


def test_esbr_synth_matches_reference(oracle, ref):
    n = 200
    qmf, fs, pos = oracle_util.synth_esbr_units(n, 5)
    o1, f1, p1 = oracle.esbr_synth_batch(qmf, fs, pos)
    o2, f2, p2 = ref.esbr_synth_batch(qmf, fs, pos)
    assert np.array_equal(p1, p2)
    for u in range(n):
        assert np.array_equal(f1[u], f2[u]), f"unit {u}: state differs at {np.argwhere(f1[u] != f2[u]).ravel()[:8]}"
        assert np.array_equal(o1[u].view(np.int32), o2[u].view(np.int32)), f"unit {u}: output differs"
    assert np.abs(o2).max() > 100


def test_esbr_synth_stream_state_carry(oracle, ref):
    n, frames = 24, 6
    _, fs1, pos1 = oracle_util.synth_esbr_units(n, 9)
    fs1[:] = 0
    pos1[:] = 0
    fs2, pos2 = fs1.copy(), pos1.copy()
    for f in range(frames):
        qmf, _, _ = oracle_util.synth_esbr_units(n, 20 + f)
        o1, fs1, pos1 = oracle.esbr_synth_batch(qmf, fs1, pos1)
        o2, fs2, pos2 = ref.esbr_synth_batch(qmf, fs2, pos2)
        assert np.array_equal(o1.view(np.int32), o2.view(np.int32)) and np.array_equal(fs1, fs2) and np.array_equal(pos1, pos2)
    assert set(map(tuple, pos1.tolist())) <= {(0, 0), (256, 512), (512, 384), (768, 256), (1024, 128)}


def test_esbr_anal_matches_reference(oracle, ref):
    """ixheaacd_esbr_analysis_filt_block itself (32 channels, 32 slots) on a minimal ia_sbr_dec_struct"""
    n = 200
    x, st, pos = oracle_util.synth_esbr_anal_units(n, 3)
    q1, s1, p1 = oracle.esbr_anal_batch(x, st, pos)
    q2, s2, p2 = ref.esbr_anal_batch(x, st, pos)
    assert np.array_equal(p1, p2) and np.array_equal(s1, s2)
    for u in range(n):
        assert np.array_equal(q1[u].view(np.int32), q2[u].view(np.int32)), f"unit {u}: {np.argwhere(q1[u] != q2[u])[:4].tolist()}"
    assert np.abs(q2).max() > 0.01


def test_esbr_anal_synth_stream(oracle, ref):
    """analysis -> synthesis over 6 frames with both states carried (the low band passes straight through)"""
    n = 8
    x0, st, pos = oracle_util.synth_esbr_anal_units(n, 4)
    st[:] = 0
    pos[:] = 0
    st2, pos2 = st.copy(), pos.copy()
    fs = np.zeros((n, 1280), np.int32)
    sp = np.zeros((n, 2), np.int32)
    fs2, sp2 = fs.copy(), sp.copy()
    for f in range(6):
        x, _, _ = oracle_util.synth_esbr_anal_units(n, 30 + f)
        q1, st, pos = oracle.esbr_anal_batch(x, st, pos)
        q2, st2, pos2 = ref.esbr_anal_batch(x, st2, pos2)
        assert np.array_equal(q1.view(np.int32), q2.view(np.int32)), f"frame {f}"
        o1, fs, sp = oracle.esbr_synth_batch(q1, fs, sp)
        o2, fs2, sp2 = ref.esbr_synth_batch(q2, fs2, sp2)
        assert np.array_equal(o1.view(np.int32), o2.view(np.int32)), f"frame {f}"
    assert np.array_equal(st, st2) and np.array_equal(fs, fs2)


def _same_hfgen(a, b, what):
    dr1, di1, bw1, pt1, e1 = a
    dr2, di2, bw2, pt2, e2 = b
    assert np.array_equal(e1, e2), f"{what}: err differs at {np.argwhere(e1 != e2).ravel()[:8]}: {e1[e1 != e2][:8]} vs {e2[e1 != e2][:8]}"
    ok = e2 == 0
    for u in np.flatnonzero(ok):
        assert np.array_equal(pt1[u], pt2[u]), f"{what}: unit {u} patches {pt1[u]} vs {pt2[u]}"
        assert np.array_equal(bw1[u].view(np.int32), bw2[u].view(np.int32)), f"{what}: unit {u} bw"
        for x, y, nm in ((dr1, dr2, "re"), (di1, di2, "im")):
            if not np.array_equal(x[u].view(np.int32), y[u].view(np.int32)):
                w = np.argwhere(x[u].view(np.int32) != y[u].view(np.int32))
                raise AssertionError(f"{what}: unit {u} {nm} differs at {w[:6].tolist()} ({len(w)} cells)")
    return int(ok.sum())


def test_esbr_generate_hf_matches_reference(oracle, ref):
    """ixheaacd_generate_hf itself (2:1 system, no pre-processing) — float results bit for bit, incl. untouched cells,
    patch table, chirp-factor state and the -1 returns"""
    d = oracle_util.synth_esbr_hfgen_units(400, 11)
    a = oracle_util.oracle_esbr_hfgen_batch(oracle, d)
    b = oracle_util.ref_esbr_hfgen_batch(ref, d)
    good = _same_hfgen(a, b, "hbe buffers")
    assert good > 300 and (b[4] == -1).sum() > 10
    assert np.abs(b[0] - d["dst_re"]).max() > 0
    d2 = oracle_util.synth_esbr_hfgen_units(200, 12, hbe=False)
    _same_hfgen(oracle_util.oracle_esbr_hfgen_batch(oracle, d2, with_pv=False),
                oracle_util.ref_esbr_hfgen_batch(ref, d2, with_pv=False), "no hbe")


def same_envcalc(a, b, what):
    re1, im1, ip1, st1, e1 = a
    re2, im2, ip2, st2, e2 = b
    assert np.array_equal(e1, e2), f"{what}: err differs at {np.argwhere(e1 != e2).ravel()[:8]}: {e1[e1 != e2][:8]} vs {e2[e1 != e2][:8]}"
    ok = np.flatnonzero(e2 == 0)
    for u in ok:
        assert np.array_equal(ip1[u], ip2[u]), f"{what}: unit {u} ipar words {np.argwhere(ip1[u] != ip2[u]).ravel()[:8]}"
        for x, y, nm in ((re1, re2, "re"), (im1, im2, "im"), (st1, st2, "state")):
            if not np.array_equal(x[u].view(np.int32), y[u].view(np.int32)):
                w = np.argwhere(x[u].view(np.int32) != y[u].view(np.int32))
                raise AssertionError(f"{what}: unit {u} {nm} differs at {w[:6].tolist()} ({len(w)} cells)")
    return len(ok)


def test_esbr_env_calc_matches_reference(oracle, ref):
    """ixheaacd_sbr_env_calc itself (ORIG_SBR, 2:1, no reset) — adjusted QMF cells, smoothing history, harmonic flags, phase
    and harmonic indices bit for bit"""
    rp = oracle_util.esbr_random_phase(ref)
    d = oracle_util.synth_esbr_envcalc_units(400, 17)
    b = oracle_util.ref_esbr_envcalc_batch(ref, d)
    good = same_envcalc(oracle_util.oracle_esbr_envcalc_batch(oracle, d, rp), b, "env calc")
    assert good > 350 and (b[4] != 0).sum() >= 12
    assert np.abs(b[0] - d["re"]).max() > 0


def load_esbr_golden(name):
    import os
    return np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", name))


def test_esbr_generate_hf_pre_processing_matches_reference(oracle, ref):
    """pre_proc_flag = 1: ixheaacd_pre_processing (decoder/ixheaacd_sbrdec_lpfuncs.c:928-979 — low-band level in dB, third-order
    ixheaacd_polyfit with ixheaacd_gausssolve, gain = 10^((mean - slope) / 20)) scales every patch by the gain of its source band.
    The restatement calls the same libm, so the floats are identical to the compiled function's."""
    for hbe in (True, False):
        d = oracle_util.synth_esbr_hfgen_units(300, 31 + hbe, hbe=hbe)
        d["par"][:, oracle_util.EHF["PRE_PROC"]] = 1
        if hbe:
            d["par"][::2, oracle_util.EHF["PATCHING_MODE"]] = 1  # the gains only act in the LPP patch branch
        a = oracle_util.oracle_esbr_hfgen_batch(oracle, d, with_pv=hbe)
        b = oracle_util.ref_esbr_hfgen_batch(ref, d, with_pv=hbe)
        good = _same_hfgen(a, b, f"pre-processing, hbe={hbe}")
        d0 = dict(d)
        d0["par"] = d["par"].copy()
        d0["par"][:, oracle_util.EHF["PRE_PROC"]] = 0
        a0 = oracle_util.oracle_esbr_hfgen_batch(oracle, d0, with_pv=hbe)
        assert not np.array_equal(a[0].view(np.int32), a0[0].view(np.int32)), "the gains changed nothing"


def golden_hfgen_units(g):
    return dict(par=g["par"], src_re=g["src_re"], src_im=g["src_im"], pv_re=g["pv_re"], pv_im=g["pv_im"],
                dst_re=g["dst_in_re"], dst_im=g["dst_in_im"], bw_prev=g["bw_in"], patch_in=g["patch_in"])


def check_hfgen_golden(out, g, what):
    dr, di, bw, patch, err = out
    assert np.array_equal(err, g["ret"]), f"{what}: return values"
    assert np.array_equal(patch, g["patch"]), f"{what}: patch table"
    assert np.array_equal(bw.view(np.int32), g["bw_out"].view(np.int32)), f"{what}: bw_array_prev"
    for x, y, nm in ((dr, g["dst_out_re"], "re"), (di, g["dst_out_im"], "im")):
        bad = np.argwhere(x.view(np.int32) != y.view(np.int32))
        assert len(bad) == 0, f"{what}: {nm} differs in {len(bad)} cells, first {bad[0].tolist()}"


def golden_envcalc_units(g):
    return dict(re=g["re_in"], im=g["im_in"], ipar=g["ipar_in"], fpar=g["fpar"], state=g["state_in"])


def check_envcalc_golden(out, g, what):
    re, im, ipar, state, err = out
    assert np.array_equal(err, g["ret"]), f"{what}: return values"
    assert np.array_equal(ipar, g["ipar_out"]), f"{what}: ipar words {np.argwhere(ipar != g['ipar_out'])[:6].tolist()}"
    for x, y, nm in ((re, g["re_out"], "re"), (im, g["im_out"], "im"), (state, g["state_out"], "state")):
        bad = np.argwhere(x.view(np.int32) != y.view(np.int32))
        assert len(bad) == 0, f"{what}: {nm} differs in {len(bad)} cells, first {bad[0].tolist()}"


def test_esbr_generate_hf_golden(oracle):
    """records tapped from real USAC decodes (default patching and harmonic transposer), tools/make_golden.py esbr"""
    g = load_esbr_golden("esbr_hfgen_tapped.npz")
    assert (g["has_pv"] != 0).sum() >= 4 and (g["has_pv"] == 0).sum() >= 4
    check_hfgen_golden(oracle_util.oracle_esbr_hfgen_batch(oracle, golden_hfgen_units(g)), g, "oracle vs tapped decode")
    assert np.abs(g["dst_out_re"] - g["dst_in_re"]).max() > 0


def test_esbr_env_calc_golden(oracle):
    g = load_esbr_golden("esbr_envcalc_tapped.npz")
    rp = oracle_util.esbr_random_phase()
    check_envcalc_golden(oracle_util.oracle_esbr_envcalc_batch(oracle, golden_envcalc_units(g), rp), g, "oracle vs tapped decode")
    assert len(set(g["ipar_in"][:, oracle_util.EEC["NUM_ENV"]].tolist())) >= 3


def test_samples_sat_matches_reference(oracle, ref):
    """float -> PCM16 hand-over (clamp, truncation towards zero, interleave) against ixheaacd_samples_sat itself"""
    rng = np.random.default_rng(3)
    n, nch = 2048, 2
    x = (rng.standard_normal((nch, n)) * 20000).astype(np.float32)
    x[0, :8] = [32767.0, 32767.5, 32768.0, -32768.0, -32768.5, -32769.0, 0.99, -0.99]
    x[1, :4] = [1e9, -1e9, 0.0, -0.0]
    want = np.zeros((n, nch), np.int16)
    ref.lib.ref_samples_sat16(oracle_util.P(x), nch, n, oracle_util.P(want))
    got = np.zeros((n, nch), np.int16)
    for c in range(nch):
        oracle.lib.xo_samples_sat16(oracle_util.P(np.ascontiguousarray(x[c])), nch, c, oracle_util.P(got), n)
    assert np.array_equal(got, want)
    assert list(want[:8, 0]) == [32767, 32767, 32767, -32768, -32768, -32768, 0, 0]


def esbr_stage_golden_frames(g):
    """yields per frame (records 2f, 2f+1 = channels 0, 1) the inputs and expected outputs of the tapped stage calls"""
    head = g["head"]
    assert (head[0::2, 12] == 0).all() and (head[1::2, 12] == 1).all()
    for f in range(len(head) // 2):
        r = slice(2 * f, 2 * f + 2)
        rg = np.stack([head[r, 7], head[r, 8], 2 * head[r, 9], 0 * head[r, 9]], 1).astype(np.int32)
        yield f, r, rg


def test_esbr_stage_golden(oracle):
    """the composed oracle stage against 8 consecutive frames x 2 channels tapped around ixheaacd_sbr_dec in a real USAC
    decode: time output, both bank states, chirp factors, patch table, smoothing history, in/out parameter words per
    frame, the QMF history arrays after the last frame"""
    g = load_esbr_golden("esbr_stage_tapped.npz")
    rp = oracle_util.esbr_random_phase()
    st = {k: g["in0_" + k] for k in oracle_util.ESD_KEYS}
    for f, r, rg in esbr_stage_golden_frames(g):
        out, st, ipar2, err = oracle_util.oracle_esbr_stage(oracle, rp, st, g["time_in"][r], g["hf_par"][r], g["ec_ipar_in"][r],
                                                            g["ec_fpar"][r], rg)
        assert not err.any(), f"frame {f}: {err}"
        assert np.array_equal(out.view(np.int32), g["time_out"][r].view(np.int32)), f"frame {f}: time output"
        assert np.array_equal(ipar2, g["ec_ipar_out"][r]), f"frame {f}: in/out parameter words"
        for k in ("anal_states", "anal_pos", "synth_states", "synth_pos", "bw_prev", "patch", "ec_state"):
            assert np.array_equal(st[k].view(np.int32), g["out_" + k][r].view(np.int32)), f"frame {f}: {k}"
    for k in ("qmf_re", "qmf_im", "out_re", "out_im"):
        assert np.array_equal(st[k].view(np.int32), g["out_" + k].view(np.int32)), k
    assert np.abs(g["time_out"]).max() > 100
