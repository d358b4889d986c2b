"""CPU tests: the QMF harmonic transposer (ixheaacd_qmf_hbe_apply) — our C restatement against records tapped
from a real USAC decode (-harmonic_sbr:1) and against the compiled reference function driven through the shim over all bank
sizes (FFT banks 4 / 8 / 12 / 16 and the direct-form size 20), stretch orders 2 / 3 / 4 and the pitch-driven cross products.
Float results are compared as bit patterns."""
import os

import numpy as np
import pytest

from tests import oracle_util

GOLD = os.path.join(os.path.dirname(__file__), "golden", "esbr_hbe_tapped.npz")


def bits(a):
    return np.ascontiguousarray(a, np.float32).view(np.int32)


def band_mask(cfg):
    m = np.zeros((len(cfg), 1, 64), bool)
    for u, c in enumerate(cfg):
        m[u, 0, c[2]:c[3]] = True
    return np.broadcast_to(m, (len(cfg), 32, 64))


def test_hbe_oracle_reproduces_tapped_records(oracle):
    g = np.load(GOLD)
    assert (g["ret"] == 0).all()
    pr, pi, st, err = oracle_util.oracle_hbe_batch(oracle, g["cfg"], g["state_in"], g["qmf_re"], g["qmf_im"], g["pv_re"], g["pv_im"])
    assert (err == 0).all()
    m = band_mask(g["cfg"])
    assert np.array_equal(bits(pr)[m], bits(g["pv_re"])[m]) and np.array_equal(bits(pi)[m], bits(g["pv_im"])[m])
    assert np.array_equal(bits(st), bits(g["state_out"]))
    assert np.abs(g["pv_re"][m]).max() > 1.0
    assert set(g["cfg"][:, 5].tolist()) >= {0, 24}  # plain and cross-product frames


def test_hbe_oracle_state_carry_over_consecutive_tapped_frames(oracle):
    g = np.load(GOLD)
    r0, cnt = g["run"]
    for ch in range(2):
        idx = list(range(r0 + ch, r0 + cnt, 2))
        st = g["state_in"][idx[0]:idx[0] + 1]
        for i in idx:
            pr, pi, st, err = oracle_util.oracle_hbe_batch(oracle, g["cfg"][i:i + 1], st, g["qmf_re"][i:i + 1], g["qmf_im"][i:i + 1])
            assert err[0] == 0
            m = band_mask(g["cfg"][i:i + 1])
            assert np.array_equal(bits(pr)[m], bits(g["pv_re"][i:i + 1])[m])
            assert np.array_equal(bits(st), bits(g["state_out"][i:i + 1]))


@pytest.mark.parametrize("mode", ["zero", "pitch", "mixed"])
def test_hbe_oracle_matches_compiled_reference(oracle, ref, mode):
    n = 120
    cfg, tbl, state, qre, qim = oracle_util.synth_hbe_units(n, {"zero": 1, "pitch": 2, "mixed": 3}[mode], ref, mode)
    assert set(cfg[:, 0].tolist()) == {4, 8, 12, 16, 20} and set(cfg[:, 4].tolist()) >= {2, 3, 4}
    p1, i1, s1, e1 = oracle_util.oracle_hbe_batch(oracle, cfg, state, qre, qim)
    p2, i2, s2, e2 = oracle_util.ref_hbe_batch(ref, cfg, state, qre, qim, tbl)
    assert np.array_equal(e1, e2) and (e1 == 0).all(), (e1[:10], e2[:10])
    for u in range(n):
        b0, b1 = cfg[u, 2], cfg[u, 3]
        assert np.array_equal(bits(p1[u][:, b0:b1]), bits(p2[u][:, b0:b1])), f"unit {u} cfg {cfg[u].tolist()}: pv_re differs"
        assert np.array_equal(bits(i1[u][:, b0:b1]), bits(i2[u][:, b0:b1])), f"unit {u}: pv_im differs"
        assert np.array_equal(bits(s1[u]), bits(s2[u])), f"unit {u} cfg {cfg[u].tolist()}: state differs at {np.flatnonzero(bits(s1[u]) != bits(s2[u]))[:8]}"


def test_hbe_oracle_streams_match_reference(oracle, ref):
    """state carried over 5 frames by both implementations"""
    n = 30
    cfg, tbl, state, _, _ = oracle_util.synth_hbe_units(n, 11, ref, "mixed")
    s1, s2 = state.copy(), state.copy()
    for f in range(5):
        _, _, _, qre, qim = oracle_util.synth_hbe_units(n, 100 + f, ref, "zero")
        p1, i1, s1, e1 = oracle_util.oracle_hbe_batch(oracle, cfg, s1, qre, qim)
        p2, i2, s2, e2 = oracle_util.ref_hbe_batch(ref, cfg, s2, qre, qim, tbl)
        assert (e1 == 0).all() and (e2 == 0).all()
        m = band_mask(cfg)
        assert np.array_equal(bits(p1)[m], bits(p2)[m]) and np.array_equal(bits(i1)[m], bits(i2)[m]) and np.array_equal(bits(s1), bits(s2))


def test_esbr_stage_with_hbe_golden(oracle):
    """the composed oracle stage (analysis bank -> harmonic transposer -> HF generator -> envelope adjuster -> regrouping ->
    synthesis bank) against 6 consecutive frames x 2 channels tapped around ixheaacd_sbr_dec in a real -harmonic_sbr:1 decode"""
    from tests.test_oracle_esbr import esbr_stage_golden_frames
    g = np.load(os.path.join(os.path.dirname(GOLD), "esbr_hbe_stage_tapped.npz"))
    rp = oracle_util.esbr_random_phase()
    st = {k: g["in0_" + k] for k in oracle_util.ESH_KEYS}
    for f, r, rg in esbr_stage_golden_frames(g):
        out, st, ipar2, err = oracle_util.oracle_esbr_hbe_stage(oracle, rp, st, g["time_in"][r], g["hbe_cfg"][r], g["hf_par"][r],
                                                                g["ec_ipar_in"][r], g["ec_fpar"][r], rg)
        assert not err.any(), f"frame {f}: {err}"
        assert np.array_equal(bits(out), bits(g["time_out"][r])), f"frame {f}: time output"
        assert np.array_equal(ipar2, g["ec_ipar_out"][r]), f"frame {f}: in/out parameter words"
        for k in ("anal_states", "anal_pos", "synth_states", "synth_pos", "bw_prev", "patch", "ec_state", "hbe_state"):
            assert np.array_equal(st[k].view(np.int32), g["out_" + k][r].view(np.int32)), f"frame {f}: {k}"
    for k in ("qmf_re", "qmf_im", "out_re", "out_im", "pv_re", "pv_im"):
        assert np.array_equal(st[k].view(np.int32), g["out_" + k].view(np.int32)), k
    assert np.abs(g["time_out"]).max() > 100
