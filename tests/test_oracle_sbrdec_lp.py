"""CPU tests: the low-power (real-valued) fixed-point SBR stage the reference runs for stereo HE-AACv1
(ixheaacd_sbr_dec with low_pow_flag = 1: dct3_32 / dct2_64 banks, low-power HF generator with alias-degree estimation,
low-power envelope functions with alias reduction).  Oracle vs tapped whole-stage records, vs the compiled reference on
perturbed units, and over consecutive frames."""
import os

import numpy as np

from tests import oracle_util

GOLD = os.path.join(os.path.dirname(__file__), "golden", "sbrdec_lp_tapped.npz")


def test_oracle_matches_golden(oracle):
    g = np.load(GOLD)
    assert len(g["side"]) >= 30 and (g["hdr"][:, 5] == 1).all()
    for u in range(len(g["side"])):
        st, out, err = oracle.sbr_dec_lp(g["side"][u], g["st_in"][u], g["tin"][u])
        assert err == g["hdr"][u][4]
        assert np.array_equal(st, g["st_out"][u]), f"record {u}: state differs at {np.argwhere(st != g['st_out'][u]).ravel()[:8]}"
        assert np.array_equal(out, g["out_l"][u]), f"record {u}: PCM"
    assert np.abs(g["out_l"].astype(np.int32)).max() > 1000


def test_oracle_stream_state_carry(oracle):
    """records 2..25 are 12 consecutive frames of the two channels (even / odd records): carry each channel's state"""
    g = np.load(GOLD)
    for ch in (0, 1):
        st = g["st_in"][2 + ch].copy()
        for k in range(12):
            u = 2 + ch + 2 * k
            st, out, err = oracle.sbr_dec_lp(g["side"][u], st, g["tin"][u])
            assert err == 0 and np.array_equal(out, g["out_l"][u]) and np.array_equal(st, g["st_out"][u]), f"record {u}"


def test_oracle_matches_reference_random(oracle, ref):
    g = np.load(GOLD)
    side, st, tin = oracle_util.synth_sbr_lp_units(160, 91, g)
    for u in range(len(side)):
        s1, o1, e1 = oracle.sbr_dec_lp(side[u], st[u], tin[u])
        s2, o2, e2 = ref.sbr_dec_lp(side[u], st[u], tin[u])
        assert e1 == e2, f"unit {u}"
        assert np.array_equal(s1, s2), f"unit {u}: state differs at {np.argwhere(s1 != s2).ravel()[:8]}"
        assert np.array_equal(o1, o2), f"unit {u}: PCM"
