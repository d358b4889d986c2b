"""CPU tests: whole fixed-point HQ SBR stage (ixheaacd_sbr_dec incl. parametric stereo).
 * oracle vs whole-stage records tapped from real HE-AAC v1 / v2 decodes of the compiled reference (golden),
 * oracle vs the compiled reference driven through the flat-record shim on perturbed units,
 * PS rotation "as written" vs the -fno-strict-aliasing build of the same reference source,
 * state carried over consecutive frames."""
import os

import numpy as np
import pytest

from tests import oracle_util

GOLD = os.path.join(os.path.dirname(__file__), "golden", "sbrdec_tapped.npz")


def check(got, exp, ps_active, what):
    st, ps, ol, orr, err = got
    est, eps, eol, eorr, eerr = exp
    assert err == eerr, what
    assert np.array_equal(st, est), f"{what}: state differs at {np.argwhere(st != est).ravel()[:8]}"
    assert np.array_equal(ol, eol), f"{what}: left PCM differs"
    if ps_active:
        assert np.array_equal(ps, eps), f"{what}: PS state differs at {np.argwhere(ps != eps).ravel()[:8]}"
        assert np.array_equal(orr, eorr), f"{what}: right PCM differs"


def test_oracle_matches_golden(oracle):
    g = np.load(GOLD)
    n = len(g["side"])
    assert n >= 30 and g["side"][:, 737].any() and not g["side"][:, 737].all()
    for u in range(n):
        got = oracle.sbr_dec(g["side"][u], g["st_in"][u], g["ps_in"][u], g["tin"][u])
        check(got, (g["st_out"][u], g["ps_out"][u], g["out_l"][u], g["out_r"][u], g["hdr"][u][4]), g["side"][u, 737],
              f"record {u}")
    assert np.abs(g["out_l"].astype(np.int32)).max() > 1000


def test_oracle_stream_state_carry(oracle):
    """records 1..11 of each stream are consecutive frames (record 0 is the decoder's init-phase call, after which the
    reference resets its state): feed the oracle its own state"""
    g = np.load(GOLD)
    for base in (1, int(np.argmin(g["side"][:, 737])) + 1):
        st, ps = g["st_in"][base].copy(), g["ps_in"][base].copy()
        for k in range(11):
            u = base + k
            assert g["index"][u] == g["index"][base] + k
            st, ps, ol, orr, err = oracle.sbr_dec(g["side"][u], st, ps, g["tin"][u])
            check((st, ps, ol, orr, err), (g["st_out"][u], g["ps_out"][u], g["out_l"][u], g["out_r"][u], 0),
                  g["side"][u, 737], f"stream frame {u}")


def test_oracle_matches_reference_random(oracle, ref):
    g = np.load(GOLD)
    n = 120
    side, st, ps, tin = oracle_util.synth_sbr_units(n, 77, g)
    side[side[:, 737] == 2, 737] = 1   # the reference as built here
    for u in range(n):
        got = oracle.sbr_dec(side[u], st[u], ps[u], tin[u])
        exp = ref.sbr_dec(side[u], st[u], ps[u], tin[u])
        check(got, exp, side[u, 737], f"unit {u}")


def test_ps_rotation_as_written_matches_nsa_build(oracle):
    ref_nsa = oracle_util.Ref.try_load(oracle_util.REF_NSA_SO)
    if ref_nsa is None:
        pytest.skip("oracle/_ref/libxaac_ref_nsa.so not built")
    g = np.load(GOLD)
    psrec = np.nonzero(g["side"][:, 737])[0]
    rng = np.random.default_rng(5)
    for it in range(60):
        u = psrec[rng.integers(0, len(psrec))]
        s = rng.integers(10, 31)
        m = ((rng.random((38, 128)) * 2 - 1) * 2.0 ** s).astype(np.int64).astype(np.int32)
        sf = g["st_in"][u][324:332].copy()
        sf[0], sf[6] = rng.integers(-6, 6), rng.integers(-8, 4)
        usb, cs = int(g["st_in"][u][332 + 14]), int(rng.integers(-3, 4))
        l1, r1, p1 = oracle.ps_apply_frame(g["side"][u][744:], g["ps_in"][u], m, usb, int(sf[0] - sf[6]), cs, 0)
        l2, r2, p2 = ref_nsa.ps_apply_frame(g["side"][u], g["st_in"][u], g["ps_in"][u], sf, m, usb, cs)
        n = oracle_util.PS_ST_DSP_WORDS
        assert np.array_equal(l1, l2) and np.array_equal(r1, r2) and np.array_equal(p1[:n], p2[:n]), f"iteration {it}"
        assert (l1[:32, 3:usb] != 0).any()


def test_ps_rotation_as_built_matches_reference(oracle, ref):
    g = np.load(GOLD)
    psrec = np.nonzero(g["side"][:, 737])[0]
    rng = np.random.default_rng(6)
    for it in range(30):
        u = psrec[rng.integers(0, len(psrec))]
        m = ((rng.random((38, 128)) * 2 - 1) * 2.0 ** rng.integers(10, 31)).astype(np.int64).astype(np.int32)
        sf = g["st_in"][u][324:332].copy()
        usb, cs = int(g["st_in"][u][332 + 14]), int(rng.integers(-3, 4))
        l1, r1, p1 = oracle.ps_apply_frame(g["side"][u][744:], g["ps_in"][u], m, usb, int(sf[0] - sf[6]), cs, 1)
        l2, r2, p2 = ref.ps_apply_frame(g["side"][u], g["st_in"][u], g["ps_in"][u], sf, m, usb, cs)
        n = oracle_util.PS_ST_DSP_WORDS
        assert np.array_equal(l1, l2) and np.array_equal(r1, r2) and np.array_equal(p1[:n], p2[:n]), f"iteration {it}"
