"""CPU tests: the USAC frequency-domain core transform (ixheaacd_fd_frm_dec with the saturating radix-4
FFT ixheaacd_complex_fft_p2_dec) — our C restatement against the compiled reference."""
import numpy as np
import pytest

from tests import oracle_util


@pytest.mark.parametrize("n", [512, 64])
def test_complex_fft_matches_reference(oracle, ref, n):
    rng = np.random.default_rng(n)
    for t in range(60):
        mag = 2.0 ** rng.integers(8, 32)
        xr = ((rng.random(n) * 2 - 1) * mag).astype(np.int64).clip(-2 ** 31, 2 ** 31 - 1).astype(np.int32)
        xi = ((rng.random(n) * 2 - 1) * mag).astype(np.int64).clip(-2 ** 31, 2 ** 31 - 1).astype(np.int32)
        if t == 0:
            xr[:] = 2 ** 31 - 1
            xi[:] = -2 ** 31
        a1, b1, p1 = oracle.usac_complex_fft(xr, xi, 10 if n == 512 else 7)
        a2, b2, p2 = ref.usac_complex_fft(xr, xi, 10 if n == 512 else 7)
        assert p1 == p2
        assert np.array_equal(a1, a2) and np.array_equal(b1, b2), f"trial {t}"


def test_fd_frm_dec_matches_reference_all_sequences(oracle, ref):
    n = 600
    coef, ov = oracle_util.synth_usac_units(n, 7)
    rng = np.random.default_rng(7)
    seq = rng.integers(0, 5, n).astype(np.int32)
    shape = rng.integers(0, 2, n).astype(np.int32)
    shape_prev = rng.integers(0, 2, n).astype(np.int32)
    o1, v1, e1 = oracle.usac_fd_batch(coef, ov, seq, shape, shape_prev)
    o2, v2, e2 = ref.usac_fd_batch(coef, ov, seq, shape, shape_prev)
    assert np.array_equal(e1, e2)
    for u in range(n):
        assert np.array_equal(o1[u], o2[u]), f"unit {u} (seq {seq[u]}): output differs at {np.argwhere(o1[u] != o2[u]).ravel()[:8]}"
        assert np.array_equal(v1[u], v2[u]), f"unit {u} (seq {seq[u]}): overlap differs"
    assert np.abs(o2.astype(np.int64)).max() > 1000


def test_fd_streams_state_carry(oracle, ref):
    """12 consecutive frames per stream with a legal window-sequence walk, overlap and previous shape carried"""
    n, frames = 64, 12
    seq, shape = oracle_util.usac_seq_walk(n, frames, 3)
    ov1 = np.zeros((n, 1024), np.int32)
    ov2 = ov1.copy()
    prev = np.zeros(n, np.int32)
    for f in range(frames):
        coef, _ = oracle_util.synth_usac_units(n, 100 + f)
        o1, ov1, _ = oracle.usac_fd_batch(coef, ov1, seq[f], shape[f], prev)
        o2, ov2, _ = ref.usac_fd_batch(coef, ov2, seq[f], shape[f], prev)
        assert np.array_equal(o1, o2) and np.array_equal(ov1, ov2), f"frame {f}"
        prev = shape[f]
