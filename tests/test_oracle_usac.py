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


GOLD = __import__("os").path.join(__import__("os").path.dirname(__file__), "golden", "usac_fd_tapped.npz")


def test_oracle_matches_tapped_real_decode(oracle):
    """records tapped (ld --wrap=ixheaacd_fd_frm_dec) from the reference decoding a real xHE-AAC stream made by the
    reference encoder (tools/make_golden.py usac_fd): every record bit-exact, all four occurring window sequences"""
    g = np.load(GOLD)
    h = g["hdr"]  # ccfl, window_sequence, window_shape, window_shape_prev, td_frame_prev, fac, ec, ret
    assert (h[:, 0] == 1024).all() and (h[:, 4:8] == 0).all() and len(set(h[:, 1].tolist())) >= 4
    out, ov, err = oracle.usac_fd_batch(g["coef"], g["ov_in"], h[:, 1], h[:, 2], h[:, 3])
    assert (err == 0).all()
    assert np.array_equal(out, g["out"]) and np.array_equal(ov, g["ov_out"])
    assert np.abs(g["out"].astype(np.int64)).max() > 1 << 20


def test_oracle_tapped_stream_state_carry(oracle):
    """24 consecutive tapped calls = 12 frames of both channels: the overlap produced for one frame must be the tapped
    overlap input of the channel's next frame, and carrying it reproduces every output"""
    g = np.load(GOLD)
    r0 = int(g["run_start"][0])
    assert r0 >= 0
    for ch in (0, 1):
        ov = g["ov_in"][r0 + ch].copy()
        for k in range(12):
            u = r0 + ch + 2 * k
            h = g["hdr"][u]
            assert np.array_equal(ov, g["ov_in"][u])
            o, ov2, _ = oracle.usac_fd_batch(g["coef"][u:u + 1], ov[None], h[1:2], h[2:3], h[3:4])
            assert np.array_equal(o[0], g["out"][u]), f"record {u}"
            ov = ov2[0]
