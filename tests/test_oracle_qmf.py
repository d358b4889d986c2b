"""CPU tests: the C oracle for the fixed-point HQ QMF banks against the compiled reference on seeded inputs
(cos/sin modulation for both sizes, 64-band synthesis, 32-band analysis, state carry across frames)."""
import numpy as np
import pytest

from tests import oracle_util


def test_qmf_rom_blob_matches_reference(ref):
    assert np.array_equal(oracle_util.rom("qmf_rom.bin"), ref.rom_qmf())


@pytest.mark.parametrize("nch", [64, 32])
def test_cos_sin_mod(oracle, ref, nch):
    rng = np.random.default_rng(nch)
    for t in range(200):
        s = int(rng.integers(2, 32))
        x = rng.integers(-(1 << s), 1 << s, 128, dtype=np.int64).astype(np.int32)
        if t == 0:
            x[:] = 2 ** 31 - 1
        if t == 1:
            x[:] = -(2 ** 31)
        assert np.array_equal(oracle.cos_sin_mod(x, nch), ref.cos_sin_mod(x, nch)), f"trial {t}"


def test_synthesis_matches_reference(oracle, ref):
    n = 96
    matrix, fs, pos, params = oracle_util.synth_qmf_units(n, 11)
    pcm, fs2, pos2 = oracle.synth_batch(matrix, fs, pos, params)
    for u in range(n):
        r_pcm, r_fs, r_pos = ref.synth(matrix[u], fs[u], pos[u], params[u])
        assert np.array_equal(pcm[u], r_pcm), f"unit {u} pcm"
        assert np.array_equal(fs2[u], r_fs), f"unit {u} filter states"
        assert np.array_equal(pos2[u], r_pos), f"unit {u} offsets"


def test_synthesis_interleaved_and_stream(oracle, ref):
    """ch_fac = 2 output stride and 12 consecutive frames of one channel with carried state."""
    rng = np.random.default_rng(3)
    fs_o = np.zeros(1280, np.int16)
    fs_r = fs_o.copy()
    pos_o = np.zeros(2, np.int16)
    pos_r = pos_o.copy()
    for f in range(12):
        matrix, _, _, params = oracle_util.synth_qmf_units(1, 100 + f)
        pcm_o, fs_o2, pos_o2 = oracle.synth_batch(matrix, fs_o[None], pos_o[None], params)
        r_pcm, fs_r, pos_r = ref.synth(matrix[0], fs_r, pos_r, params[0], ch_fac=2)
        fs_o, pos_o = fs_o2[0], pos_o2[0]
        assert np.array_equal(pcm_o[0], r_pcm) and np.array_equal(fs_o, fs_r) and np.array_equal(pos_o, pos_r), f
    assert pos_o[0] == (0 - 12 * 32 * 128) % 1280 and pos_o[1] == (12 * 32 * 64) % 640


def test_analysis_matches_reference(oracle, ref):
    rng = np.random.default_rng(21)
    n = 64
    s = rng.integers(2, 16, size=(n, 1))
    tin = ((rng.random((n, 1024)) * 2 - 1) * (2.0 ** s)).astype(np.int16)
    tin[0] = 32767
    tin[1] = -32768
    st = rng.integers(-32768, 32768, (n, 320)).astype(np.int16)
    pos = np.stack([rng.integers(0, 10, n) * 32, rng.integers(0, 5, n) * 128], 1).astype(np.int16)
    usb = rng.integers(0, 33, n)
    m, st2, pos2 = oracle.anal_batch(tin, st, pos, usb)
    for u in range(n):
        r_m, r_st, r_pos, lb = ref.anal(tin[u], st[u], pos[u], usb[u])
        assert lb == -8
        assert np.array_equal(m[u], r_m), f"unit {u} matrix"
        assert np.array_equal(st2[u], r_st) and np.array_equal(pos2[u], r_pos), f"unit {u} state"


def test_analysis_synthesis_chain_stream(oracle, ref):
    """analysis -> synthesis over 6 frames (no SBR processing): both banks carry state; oracle == reference."""
    rng = np.random.default_rng(8)
    a_st_o = np.zeros((1, 320), np.int16); a_pos_o = np.zeros((1, 2), np.int16)
    a_st_r = np.zeros(320, np.int16); a_pos_r = np.zeros(2, np.int16)
    s_fs_o = np.zeros((1, 1280), np.int16); s_pos_o = np.zeros((1, 2), np.int16)
    s_fs_r = np.zeros(1280, np.int16); s_pos_r = np.zeros(2, np.int16)
    params = np.array([[-8, -8, -8, -6, 32, 32, 6, 0]], np.int16)
    for f in range(6):
        t = np.arange(1024) + 1024 * f
        tin = (8000 * np.sin(2 * np.pi * 0.013 * t) + 200 * rng.standard_normal(1024)).astype(np.int16)
        m_o, a_st_o, a_pos_o = oracle.anal_batch(tin[None], a_st_o, a_pos_o, np.array([32]))
        m_r, a_st_r, a_pos_r, _ = ref.anal(tin, a_st_r, a_pos_r, 32)
        assert np.array_equal(m_o[0], m_r)
        pcm_o, s_fs_o, s_pos_o = oracle.synth_batch(m_o, s_fs_o, s_pos_o, params)
        pcm_r, s_fs_r, s_pos_r = ref.synth(m_r, s_fs_r, s_pos_r, params[0])
        assert np.array_equal(pcm_o[0], pcm_r), f"frame {f}"
    assert np.abs(pcm_o).max() > 100
