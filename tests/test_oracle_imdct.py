"""CPU tests: the C oracle against (a) the golden fixtures tapped from the compiled reference decoding real
reference-encoded streams and (b) the compiled reference itself on seeded synthetic units (all 16 window
sequence transitions, both window shapes, corner-case spectra)."""
import os

import numpy as np
import pytest

from tests import oracle_util

GOLD = os.path.join(os.path.dirname(__file__), "golden", "imdct_tapped.npz")


def test_rom_blob_matches_reference(ref):
    assert np.array_equal(oracle_util.rom(), ref.rom_imdct())


def test_oracle_matches_golden(oracle):
    g = np.load(GOLD)
    hdr = g["hdr"]
    assert len(hdr) >= 20
    seen = set()
    for i in range(len(hdr)):
        _, _, ch_fac, pshape, pseq, wseq, wshape, qadj = [int(x) for x in hdr[i]]
        out, ovl, ps, pq, adj = oracle.imdct_process(g["spec"][i], g["ovl_in"][i], pshape, pseq, wseq, wshape)
        assert np.array_equal(out, g["out"][i]), f"record {i}: PCM mismatch"
        assert np.array_equal(ovl, g["ovl_out"][i]), f"record {i}: overlap mismatch"
        assert (ps, pq, adj) == (wshape, wseq, qadj)
        seen.add((pseq, wseq))
    # real streams contain the legal transitions long->long, long->start, start->short, short->stop, stop->long
    assert {(0, 0), (0, 1), (1, 2), (2, 3), (3, 0)} <= seen


def test_golden_config1_is_sine_long(oracle):
    """BASELINE.json configs[0]: AAC-LC mono 48 kHz long block, sine-window OLA, 1 frame, all 1024 outputs + 512 overlap words."""
    g = np.load(GOLD)
    assert list(g["hdr"][0][2:7]) == [1, 0, 0, 0, 0]
    out, ovl, *_ = oracle.imdct_process(g["spec"][0], g["ovl_in"][0], 0, 0, 0, 0)
    assert np.array_equal(out, g["out"][0]) and np.array_equal(ovl, g["ovl_out"][0])
    assert np.abs(g["out"][0]).max() > 0


@pytest.mark.parametrize("seed", [1, 2])
def test_oracle_matches_reference_random(oracle, ref, seed):
    n = 640
    spec, ovl, wstate, ics = oracle_util.synth_units(n, seed)
    o_out, o_ovl, o_ws, o_adj = oracle.imdct_batch(spec, ovl, wstate, ics)
    combos = set()
    for u in range(n):
        r_out, r_ovl, ps, pq, adj = ref.imdct_process(spec[u], ovl[u], wstate[u, 0], wstate[u, 1], ics[u, 0], ics[u, 1])
        assert np.array_equal(o_out[u], r_out), f"unit {u} PCM"
        assert np.array_equal(o_ovl[u], r_ovl), f"unit {u} overlap"
        assert (int(o_ws[u, 0]), int(o_ws[u, 1]), int(o_adj[u])) == (ps, pq, adj)
        combos.add((int(wstate[u, 1]), int(ics[u, 0])))
    assert len(combos) == 16


def test_oracle_interleaved_stride_matches_reference(oracle, ref):
    spec, ovl, wstate, ics = oracle_util.synth_units(16, 7)
    for u in range(16):
        a = oracle.imdct_process(spec[u], ovl[u], wstate[u, 0], wstate[u, 1], ics[u, 0], ics[u, 1], ch_fac=2)
        b = ref.imdct_process(spec[u], ovl[u], wstate[u, 0], wstate[u, 1], ics[u, 0], ics[u, 1], ch_fac=2)
        assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1]) and a[2:] == b[2:]


def test_oracle_stream_state_carry(oracle, ref):
    """40 consecutive frames of one stream: overlap + window state carried, legal sequence walk."""
    rng = np.random.default_rng(5)
    seqs = [0, 0, 1, 2, 2, 3, 0, 1, 3, 0] * 4
    ovl_o = np.zeros(512, np.int32)
    ovl_r = ovl_o.copy()
    so = sq = ro = rq = 0
    for f, ws in enumerate(seqs):
        spec = ((rng.random(1024) * 2 - 1) * 2.0 ** rng.integers(10, 29)).astype(np.int64).astype(np.int32)
        shape = int(rng.integers(0, 2))
        out_o, ovl_o, so, sq, adj_o = oracle.imdct_process(spec, ovl_o, so, sq, ws, shape)
        out_r, ovl_r, ro, rq, adj_r = ref.imdct_process(spec, ovl_r, ro, rq, ws, shape)
        assert np.array_equal(out_o, out_r) and np.array_equal(ovl_o, ovl_r), f"frame {f}"
        assert (so, sq, adj_o) == (ro, rq, adj_r)
