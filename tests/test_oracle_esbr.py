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
