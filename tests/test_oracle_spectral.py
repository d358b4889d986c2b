"""CPU tests for the AAC pre-IMDCT spectral stage's checker: the committed golden records (tests/golden/aac_spectral_ref.npz, made
by tools/make_golden.py sps from the COMPILED ixheaacd_channel_pair_process) reproduce, and the record generator exercises every
tool of the stage."""
import os

import numpy as np

from tests import oracle_util as ou

GOLD = os.path.join(os.path.dirname(__file__), "golden", "aac_spectral_ref.npz")


def test_golden_records_reproduce(ref):
    g = np.load(GOLD)
    out, err, seed = ou.ref_channel_pair_process(ref, g["spec_in"], g["rec"], g["seed_in"])
    assert (err == 0).all() and np.array_equal(out, g["spec_out"]) and np.array_equal(seed, g["seed_out"])


def test_golden_covers_the_tools():
    g = np.load(GOLD)
    rec = g["rec"]
    assert rec.shape[1] == ou.SPS_BYTES
    two = rec[:, :4].view(np.int32)[:, 0] == 2
    ch = lambda u, c: rec[u][ou.SPS_CH + c * ou.SPS_CH_BYTES: ou.SPS_CH + (c + 1) * ou.SPS_CH_BYTES]
    assert two.sum() >= 10 and (~two).sum() >= 1
    assert any(ch(u, 0)[:32].view(np.int32)[0] == 2 for u in range(len(rec)))                       # eight-short
    assert any(ch(u, c)[424:428].view(np.int32)[0] for u in range(len(rec)) for c in range(2))     # TNS
    assert any(ch(u, c)[:32].view(np.int32)[3] for u in range(len(rec)) for c in range(2))         # PNS
    assert any((ch(u, 1)[40:168].view(np.int8) >= 14).any() for u in np.flatnonzero(two))           # intensity
    assert (g["spec_out"] != g["spec_in"]).sum() > 5000 and (g["seed_out"] != g["seed_in"]).sum() >= 5


def test_reference_state_chain(ref):
    """three frames with the generator state carried: the same elements in one batch or one at a time give the same result"""
    spec, rec = ou.synth_sps_units(40, 9, pns=True)
    seed = np.arange(40, dtype=np.int32) * 1000003
    a = ou.ref_channel_pair_process(ref, spec, rec, seed)
    for u in (0, 7, 39):
        b = ou.ref_channel_pair_process(ref, spec[u:u + 1], rec[u:u + 1], seed[u:u + 1])
        assert np.array_equal(a[0][u], b[0][0]) and a[2][u] == b[2][0]
