"""CPU tests: the AAC-LC output stage (ixheaacd_peak_limiter_process + round16) — our phase-split C
restatement against the compiled reference, frame by frame with the limiter state carried."""
import numpy as np
import pytest

from tests import oracle_util


def test_reset_state_matches_reference_init(ref):
    for ch, fs in ((1, 48000), (2, 44100), (2, 32000), (2, 48000)):
        st, delay = ref.peak_limiter_init(ch, fs)
        mine = oracle_util.peak_limiter_reset_state(ch, fs)
        assert delay == mine[6]
        assert np.array_equal(st, mine), f"{ch} ch {fs} Hz: {np.argwhere(st != mine).ravel()[:8]}"


@pytest.mark.parametrize("ch,fs", [(2, 44100), (1, 48000), (2, 48000)])
def test_streams_match_reference(oracle, ref, ch, fs):
    """16 streams x 10 frames: quiet and clipping frames mixed, so that the limiter attacks, holds and releases across
    frame boundaries; samples, PCM16 and the whole state record must agree after every frame"""
    n, frames = 16, 10
    st1 = np.tile(oracle_util.peak_limiter_reset_state(ch, fs), (n, 1))
    st2 = st1.copy()
    engaged = 0
    for f in range(frames):
        x, q = oracle_util.synth_peaklim_units(n, ch, 50 + f, loud_fraction=0.4 if f % 3 else 0.9)
        st1, y1, p1, err = oracle.peak_limiter_batch(st1, x, q, ch)
        st2, y2, p2 = ref.peak_limiter_batch(st2, x, q, ch)
        assert (err == 0).all()
        assert np.array_equal(y1, y2), f"frame {f}: samples differ for units {np.unique(np.argwhere(y1 != y2)[:, 0])[:8]}"
        assert np.array_equal(p1, p2)
        assert np.array_equal(st1, st2), f"frame {f}: state differs at {np.argwhere(st1 != st2)[:6].tolist()}"
        engaged += int((st1[:, 3].view(np.float32) < 1.0).sum())
    assert engaged > 10


def test_product_state_init_matches_reference(ref):
    """xaac_b200_peak_limiter_state_init is host-only code of the product library (no device needed)"""
    import libxaac_b200.output as out
    for ch, fs in ((1, 48000), (2, 44100), (2, 32000), (2, 24000)):
        st, _ = ref.peak_limiter_init(ch, fs)
        assert np.array_equal(out.peak_limiter_reset_state(ch, fs), st)
