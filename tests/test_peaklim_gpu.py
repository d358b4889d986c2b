"""GPU parity tests for the AAC-LC output stage (xaac_b200_peak_limiter_dev = ixheaacd_peak_limiter_process + round16)
against the CPU oracle (pinned to the compiled reference by tests/test_oracle_peaklim.py) over multi-frame streams in
which the limiter rests, attacks, holds and releases, and against the compiled reference where oracle/_ref is present;
plus the whole AAC-LC stereo chain IMDCT -> limiter -> PCM16."""
import numpy as np
import pytest

from tests import oracle_util

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("ch,fs,n", [(2, 44100, 96), (1, 48000, 40), (2, 48000, 700)])
def test_streams_vs_oracle(ctx, oracle, ch, fs, n):
    import torch
    import libxaac_b200 as xb
    frames = 8
    st = xb.PeakLimiterBatch(n, ch, fs)
    est = np.tile(oracle_util.peak_limiter_reset_state(ch, fs), (n, 1))
    assert np.array_equal(st.state.cpu().numpy(), est)
    engaged = 0
    for f in range(frames):
        x, q = oracle_util.synth_peaklim_units(n, ch, 70 + f, loud_fraction=0.3 if f % 3 else 0.8)
        out32 = torch.empty((n, 1024, ch), dtype=torch.int32, device="cuda")
        pcm, err = xb.peak_limiter_process(ctx, st, torch.from_numpy(x).cuda(), torch.from_numpy(q).cuda(), out32=out32)
        torch.cuda.synchronize()
        est, ey, ep, eerr = oracle.peak_limiter_batch(est, x, q, ch)
        assert int(err.abs().max().item()) == 0 and (eerr == 0).all()
        got = out32.cpu().numpy()
        bad = np.unique(np.argwhere(got != ey)[:, 0])
        assert bad.size == 0, f"frame {f}: WORD32 output differs for units {bad[:8]}"
        assert np.array_equal(pcm.cpu().numpy(), ep), f"frame {f}: PCM16"
        gst = st.state.cpu().numpy()
        assert np.array_equal(gst, est), f"frame {f}: state differs at {np.argwhere(gst != est)[:6].tolist()}"
        engaged += int((est[:, 3].view(np.float32) < 1.0).sum())
    assert engaged > 10


def test_vs_compiled_reference(ctx, ref):
    import torch
    import libxaac_b200 as xb
    n, ch, fs = 64, 2, 44100
    st = xb.PeakLimiterBatch(n, ch, fs)
    rst = np.tile(ref.peak_limiter_init(ch, fs)[0], (n, 1))
    for f in range(5):
        x, q = oracle_util.synth_peaklim_units(n, ch, 90 + f, loud_fraction=0.6)
        pcm, _ = xb.peak_limiter_process(ctx, st, torch.from_numpy(x).cuda(), torch.from_numpy(q).cuda())
        torch.cuda.synchronize()
        rst, _, rp = ref.peak_limiter_batch(rst, x, q, ch)
        assert np.array_equal(pcm.cpu().numpy(), rp), f"frame {f}"
    assert np.array_equal(st.state.cpu().numpy(), rst)


def test_aac_lc_stereo_chain(ctx, oracle):
    """AAC-LC stereo with the reference's default flags: ixheaacd_imdct_process per channel (interleaved WORD32 out,
    ch_fac = 2) -> peak limiter -> round16, 4 frames, against the chained oracles"""
    import torch
    import libxaac_b200 as xb
    n_frames, frames = 150, 4
    n = 2 * n_frames
    ist = xb.ImdctBatch(n)
    lst = xb.PeakLimiterBatch(n_frames, 2, 44100)
    est = np.tile(oracle_util.peak_limiter_reset_state(2, 44100), (n_frames, 1))
    ovl = np.zeros((n, 512), np.int32)
    wstate = np.zeros((n, 2), np.uint8)
    rng = np.random.default_rng(4)
    for f in range(frames):
        s = rng.integers(14, 30, (n, 1))
        spec = ((rng.random((n, 1024)) * 2 - 1) * 2.0 ** s).astype(np.int64).astype(np.int32)
        ics = np.zeros((n, 2), np.uint8)
        ics[:, 1] = np.repeat(rng.integers(0, 2, n_frames), 2)
        w32, adj = xb.imdct_process(ctx, ist, torch.from_numpy(spec).cuda(), torch.from_numpy(ics).cuda(), ch_fac=2)
        pcm, err = xb.peak_limiter_process(ctx, lst, w32.view(n_frames, 1024, 2), adj.view(n_frames, 2))
        torch.cuda.synchronize()
        eo, ovl, wstate, eadj = oracle.imdct_batch(spec, ovl, wstate, ics)
        inter = np.ascontiguousarray(eo.reshape(n_frames, 2, 1024).transpose(0, 2, 1))
        est, _, ep, _ = oracle.peak_limiter_batch(est, inter, eadj.reshape(n_frames, 2), 2)
        assert np.array_equal(pcm.cpu().numpy(), ep), f"frame {f}"
    assert np.array_equal(lst.state.cpu().numpy(), est)
