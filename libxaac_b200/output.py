"""Host-side mirror of the reference's AAC-LC output stage, batched.

Reference: ixheaacd_peak_limiter_process(ia_peak_limiter_struct *peak_limiter, VOID *samples, UWORD32 frame_len,
UWORD8 *qshift_adj) (decoder/ixheaacd_peak_limiter.c:177) + the round16 loop of ixheaacd_dec_execute
(decoder/ixheaacd_api.c:3676-3681).  Unit = one stream:

  PeakLimiterBatch.state <- ia_peak_limiter_struct              int32 [n, 1548] (XAAC_PL_* words)
  samples                <- time_data (WORD32, interleaved)     int32 [n, 1024, ch]
  qshift_adj             <- p_state_aac->qshift_adj             int8  [n, ch]
"""
import ctypes

import numpy as np
import torch

from . import _lib
from .imdct import _chk, _ptr

PL_STATE_WORDS = 1548


def peak_limiter_reset_state(num_channels, sample_rate):
    """ixheaacd_peak_limiter_init (decoder/ixheaacd_peak_limiter.c:45): numpy int32 [1548]"""
    st = np.zeros(PL_STATE_WORDS, np.int32)
    rc = _lib.load().xaac_b200_peak_limiter_state_init(st.ctypes.data_as(ctypes.c_void_p), int(num_channels), int(sample_rate))
    if rc != 0:
        raise _lib.XaacB200Error("xaac_b200_peak_limiter_state_init: unsupported channel count / sample rate")
    return st


class PeakLimiterBatch:
    """Device-resident limiter state of n streams, initialised like ixheaacd_peak_limiter_init."""

    def __init__(self, n_units, num_channels, sample_rate, device="cuda:0"):
        self.n, self.ch = int(n_units), int(num_channels)
        st = peak_limiter_reset_state(num_channels, sample_rate)
        self.state = torch.from_numpy(st).to(device).repeat(self.n, 1).contiguous()


def peak_limiter_process(ctx, state, samples, qshift_adj, pcm16=None, out32=None, err=None, stream=None):
    """Batched drop-in for ixheaacd_peak_limiter_process + round16 on device tensors.  Returns (pcm16, err)."""
    n, ch = state.n, state.ch
    _chk(samples, torch.int32, (n, 1024, ch), "samples", "cuda")
    _chk(qshift_adj, torch.int8, (n, ch), "qshift_adj", "cuda")
    if pcm16 is None:
        pcm16 = torch.empty((n, 1024, ch), dtype=torch.int16, device=samples.device)
    if err is None:
        err = torch.empty((n,), dtype=torch.int32, device=samples.device)
    else:
        _chk(err, torch.int32, (n,), "err", "cuda")
    if stream is None:
        stream = torch.cuda.current_stream(samples.device)
    rc = ctx._lib.xaac_b200_peak_limiter_dev(ctx.handle, _ptr(state.state), _ptr(samples), _ptr(qshift_adj),
                                            None if out32 is None else _ptr(out32), _ptr(pcm16), _ptr(err), n, ch,
                                            ctypes.c_void_p(stream.cuda_stream))
    ctx.check(rc, "xaac_b200_peak_limiter_dev")
    return pcm16, err
