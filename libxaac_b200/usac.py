"""Host-side mirror of the reference's USAC frequency-domain core transform, batched.

Reference: ixheaacd_fd_frm_dec(ia_usac_data_struct *usac_data, WORD32 i_ch) (decoder/ixheaacd_imdct.c:596), called per
channel from ixheaacd_core_coder_data (decoder/ixheaacd_ext_ch_ele.c:991).  The members of ia_usac_data_struct the
function touches become structure-of-arrays over a batch of independent units (one unit = one frame x one core channel):

  UsacFdBatch.overlap <- usac_data->overlap_data_ptr[ch]   int32 [n, 1024]
  UsacFdBatch.wstate  <- usac_data->window_shape_prev[ch]  uint8 [n]
  ics                 <- usac_data->{window_sequence, window_shape}[ch]   uint8 [n, 2]
  coef                <- usac_data->coef_fix[ch]            int32 [n, 1024]
  out                 <- usac_data->output_data_ptr[ch]     int32 [n, 1024] (Q15)

Pure frequency-domain streams only (td_frame_prev = 0, no FAC data): LPD / FAC transitions need the ACELP state and stay
on the host.
"""
import ctypes

import torch

from .imdct import _chk, _ptr

STOP_START_SEQUENCE = 4


class UsacFdBatch:
    """Persistent per-unit state of the stage, device-resident."""

    def __init__(self, n_units, device="cuda:0"):
        self.n = int(n_units)
        self.overlap = torch.zeros((self.n, 1024), dtype=torch.int32, device=device)
        self.wstate = torch.zeros((self.n,), dtype=torch.uint8, device=device)


def usac_fd_frm_dec(ctx, state, coef, ics, out=None, stream=None):
    """Batched drop-in for ixheaacd_fd_frm_dec on device tensors (asynchronous on `stream`)."""
    n = state.n
    _chk(coef, torch.int32, (n, 1024), "coef", "cuda")
    _chk(ics, torch.uint8, (n, 2), "ics", "cuda")
    if out is None:
        out = torch.empty((n, 1024), dtype=torch.int32, device=coef.device)
    _chk(out, torch.int32, (n, 1024), "out", "cuda")
    if stream is None:
        stream = torch.cuda.current_stream(coef.device)
    rc = ctx._lib.xaac_b200_usac_fd_frm_dec_dev(ctx.handle, _ptr(coef), _ptr(state.overlap), _ptr(state.wstate), _ptr(ics),
                                               _ptr(out), n, ctypes.c_void_p(stream.cuda_stream))
    ctx.check(rc, "xaac_b200_usac_fd_frm_dec_dev")
    return out
