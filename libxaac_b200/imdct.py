"""Host-side mirror of the reference's IMDCT stage interface, batched.

Reference: ixheaacd_imdct_process(ia_aac_dec_overlap_info*, WORD32 *ptr_spec_coeff, ia_ics_info_struct*,
VOID *out_samples, const WORD16 ch_fac, WORD32 *scratch, ia_aac_dec_tables_struct*, ...)
(decoder/ixheaacd_lpfuncs.c:347-353).  Here the per-channel structs become structure-of-arrays over a batch
of independent units (one unit = one frame x one core channel):

  ImdctBatch.overlap  <- ia_aac_dec_overlap_info.ptr_overlap_buf   int32 [n, 512]
  ImdctBatch.wstate   <- ia_aac_dec_overlap_info.{window_shape, window_sequence}  uint8 [n, 2]
  ics                 <- ia_ics_info_struct.{window_sequence, window_shape}       uint8 [n, 2]
  spec_coeff          <- ptr_spec_coeff                             int32 [n, 1024]
  out_samples         <- out_samples (WORD32, stride ch_fac)        int32 [n, 1024] (ch_fac=1)
  qshift_adj          <- ia_ics_info_struct.qshift_adj              int8  [n]
"""
import ctypes

import torch

ONLY_LONG_SEQUENCE = 0
LONG_START_SEQUENCE = 1
EIGHT_SHORT_SEQUENCE = 2
LONG_STOP_SEQUENCE = 3


def _ptr(t):
    return ctypes.c_void_p(t.data_ptr())


def _chk(t, dtype, shape, name, device_type):
    if t.dtype != dtype or tuple(t.shape) != tuple(shape) or not t.is_contiguous():
        raise ValueError(f"{name}: expected contiguous {dtype} {tuple(shape)}, got {t.dtype} {tuple(t.shape)}")
    if t.device.type != device_type:
        raise ValueError(f"{name}: expected a {device_type} tensor, got {t.device}")


class ImdctBatch:
    """Persistent per-unit state of the stage (the reference's ia_aac_dec_overlap_info), device-resident."""

    def __init__(self, n_units, device="cuda:0"):
        self.n = int(n_units)
        self.overlap = torch.zeros((self.n, 512), dtype=torch.int32, device=device)
        self.wstate = torch.zeros((self.n, 2), dtype=torch.uint8, device=device)  # sine, ONLY_LONG


def imdct_process(ctx, state, spec_coeff, ics, out_samples=None, qshift_adj=None, ch_fac=1, stream=None):
    """Batched drop-in for ixheaacd_imdct_process on device tensors (asynchronous on `stream`)."""
    n = state.n
    _chk(spec_coeff, torch.int32, (n, 1024), "spec_coeff", "cuda")
    _chk(ics, torch.uint8, (n, 2), "ics", "cuda")
    if out_samples is None:
        shape = (n, 1024) if ch_fac == 1 else (n // ch_fac, 1024, ch_fac)
        out_samples = torch.empty(shape, dtype=torch.int32, device=spec_coeff.device)
    if qshift_adj is None:
        qshift_adj = torch.empty((n,), dtype=torch.int8, device=spec_coeff.device)
    if out_samples.numel() != n * 1024 or out_samples.dtype != torch.int32 or not out_samples.is_contiguous():
        raise ValueError("out_samples: expected contiguous int32 with n*1024 elements")
    if stream is None:
        stream = torch.cuda.current_stream(spec_coeff.device)
    rc = ctx._lib.xaac_b200_imdct_process_dev(
        ctx.handle, _ptr(spec_coeff), _ptr(state.overlap), _ptr(state.wstate), _ptr(ics), _ptr(out_samples),
        _ptr(qshift_adj), n, int(ch_fac), ctypes.c_void_p(stream.cuda_stream))
    ctx.check(rc, "xaac_b200_imdct_process_dev")
    return out_samples, qshift_adj


class ImdctHostState:
    """Library-owned device-resident state for the host-buffer entry point (xaac_b200_imdct_state_*)."""

    def __init__(self, ctx, n_units):
        self.ctx = ctx
        self.n = int(n_units)
        self._h = ctypes.c_void_p()
        ctx.check(ctx._lib.xaac_b200_imdct_state_create(ctx.handle, self.n, ctypes.byref(self._h)),
                  "xaac_b200_imdct_state_create")

    def upload(self, overlap, wstate):
        _chk(overlap, torch.int32, (self.n, 512), "overlap", "cpu")
        _chk(wstate, torch.uint8, (self.n, 2), "wstate", "cpu")
        self.ctx.check(self.ctx._lib.xaac_b200_imdct_state_upload(self.ctx.handle, self._h, _ptr(overlap), _ptr(wstate)),
                       "xaac_b200_imdct_state_upload")

    def download(self):
        overlap = torch.empty((self.n, 512), dtype=torch.int32)
        wstate = torch.empty((self.n, 2), dtype=torch.uint8)
        self.ctx.check(self.ctx._lib.xaac_b200_imdct_state_download(self.ctx.handle, self._h, _ptr(overlap), _ptr(wstate)),
                       "xaac_b200_imdct_state_download")
        return overlap, wstate

    def close(self):
        if self._h:
            self.ctx._lib.xaac_b200_imdct_state_destroy(self.ctx.handle, self._h)
            self._h = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def imdct_process_host(ctx, state, spec_coeff, ics, out_samples, qshift_adj, ch_fac=1):
    """Same stage through the host-buffer C-ABI entry point (copies + kernel + copies, synchronous).
    spec_coeff/ics/out_samples/qshift_adj are CPU tensors (pinned memory lets copies overlap the kernel);
    `state` is an ImdctHostState that stays in HBM."""
    n = state.n
    _chk(spec_coeff, torch.int32, (n, 1024), "spec_coeff", "cpu")
    _chk(ics, torch.uint8, (n, 2), "ics", "cpu")
    _chk(qshift_adj, torch.int8, (n,), "qshift_adj", "cpu")
    if out_samples.numel() != n * 1024 or out_samples.dtype != torch.int32 or not out_samples.is_contiguous():
        raise ValueError("out_samples: expected contiguous int32 with n*1024 elements")
    rc = ctx._lib.xaac_b200_imdct_process_host(
        ctx.handle, state._h, _ptr(spec_coeff), _ptr(ics), _ptr(out_samples), _ptr(qshift_adj), int(ch_fac))
    ctx.check(rc, "xaac_b200_imdct_process_host")
    return out_samples, qshift_adj


def imdct_out_to_pcm16(ctx, samples, qshift_adj, mode=0, out=None, stream=None):
    """WORD32 IMDCT output [n,1024] -> PCM16 [n,1024].  mode 0: SBR hand-over of ixheaacd_allocate_sbr_scr
    (decoder/ixheaacd_api.c:337-370); mode 1: AAC-LC output (ixheaacd_scale_adjust + round16, api.c:3676-3681)."""
    n = samples.shape[0]
    _chk(samples, torch.int32, (n, 1024), "samples", "cuda")
    _chk(qshift_adj, torch.int8, (n,), "qshift_adj", "cuda")
    if out is None:
        out = torch.empty((n, 1024), dtype=torch.int16, device=samples.device)
    if stream is None:
        stream = torch.cuda.current_stream(samples.device)
    rc = ctx._lib.xaac_b200_imdct_out_to_pcm16_dev(ctx.handle, _ptr(samples), _ptr(qshift_adj), _ptr(out), n, int(mode),
                                                  ctypes.c_void_p(stream.cuda_stream))
    ctx.check(rc, "xaac_b200_imdct_out_to_pcm16_dev")
    return out
