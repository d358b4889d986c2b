"""Host-side mirror of the reference's fixed-point SBR QMF bank interface, batched.

Reference: ixheaacd_cplx_synt_qmffilt(WORD32 **qmf_real, WORD32 **qmf_imag, WORD32 split, ...,
ia_sbr_scale_fact_struct *sbr_scale_factor, WORD16 *time_out, ia_sbr_qmf_filter_bank_struct *qmf_bank, ...)
(decoder/ixheaacd_qmf_dec.c:811).  Per-channel structs become structure-of-arrays over a batch of independent
units (one unit = one frame of one output channel):

  matrix         <- qmf_real[i] / qmf_imag[i] rows (slot stride 128 words: re[64] | im[64])  int32 [n, 32, 128]
  filter_states  <- ia_sbr_qmf_filter_bank_struct.filter_states                              int16 [n, 1280]
  pos            <- {ixheaacd_drc_offset, filter_pos_syn - p_filter}                          int16 [n, 2]
  params         <- {ov_lb_scale, lb_scale, hb_scale, st_syn_scale, lsb, usb, split, 0}       int16 [n, 8]
  time_out       <- PCM16, 2048 per unit                                                      int16 [n, 2048]
"""
import ctypes

import torch

from .imdct import _chk, _ptr

ST_SYN_SCALE = -6  # decoder/ixheaacd_sbrdec_initfuncs.c:1135
OP_DELAY = 6       # decoder/ixheaacd_sbr_dec.c:709 (split between overlap and current slots)


def synth_params(ov_lb_scale, lb_scale, hb_scale, lsb, usb, st_syn_scale=ST_SYN_SCALE, split=OP_DELAY):
    """Pack per-unit parameter rows (each argument: int or 1-D integer tensor/array of length n)."""
    cols = [torch.as_tensor(c, dtype=torch.int16).reshape(-1) for c in
            (ov_lb_scale, lb_scale, hb_scale, st_syn_scale, lsb, usb, split, 0)]
    n = max(c.numel() for c in cols)
    cols = [c.expand(n) if c.numel() == 1 else c for c in cols]
    return torch.stack(cols, dim=1).contiguous()


class QmfSynthBatch:
    """Persistent synthesis-bank state of a batch of output channels (device tensors)."""

    def __init__(self, n_units, device="cuda:0"):
        self.n = int(n_units)
        self.filter_states = torch.zeros((self.n, 1280), dtype=torch.int16, device=device)
        self.pos = torch.zeros((self.n, 2), dtype=torch.int16, device=device)


def cplx_synt_qmffilt(ctx, state, matrix, params, time_out=None, ch_fac=1, stream=None):
    """Batched drop-in for ixheaacd_cplx_synt_qmffilt (HQ, 64 bands) on device tensors; asynchronous on `stream`."""
    n = state.n
    _chk(matrix, torch.int32, (n, 32, 128), "matrix", "cuda")
    _chk(params, torch.int16, (n, 8), "params", "cuda")
    if n % ch_fac:
        raise ValueError("n_units must be a multiple of ch_fac")
    if time_out is None:
        shape = (n, 2048) if ch_fac == 1 else (n // ch_fac, 2048, ch_fac)
        time_out = torch.empty(shape, dtype=torch.int16, device=matrix.device)
    if time_out.numel() != n * 2048 or time_out.dtype != torch.int16 or not time_out.is_contiguous():
        raise ValueError("time_out: expected contiguous int16 with n*2048 elements")
    if stream is None:
        stream = torch.cuda.current_stream(matrix.device)
    rc = ctx._lib.xaac_b200_qmf_synth_hq_dev(ctx.handle, _ptr(matrix), _ptr(state.filter_states), _ptr(state.pos),
                                            _ptr(params), _ptr(time_out), n, int(ch_fac),
                                            ctypes.c_void_p(stream.cuda_stream))
    ctx.check(rc, "xaac_b200_qmf_synth_hq_dev")
    return time_out


class QmfSynthHostState:
    """Library-owned device-resident synthesis state for the host-buffer entry point."""

    def __init__(self, ctx, n_units):
        self.ctx = ctx
        self.n = int(n_units)
        self._h = ctypes.c_void_p()
        ctx.check(ctx._lib.xaac_b200_qmf_synth_state_create(ctx.handle, self.n, ctypes.byref(self._h)),
                  "xaac_b200_qmf_synth_state_create")

    def upload(self, filter_states, pos):
        _chk(filter_states, torch.int16, (self.n, 1280), "filter_states", "cpu")
        _chk(pos, torch.int16, (self.n, 2), "pos", "cpu")
        self.ctx.check(self.ctx._lib.xaac_b200_qmf_synth_state_upload(self.ctx.handle, self._h, _ptr(filter_states),
                                                                     _ptr(pos)), "xaac_b200_qmf_synth_state_upload")

    def download(self):
        fs = torch.empty((self.n, 1280), dtype=torch.int16)
        pos = torch.empty((self.n, 2), dtype=torch.int16)
        self.ctx.check(self.ctx._lib.xaac_b200_qmf_synth_state_download(self.ctx.handle, self._h, _ptr(fs), _ptr(pos)),
                       "xaac_b200_qmf_synth_state_download")
        return fs, pos

    def close(self):
        if self._h:
            self.ctx._lib.xaac_b200_qmf_synth_state_destroy(self.ctx.handle, self._h)
            self._h = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def cplx_synt_qmffilt_host(ctx, state, matrix, params, time_out, ch_fac=1):
    """Host-buffer entry point (H2D + kernel + D2H, chunked and pipelined; synchronous)."""
    n = state.n
    _chk(matrix, torch.int32, (n, 32, 128), "matrix", "cpu")
    _chk(params, torch.int16, (n, 8), "params", "cpu")
    if time_out.numel() != n * 2048 or time_out.dtype != torch.int16 or not time_out.is_contiguous():
        raise ValueError("time_out: expected contiguous int16 with n*2048 elements")
    rc = ctx._lib.xaac_b200_qmf_synth_hq_host(ctx.handle, state._h, _ptr(matrix), _ptr(params), _ptr(time_out),
                                             int(ch_fac))
    ctx.check(rc, "xaac_b200_qmf_synth_hq_host")
    return time_out


class QmfAnalBatch:
    """Persistent analysis-bank state of a batch of core channels (device tensors): the 320-sample WORD16 ring
    (anal_filter_states) and {core_samples_buffer offset, filter_pos offset}."""

    def __init__(self, n_units, device="cuda:0"):
        self.n = int(n_units)
        self.states = torch.zeros((self.n, 320), dtype=torch.int16, device=device)
        self.pos = torch.zeros((self.n, 2), dtype=torch.int16, device=device)


ANAL_LB_SCALE_HQ = -8  # sbr_scale_factor->lb_scale set by the stage, generic/ixheaacd_qmf_dec_generic.c:635


def cplx_anal_qmffilt(ctx, state, time_in, usb, matrix=None, ch_fac=1, stream=None):
    """Batched drop-in for ixheaacd_cplx_anal_qmffilt (HQ, 32 bands). time_in: int16 [n,1024] (ch_fac=1) or
    [n/ch_fac,1024,ch_fac]; usb: int16 [n]; matrix: int32 [n,32,128] (rows: re at 0..31, im at 64..95)."""
    n = state.n
    if time_in.numel() != n * 1024 or time_in.dtype != torch.int16 or not time_in.is_contiguous():
        raise ValueError("time_in: expected contiguous int16 with n*1024 elements")
    _chk(usb, torch.int16, (n,), "usb", "cuda")
    if n % ch_fac:
        raise ValueError("n_units must be a multiple of ch_fac")
    if matrix is None:
        matrix = torch.zeros((n, 32, 128), dtype=torch.int32, device=time_in.device)
    _chk(matrix, torch.int32, (n, 32, 128), "matrix", "cuda")
    if stream is None:
        stream = torch.cuda.current_stream(time_in.device)
    rc = ctx._lib.xaac_b200_qmf_anal_hq_dev(ctx.handle, _ptr(time_in), _ptr(state.states), _ptr(state.pos), _ptr(usb),
                                           _ptr(matrix), n, int(ch_fac), ctypes.c_void_p(stream.cuda_stream))
    ctx.check(rc, "xaac_b200_qmf_anal_hq_dev")
    return matrix
