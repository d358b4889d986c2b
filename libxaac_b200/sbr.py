"""Host-side mirror of the reference's fixed-point SBR middle stages, batched.

hf_generator  <- ixheaacd_hf_generator(ia_sbr_hf_generator_struct*, ia_sbr_scale_fact_struct*, WORD32 **qmf_real,
                 WORD32 **qmf_imag, time_step, first_slot_offset, last_slot_offset, num_if_bands,
                 max_qmf_subband_aac, sbr_invf_mode, sbr_invf_mode_prev, ...)   (decoder/ixheaacd_lpp_tran.c:956)

calc_sbrenvelope <- ixheaacd_calc_sbrenvelope(ia_sbr_scale_fact_struct*, ia_sbr_calc_env_struct*, header, frame data,
                 prev frame data, WORD32 **anal_buf_real, WORD32 **anal_buf_imag, degree_alias, low_pow_flag = 0, ...)
                 (decoder/ixheaacd_env_calc.c:692)
"""
import ctypes

import torch

from .imdct import _chk, _ptr

HF_PARAM_WORDS = 80
ENV_PARAM_WORDS = 656
ENV_STATE_WORDS = 232
SIDE_WORDS = 1232
SBR_STATE_WORDS = 3920
PS_STATE_WORDS = 3888


def hf_generator(ctx, lpc, matrix, params, bw_prev, hb_scale=None, stream=None):
    """Batched drop-in for ixheaacd_hf_generator (HQ). lpc int32 [n,2,128]; matrix int32 [n,38,128] (in place);
    params int16 [n,80] (XAAC_HF_* layout, include/xaac_b200.h); bw_prev int32 [n,6] (in place)."""
    n = matrix.shape[0]
    _chk(lpc, torch.int32, (n, 2, 128), "lpc", "cuda")
    _chk(matrix, torch.int32, (n, 38, 128), "matrix", "cuda")
    _chk(params, torch.int16, (n, HF_PARAM_WORDS), "params", "cuda")
    _chk(bw_prev, torch.int32, (n, 6), "bw_prev", "cuda")
    if hb_scale is None:
        hb_scale = torch.empty((n,), dtype=torch.int16, device=matrix.device)
    if stream is None:
        stream = torch.cuda.current_stream(matrix.device)
    rc = ctx._lib.xaac_b200_hf_generator_hq_dev(ctx.handle, _ptr(lpc), _ptr(matrix), _ptr(params), _ptr(bw_prev),
                                               _ptr(hb_scale), n, ctypes.c_void_p(stream.cuda_stream))
    ctx.check(rc, "xaac_b200_hf_generator_hq_dev")
    return hb_scale


def calc_sbrenvelope(ctx, params, sf, state, matrix, err=None, stream=None):
    """Batched drop-in for ixheaacd_calc_sbrenvelope (HQ). params int16 [n,656] (XAAC_ENV_* layout); sf int16 [n,8]
    (in place: hb_scale, ov_hb_scale); state int16 [n,232] (in place); matrix int32 [n,38,128] (in place).
    Returns err int32 [n] (0 or 0x80000000 per unit, the reference's return value)."""
    n = matrix.shape[0]
    _chk(params, torch.int16, (n, ENV_PARAM_WORDS), "params", "cuda")
    _chk(sf, torch.int16, (n, 8), "sf", "cuda")
    _chk(state, torch.int16, (n, ENV_STATE_WORDS), "state", "cuda")
    _chk(matrix, torch.int32, (n, 38, 128), "matrix", "cuda")
    if err is None:
        err = torch.empty((n,), dtype=torch.int32, device=matrix.device)
    else:
        _chk(err, torch.int32, (n,), "err", "cuda")
    if stream is None:
        stream = torch.cuda.current_stream(matrix.device)
    rc = ctx._lib.xaac_b200_calc_sbrenvelope_hq_dev(ctx.handle, _ptr(params), _ptr(sf), _ptr(state), _ptr(matrix),
                                                   _ptr(err), n, ctypes.c_void_p(stream.cuda_stream))
    ctx.check(rc, "xaac_b200_calc_sbrenvelope_hq_dev")
    return err


class SbrState:
    """Device-resident state of n SBR channels for sbr_dec (what ia_sbr_dec_struct / ia_ps_dec_struct carry between
    frames).  upload / download move host blobs: numpy int16 [n, 3920] (channel) and [n, 3888] (PS)."""

    def __init__(self, ctx, n_units, with_ps=False, low_power=False):
        """low_power=True: state for sbr_dec_lp only (XAAC_B200_SBR_STATE_LP: no stage scratch in HBM)"""
        self.ctx, self.n_units, self.with_ps = ctx, int(n_units), bool(with_ps) and not low_power
        self.low_power = bool(low_power)
        self._h = ctypes.c_void_p()
        mode = 2 if self.low_power else int(self.with_ps)
        ctx.check(ctx._lib.xaac_b200_sbr_state_create(ctx.handle, self.n_units, mode, ctypes.byref(self._h)),
                  "xaac_b200_sbr_state_create")

    @property
    def handle(self):
        return self._h

    def upload(self, st_blob=None, ps_blob=None):
        import numpy as np
        a = None if st_blob is None else np.ascontiguousarray(st_blob, np.int16)
        b = None if ps_blob is None else np.ascontiguousarray(ps_blob, np.int16)
        assert a is None or a.shape == (self.n_units, SBR_STATE_WORDS)
        assert b is None or b.shape == (self.n_units, PS_STATE_WORDS)
        rc = self.ctx._lib.xaac_b200_sbr_state_upload(self.ctx.handle, self._h,
                                                      None if a is None else a.ctypes.data_as(ctypes.c_void_p),
                                                      None if b is None else b.ctypes.data_as(ctypes.c_void_p))
        self.ctx.check(rc, "xaac_b200_sbr_state_upload")

    def download(self):
        import numpy as np
        a = np.zeros((self.n_units, SBR_STATE_WORDS), np.int16)
        b = np.zeros((self.n_units, PS_STATE_WORDS), np.int16) if self.with_ps else None
        rc = self.ctx._lib.xaac_b200_sbr_state_download(self.ctx.handle, self._h, a.ctypes.data_as(ctypes.c_void_p),
                                                        None if b is None else b.ctypes.data_as(ctypes.c_void_p))
        self.ctx.check(rc, "xaac_b200_sbr_state_download")
        return a, b

    def close(self):
        if self._h:
            self.ctx._lib.xaac_b200_sbr_state_destroy(self.ctx.handle, self._h)
            self._h = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def sbr_dec(ctx, state, side, time_in, time_out=None, err=None, stream=None):
    """Batched drop-in for ixheaacd_sbr_dec (fixed-point HQ path).  side int16 [n,1232]; time_in int16 [n,1024];
    returns (time_out, err): time_out int16 [n,2048] or, for a PS state, [n,2048,2] (interleaved L/R); err int32 [n]."""
    n = state.n_units
    _chk(side, torch.int16, (n, SIDE_WORDS), "side", "cuda")
    _chk(time_in, torch.int16, (n, 1024), "time_in", "cuda")
    shape = (n, 2048, 2) if state.with_ps else (n, 2048)
    if time_out is None:
        time_out = torch.zeros(shape, dtype=torch.int16, device=side.device)
    _chk(time_out, torch.int16, shape, "time_out", "cuda")
    if err is None:
        err = torch.empty((n,), dtype=torch.int32, device=side.device)
    else:
        _chk(err, torch.int32, (n,), "err", "cuda")
    if stream is None:
        stream = torch.cuda.current_stream(side.device)
    rc = ctx._lib.xaac_b200_sbr_dec_hq_dev(ctx.handle, state.handle, _ptr(side), _ptr(time_in), _ptr(time_out), _ptr(err),
                                          ctypes.c_void_p(stream.cuda_stream))
    ctx.check(rc, "xaac_b200_sbr_dec_hq_dev")
    return time_out, err


def sbr_dec_w32(ctx, state, side, w32, qshift_adj, time_out=None, err=None, stream=None):
    """sbr_dec fed with the core coder's WORD32 output (w32 int32 [n,1024], qshift_adj int8 [n], both as written by
    imdct_process): the WORD32 -> WORD16 hand-over of ixheaacd_allocate_sbr_scr (decoder/ixheaacd_api.c:337-370) runs in
    the analysis bank's load.  Same results as imdct_out_to_pcm16(mode 0) followed by sbr_dec."""
    n = state.n_units
    _chk(side, torch.int16, (n, SIDE_WORDS), "side", "cuda")
    _chk(w32, torch.int32, (n, 1024), "w32", "cuda")
    _chk(qshift_adj, torch.int8, (n,), "qshift_adj", "cuda")
    shape = (n, 2048, 2) if state.with_ps else (n, 2048)
    if time_out is None:
        time_out = torch.zeros(shape, dtype=torch.int16, device=side.device)
    _chk(time_out, torch.int16, shape, "time_out", "cuda")
    if err is None:
        err = torch.empty((n,), dtype=torch.int32, device=side.device)
    else:
        _chk(err, torch.int32, (n,), "err", "cuda")
    if stream is None:
        stream = torch.cuda.current_stream(side.device)
    rc = ctx._lib.xaac_b200_sbr_dec_hq_w32_dev(ctx.handle, state.handle, _ptr(side), _ptr(w32), _ptr(qshift_adj),
                                              _ptr(time_out), _ptr(err), ctypes.c_void_p(stream.cuda_stream))
    ctx.check(rc, "xaac_b200_sbr_dec_hq_w32_dev")
    return time_out, err


def sbr_dec_lp(ctx, state, side, time_in, time_out=None, out_ch=1, err=None, stream=None):
    """Batched drop-in for ixheaacd_sbr_dec with low_pow_flag = 1 (the fixed-point path of stereo HE-AACv1,
    decoder/ixheaacd_sbr_dec.c:662; one fused kernel).  side int16 [n,1232]; time_in int16 [n,1024]; returns
    (time_out, err): time_out int16 [n // out_ch, 2048, out_ch] (unit u = channel u % out_ch of frame u // out_ch), or
    [n, 2048] for out_ch = 1; err int32 [n]."""
    n = state.n_units
    _chk(side, torch.int16, (n, SIDE_WORDS), "side", "cuda")
    _chk(time_in, torch.int16, (n, 1024), "time_in", "cuda")
    shape = (n, 2048) if out_ch == 1 else (n // out_ch, 2048, out_ch)
    if time_out is None:
        time_out = torch.zeros(shape, dtype=torch.int16, device=side.device)
    _chk(time_out, torch.int16, shape, "time_out", "cuda")
    if err is None:
        err = torch.empty((n,), dtype=torch.int32, device=side.device)
    else:
        _chk(err, torch.int32, (n,), "err", "cuda")
    if stream is None:
        stream = torch.cuda.current_stream(side.device)
    rc = ctx._lib.xaac_b200_sbr_dec_lp_dev(ctx.handle, state.handle, _ptr(side), _ptr(time_in), _ptr(time_out), int(out_ch),
                                          _ptr(err), ctypes.c_void_p(stream.cuda_stream))
    ctx.check(rc, "xaac_b200_sbr_dec_lp_dev")
    return time_out, err


def sbr_dec_lp_w32(ctx, state, side, w32, qshift_adj, time_out=None, out_ch=1, err=None, stream=None):
    """sbr_dec_lp fed with the core coder's WORD32 output (w32 int32 [n,1024], qshift_adj int8 [n] as written by imdct_process);
    same results as imdct_out_to_pcm16(mode 0) followed by sbr_dec_lp."""
    n = state.n_units
    _chk(side, torch.int16, (n, SIDE_WORDS), "side", "cuda")
    _chk(w32, torch.int32, (n, 1024), "w32", "cuda")
    _chk(qshift_adj, torch.int8, (n,), "qshift_adj", "cuda")
    shape = (n, 2048) if out_ch == 1 else (n // out_ch, 2048, out_ch)
    if time_out is None:
        time_out = torch.zeros(shape, dtype=torch.int16, device=side.device)
    _chk(time_out, torch.int16, shape, "time_out", "cuda")
    if err is None:
        err = torch.empty((n,), dtype=torch.int32, device=side.device)
    else:
        _chk(err, torch.int32, (n,), "err", "cuda")
    if stream is None:
        stream = torch.cuda.current_stream(side.device)
    rc = ctx._lib.xaac_b200_sbr_dec_lp_w32_dev(ctx.handle, state.handle, _ptr(side), _ptr(w32), _ptr(qshift_adj), _ptr(time_out),
                                              int(out_ch), _ptr(err), ctypes.c_void_p(stream.cuda_stream))
    ctx.check(rc, "xaac_b200_sbr_dec_lp_w32_dev")
    return time_out, err


def heaac_frame_host(ctx, imdct_state, sbr_state, spec_coeff, ics, side, pcm, err=None):
    """One HE-AAC frame per unit from host buffers (IMDCT + hand-over + SBR stage, chunked / pipelined inside the
    library; both states stay in HBM).  spec_coeff int32 [n,1024], ics uint8 [n,2], side int16 [n,1232] and pcm int16
    [n,2048] / [n,2048,2] are CPU tensors (pinned recommended); imdct_state is an ImdctHostState."""
    n = sbr_state.n_units
    _chk(spec_coeff, torch.int32, (n, 1024), "spec_coeff", "cpu")
    _chk(ics, torch.uint8, (n, 2), "ics", "cpu")
    _chk(side, torch.int16, (n, SIDE_WORDS), "side", "cpu")
    _chk(pcm, torch.int16, (n, 2048, 2) if sbr_state.with_ps else (n, 2048), "pcm", "cpu")
    if err is not None:
        _chk(err, torch.int32, (n,), "err", "cpu")
    rc = ctx._lib.xaac_b200_heaac_frame_host(ctx.handle, imdct_state._h, sbr_state.handle, _ptr(spec_coeff), _ptr(ics),
                                            _ptr(side), _ptr(pcm), None if err is None else _ptr(err))
    ctx.check(rc, "xaac_b200_heaac_frame_host")
    return pcm


def heaac_lp_frame_host(ctx, imdct_state, sbr_state, spec_coeff, ics, side, pcm, out_ch=2, err=None):
    """One stereo HE-AACv1 frame per `out_ch` units from host buffers (IMDCT + hand-over + fused low-power SBR stage).
    spec_coeff int32 [n,1024], ics uint8 [n,2], side int16 [n,1232], pcm int16 [n // out_ch, 2048, out_ch] (CPU tensors,
    pinned recommended); sbr_state created with low_power=True."""
    n = sbr_state.n_units
    _chk(spec_coeff, torch.int32, (n, 1024), "spec_coeff", "cpu")
    _chk(ics, torch.uint8, (n, 2), "ics", "cpu")
    _chk(side, torch.int16, (n, SIDE_WORDS), "side", "cpu")
    _chk(pcm, torch.int16, (n // out_ch, 2048, out_ch), "pcm", "cpu")
    if err is not None:
        _chk(err, torch.int32, (n,), "err", "cpu")
    rc = ctx._lib.xaac_b200_heaac_lp_frame_host(ctx.handle, imdct_state._h, sbr_state.handle, _ptr(spec_coeff), _ptr(ics),
                                               _ptr(side), _ptr(pcm), int(out_ch), None if err is None else _ptr(err))
    ctx.check(rc, "xaac_b200_heaac_lp_frame_host")
    return pcm
