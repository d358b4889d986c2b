"""Host-side mirror of the reference's eSBR 64-band synthesis bank, batched.

Reference: the per-slot core of ixheaacd_esbr_synthesis_filt_block(ia_sbr_dec_struct *, ..., FLOAT32 **qmf_buf_real,
FLOAT32 **qmf_buf_imag, ...) (decoder/ixheaacd_sbr_dec.c:447, lines 583-654).  Unit = one output channel of one frame:

  EsbrSynthBatch.states <- str_synthesis_qmf_bank.filter_states_32                       int32 [n, 1280]
  EsbrSynthBatch.pos    <- {ixheaacd_drc_offset, filter_pos_syn_32 - esbr_qmf_c}        int32 [n, 2]
  qmf                   <- qmf_buf_real[i][k] | qmf_buf_imag[i][k]                       float32 [n, 32, 128]
  out                   <- time_sample_buf                                               float32 [n, 2048]
"""
import ctypes

import torch

from .imdct import _chk, _ptr


class EsbrSynthBatch:
    def __init__(self, n_units, device="cuda:0"):
        self.n = int(n_units)
        self.states = torch.zeros((self.n, 1280), dtype=torch.int32, device=device)
        self.pos = torch.zeros((self.n, 2), dtype=torch.int32, device=device)


def esbr_synthesis_filt(ctx, state, qmf, out=None, err=None, stream=None, pcm16=None, ch_fac=1, want_float=True):
    """Batched drop-in for the synthesis core of ixheaacd_esbr_synthesis_filt_block.  Returns (out, err).
    With pcm16 (int16 [n / ch_fac, 2048, ch_fac]) the store also applies ixheaacd_samples_sat (clamp + truncation,
    channel-interleaved; unit u = channel u % ch_fac of stream u // ch_fac); want_float=False skips the float output."""
    n = state.n
    _chk(qmf, torch.float32, (n, 32, 128), "qmf", "cuda")
    if out is None and want_float:
        out = torch.empty((n, 2048), dtype=torch.float32, device=qmf.device)
    if out is None and pcm16 is None:
        raise ValueError("esbr_synthesis_filt: neither float output (want_float / out) nor pcm16 requested")
    if out is not None:
        _chk(out, torch.float32, (n, 2048), "out", "cuda")
    if err is None:
        err = torch.empty((n,), dtype=torch.int32, device=qmf.device)
    else:
        _chk(err, torch.int32, (n,), "err", "cuda")
    if stream is None:
        stream = torch.cuda.current_stream(qmf.device)
    if pcm16 is None:
        rc = ctx._lib.xaac_b200_esbr_synth64_dev(ctx.handle, _ptr(qmf), _ptr(state.states), _ptr(state.pos), _ptr(out), _ptr(err),
                                                n, ctypes.c_void_p(stream.cuda_stream))
        ctx.check(rc, "xaac_b200_esbr_synth64_dev")
    else:
        _chk(pcm16, torch.int16, (n // ch_fac, 2048, ch_fac), "pcm16", "cuda")
        rc = ctx._lib.xaac_b200_esbr_synth64_pcm16_dev(ctx.handle, _ptr(qmf), _ptr(state.states), _ptr(state.pos),
                                                      _ptr(out) if out is not None else None, _ptr(pcm16), int(ch_fac),
                                                      _ptr(err), n, ctypes.c_void_p(stream.cuda_stream))
        ctx.check(rc, "xaac_b200_esbr_synth64_pcm16_dev")
    return out, err


class EsbrAnalBatch:
    """State of the eSBR 32-band analysis bank: str_codec_qmf_bank.anal_filter_states_32 int32 [n, 320] and
    {state_new_samples_pos_low_32 offset, filter_pos_32 offset} int32 [n, 2]."""

    def __init__(self, n_units, device="cuda:0"):
        self.n = int(n_units)
        self.states = torch.zeros((self.n, 320), dtype=torch.int32, device=device)
        self.pos = torch.zeros((self.n, 2), dtype=torch.int32, device=device)


def esbr_analysis_filt_block(ctx, state, time_in, qmf=None, err=None, stream=None, ch_fac=1):
    """Batched drop-in for ixheaacd_esbr_analysis_filt_block (32 channels, 32 slots).  time_in float32 [n, 1024]; returns
    (qmf float32 [n, 32, 128] with re at +0..31 and im at +64..95 of every slot row, err).
    time_in may also be the core decoder's own output, the hand-over conversion then happens in the kernel's load:
    int32 [n, 1024] (USAC core, x 2^-15) or int16 [n / ch_fac, 1024, ch_fac] (legacy interleaved PCM)."""
    n = state.n
    if time_in.dtype == torch.int32:
        _chk(time_in, torch.int32, (n, 1024), "time_in", "cuda")
    elif time_in.dtype == torch.int16:
        _chk(time_in, torch.int16, (n // ch_fac, 1024, ch_fac), "time_in", "cuda")
    else:
        _chk(time_in, torch.float32, (n, 1024), "time_in", "cuda")
    if qmf is None:
        qmf = torch.zeros((n, 32, 128), dtype=torch.float32, device=time_in.device)
    _chk(qmf, torch.float32, (n, 32, 128), "qmf", "cuda")
    if err is None:
        err = torch.empty((n,), dtype=torch.int32, device=time_in.device)
    else:
        _chk(err, torch.int32, (n,), "err", "cuda")
    if stream is None:
        stream = torch.cuda.current_stream(time_in.device)
    tail = (_ptr(state.states), _ptr(state.pos), _ptr(qmf), _ptr(err), n, ctypes.c_void_p(stream.cuda_stream))
    if time_in.dtype == torch.int32:
        ctx.check(ctx._lib.xaac_b200_esbr_anal32_core_dev(ctx.handle, _ptr(time_in), *tail), "xaac_b200_esbr_anal32_core_dev")
    elif time_in.dtype == torch.int16:
        ctx.check(ctx._lib.xaac_b200_esbr_anal32_pcm16_dev(ctx.handle, _ptr(time_in), int(ch_fac), *tail),
                  "xaac_b200_esbr_anal32_pcm16_dev")
    else:
        ctx.check(ctx._lib.xaac_b200_esbr_anal32_dev(ctx.handle, _ptr(time_in), *tail), "xaac_b200_esbr_anal32_dev")
    return qmf, err


EHF_ROWS, EHF_PAR_WORDS = 40, 96


def esbr_generate_hf(ctx, src_re, src_im, dst_re, dst_im, par, bw_prev, pv_re=None, pv_im=None, patch_out=None, err=None,
                     stream=None):
    """Batched drop-in for ixheaacd_generate_hf (decoder/ixheaacd_sbrdec_lpfuncs.c:981), 2:1 system without
    pre-processing.  QMF buffers float32 [n, 40, 64] = the reference's arrays from their first row (row r = row r - 2 of the
    pointers the reference passes); dst_* and bw_prev [n, 6] are updated in place; par int32 [n, 96] (XAAC_EHF_* words).
    Returns (patch_out int32 [n, 8] = {num_patches, start_subband[7]}, err int32 [n])."""
    n = par.shape[0]
    dev = par.device
    for t, nm in ((src_re, "src_re"), (src_im, "src_im"), (dst_re, "dst_re"), (dst_im, "dst_im")):
        _chk(t, torch.float32, (n, EHF_ROWS, 64), nm, "cuda")
    if (pv_re is None) != (pv_im is None):
        raise ValueError("pv_re and pv_im must both be given or both be None")
    if pv_re is not None:
        _chk(pv_re, torch.float32, (n, EHF_ROWS, 64), "pv_re", "cuda")
        _chk(pv_im, torch.float32, (n, EHF_ROWS, 64), "pv_im", "cuda")
    _chk(par, torch.int32, (n, EHF_PAR_WORDS), "par", "cuda")
    _chk(bw_prev, torch.float32, (n, 6), "bw_prev", "cuda")
    if patch_out is None:
        patch_out = torch.zeros((n, 8), dtype=torch.int32, device=dev)
    _chk(patch_out, torch.int32, (n, 8), "patch_out", "cuda")
    if err is None:
        err = torch.empty((n,), dtype=torch.int32, device=dev)
    else:
        _chk(err, torch.int32, (n,), "err", "cuda")
    if stream is None:
        stream = torch.cuda.current_stream(dev)
    rc = ctx._lib.xaac_b200_esbr_generate_hf_dev(ctx.handle, _ptr(src_re), _ptr(src_im), _ptr(pv_re) if pv_re is not None else None,
                                                 _ptr(pv_im) if pv_im is not None else None, _ptr(dst_re), _ptr(dst_im),
                                                 _ptr(par), _ptr(bw_prev), _ptr(patch_out), _ptr(err), n,
                                                 ctypes.c_void_p(stream.cuda_stream))
    ctx.check(rc, "xaac_b200_esbr_generate_hf_dev")
    return patch_out, err


EEC_IPAR_WORDS, EEC_FPAR_WORDS, EEC_STATE_WORDS = 288, 464, 640


def esbr_env_calc(ctx, re, im, ipar, fpar, state, err=None, stream=None, low_re=None, low_im=None):
    """Batched drop-in for ixheaacd_sbr_env_calc (decoder/ixheaacd_esbr_envcal.c:71), ORIG_SBR branch of the 2:1 system.
    re / im float32 [n, 40, 64] (sbr_qmf_out_real / imag from their first row) are adjusted in place; ipar int32 [n, 288]
    (XAAC_EEC_* words; env_short_flag_prev, harm_index, phase_index, esbr_start_up and harm_flag_prev are updated in
    place), fpar float32 [n, 464] (envelope scale factors | noise floor), state float32 [n, 640] (e_gain | noise_buf,
    updated in place).  low_re / low_im float32 [n, 40 | 72, 64] (qmf_buf_real / imag from their first row) are needed when an
    envelope uses inter-TES (ixheaacd_apply_inter_tes); without them such a frame returns -2.  Returns err int32 [n]."""
    n = ipar.shape[0]
    dev = ipar.device
    _chk(re, torch.float32, (n, EHF_ROWS, 64), "re", "cuda")
    _chk(im, torch.float32, (n, EHF_ROWS, 64), "im", "cuda")
    _chk(ipar, torch.int32, (n, EEC_IPAR_WORDS), "ipar", "cuda")
    _chk(fpar, torch.float32, (n, EEC_FPAR_WORDS), "fpar", "cuda")
    _chk(state, torch.float32, (n, EEC_STATE_WORDS), "state", "cuda")
    if err is None:
        err = torch.empty((n,), dtype=torch.int32, device=dev)
    else:
        _chk(err, torch.int32, (n,), "err", "cuda")
    if stream is None:
        stream = torch.cuda.current_stream(dev)
    if low_re is not None:
        rows = int(low_re.shape[1])
        _chk(low_re, torch.float32, (n, rows, 64), "low_re", "cuda")
        _chk(low_im, torch.float32, (n, rows, 64), "low_im", "cuda")
        rc = ctx._lib.xaac_b200_esbr_env_calc_tes_dev(ctx.handle, _ptr(re), _ptr(im), _ptr(low_re), _ptr(low_im), rows, _ptr(ipar),
                                                      _ptr(fpar), _ptr(state), _ptr(err), n, ctypes.c_void_p(stream.cuda_stream))
        ctx.check(rc, "xaac_b200_esbr_env_calc_tes_dev")
        return err
    rc = ctx._lib.xaac_b200_esbr_env_calc_dev(ctx.handle, _ptr(re), _ptr(im), _ptr(ipar), _ptr(fpar), _ptr(state), _ptr(err), n,
                                              ctypes.c_void_p(stream.cuda_stream))
    ctx.check(rc, "xaac_b200_esbr_env_calc_dev")
    return err


class _EsbrStateView(ctypes.Structure):
    _fields_ = [(k, ctypes.c_void_p) for k in ("qmf_re", "qmf_im", "out_re", "out_im", "anal_states", "anal_pos", "synth_states",
                                               "synth_pos", "bw_prev", "patch", "ec_state")]


class EsbrDecBatch:
    """Per-channel state of the float eSBR stage, resident in HBM (xaac_b200_esbr_state_view): QMF history arrays, both bank
    states, chirp factors, patch table, envelope-adjuster smoothing history."""
    SHAPES = dict(qmf_re=((40, 64), torch.float32), qmf_im=((40, 64), torch.float32), out_re=((40, 64), torch.float32),
                  out_im=((40, 64), torch.float32), anal_states=((320,), torch.int32), anal_pos=((2,), torch.int32),
                  synth_states=((1280,), torch.int32), synth_pos=((2,), torch.int32), bw_prev=((6,), torch.float32),
                  patch=((8,), torch.int32), ec_state=((640,), torch.float32))

    def __init__(self, n_units, device="cuda:0"):
        self.n = int(n_units)
        for k, (shp, dt) in self.SHAPES.items():
            setattr(self, k, torch.zeros((self.n,) + shp, dtype=dt, device=device))

    def view(self):
        return _EsbrStateView(**{k: ctypes.c_void_p(getattr(self, k).data_ptr()) for k in self.SHAPES})


def esbr_dec(ctx, state, core, hf_par, ec_ipar, ec_fpar, rg_par, out=None, pcm16=None, ch_fac=1, err=None, stream=None,
             want_float=True):
    """Batched drop-in for the eSBR branch of ixheaacd_sbr_dec (USAC channel, no harmonic transposer / PS / MPS): core float32
    or int32 [n, 1024] -> float32 [n, 2048] and / or interleaved PCM16 [n / ch_fac, 2048, ch_fac].  Returns (out, err[4, n])."""
    n = state.n
    dev = hf_par.device
    _chk(core, core.dtype if core.dtype in (torch.float32, torch.int32) else torch.float32, (n, 1024), "core", "cuda")
    _chk(hf_par, torch.int32, (n, EHF_PAR_WORDS), "hf_par", "cuda")
    _chk(ec_ipar, torch.int32, (n, EEC_IPAR_WORDS), "ec_ipar", "cuda")
    _chk(ec_fpar, torch.float32, (n, EEC_FPAR_WORDS), "ec_fpar", "cuda")
    _chk(rg_par, torch.int32, (n, 4), "rg_par", "cuda")
    if out is None and want_float:
        out = torch.empty((n, 2048), dtype=torch.float32, device=dev)
    if out is not None:
        _chk(out, torch.float32, (n, 2048), "out", "cuda")
    if pcm16 is not None:
        _chk(pcm16, torch.int16, (n // ch_fac, 2048, ch_fac), "pcm16", "cuda")
    if err is None:
        err = torch.empty((4, n), dtype=torch.int32, device=dev)
    else:
        _chk(err, torch.int32, (4, n), "err", "cuda")
    if stream is None:
        stream = torch.cuda.current_stream(dev)
    v = state.view()
    rc = ctx._lib.xaac_b200_esbr_dec_dev(ctx.handle, ctypes.byref(v), _ptr(core) if core.dtype == torch.float32 else None,
                                         _ptr(core) if core.dtype == torch.int32 else None, _ptr(hf_par), _ptr(ec_ipar),
                                         _ptr(ec_fpar), _ptr(rg_par), _ptr(out) if out is not None else None,
                                         _ptr(pcm16) if pcm16 is not None else None, int(ch_fac), _ptr(err), n,
                                         ctypes.c_void_p(stream.cuda_stream))
    ctx.check(rc, "xaac_b200_esbr_dec_dev")
    return out, err


HBE_CFG_WORDS, HBE_ST_WORDS = 16, 3616


class EsbrHbeBatch:
    """State of n harmonic-transposer instances (ia_esbr_hbe_txposer_struct): float32 [n, 3616] in the XAAC_HBE_ST_* layout."""

    def __init__(self, n_units, device="cuda:0"):
        self.n = int(n_units)
        self.state = torch.zeros((self.n, HBE_ST_WORDS), dtype=torch.float32, device=device)


def esbr_qmf_hbe_apply(ctx, hbe, qmf_re, qmf_im, pv_re, pv_im, cfg, err=None, stream=None):
    """Batched drop-in for ixheaacd_qmf_hbe_apply (decoder/ixheaacd_hbe_trans.c:224): qmf_re / qmf_im float32 [n, 32, 64] are
    the frame's new QMF slots, pv_re / pv_im float32 [n, 32, 64] receive bands start_band..end_band-1 of the phase-vocoder
    output, cfg int32 [n, 16] (XAAC_HBE_* words), hbe.state is updated in place.  Returns err int32 [n]."""
    n = hbe.n
    for t, nm in ((qmf_re, "qmf_re"), (qmf_im, "qmf_im"), (pv_re, "pv_re"), (pv_im, "pv_im")):
        _chk(t, torch.float32, (n, 32, 64), nm, "cuda")
    _chk(cfg, torch.int32, (n, HBE_CFG_WORDS), "cfg", "cuda")
    _chk(hbe.state, torch.float32, (n, HBE_ST_WORDS), "state", "cuda")
    if err is None:
        err = torch.empty((n,), dtype=torch.int32, device=cfg.device)
    else:
        _chk(err, torch.int32, (n,), "err", "cuda")
    if stream is None:
        stream = torch.cuda.current_stream(cfg.device)
    rc = ctx._lib.xaac_b200_esbr_hbe_apply_dev(ctx.handle, _ptr(qmf_re), _ptr(qmf_im), _ptr(pv_re), _ptr(pv_im), _ptr(cfg),
                                               _ptr(hbe.state), _ptr(err), n, ctypes.c_void_p(stream.cuda_stream))
    ctx.check(rc, "xaac_b200_esbr_hbe_apply_dev")
    return err


class _EsbrHbeStateView(ctypes.Structure):
    _fields_ = [("base", _EsbrStateView), ("pv_re", ctypes.c_void_p), ("pv_im", ctypes.c_void_p), ("hbe_state", ctypes.c_void_p)]


class EsbrDecHbeBatch(EsbrDecBatch):
    """State of the eSBR stage with the harmonic transposer (xaac_b200_esbr_hbe_state_view): the core QMF arrays carry the
    transposer's 32-slot delay (72 rows), plus the phase-vocoder arrays and the transposer instances."""
    SHAPES = dict(EsbrDecBatch.SHAPES, qmf_re=((72, 64), torch.float32), qmf_im=((72, 64), torch.float32),
                  pv_re=((40, 64), torch.float32), pv_im=((40, 64), torch.float32), hbe_state=((HBE_ST_WORDS,), torch.float32))

    def view(self):
        base = _EsbrStateView(**{k: ctypes.c_void_p(getattr(self, k).data_ptr()) for k in EsbrDecBatch.SHAPES})
        return _EsbrHbeStateView(base, ctypes.c_void_p(self.pv_re.data_ptr()), ctypes.c_void_p(self.pv_im.data_ptr()),
                                 ctypes.c_void_p(self.hbe_state.data_ptr()))


def esbr_dec_hbe(ctx, state, core, hbe_cfg, hf_par, ec_ipar, ec_fpar, rg_par, out=None, pcm16=None, ch_fac=1, err=None,
                 stream=None, want_float=True):
    """Batched drop-in for the eSBR branch of ixheaacd_sbr_dec with hbe_flag = 1 (USAC channel, harmonic transposer between
    the analysis bank and the HF generator).  Arguments as esbr_dec plus hbe_cfg int32 [n, 16]; state is an EsbrDecHbeBatch.
    Returns (out, err[5, n])."""
    n = state.n
    dev = hf_par.device
    _chk(core, core.dtype if core.dtype in (torch.float32, torch.int32) else torch.float32, (n, 1024), "core", "cuda")
    _chk(hbe_cfg, torch.int32, (n, HBE_CFG_WORDS), "hbe_cfg", "cuda")
    _chk(hf_par, torch.int32, (n, EHF_PAR_WORDS), "hf_par", "cuda")
    _chk(ec_ipar, torch.int32, (n, EEC_IPAR_WORDS), "ec_ipar", "cuda")
    _chk(ec_fpar, torch.float32, (n, EEC_FPAR_WORDS), "ec_fpar", "cuda")
    _chk(rg_par, torch.int32, (n, 4), "rg_par", "cuda")
    if out is None and want_float:
        out = torch.empty((n, 2048), dtype=torch.float32, device=dev)
    if out is None and pcm16 is None:
        raise ValueError("esbr_dec_hbe: neither float output nor pcm16 requested")
    if out is not None:
        _chk(out, torch.float32, (n, 2048), "out", "cuda")
    if pcm16 is not None:
        _chk(pcm16, torch.int16, (n // ch_fac, 2048, ch_fac), "pcm16", "cuda")
    if err is None:
        err = torch.empty((5, n), dtype=torch.int32, device=dev)
    else:
        _chk(err, torch.int32, (5, n), "err", "cuda")
    if stream is None:
        stream = torch.cuda.current_stream(dev)
    v = state.view()
    rc = ctx._lib.xaac_b200_esbr_dec_hbe_dev(ctx.handle, ctypes.byref(v), _ptr(core) if core.dtype == torch.float32 else None,
                                             _ptr(core) if core.dtype == torch.int32 else None, _ptr(hbe_cfg), _ptr(hf_par),
                                             _ptr(ec_ipar), _ptr(ec_fpar), _ptr(rg_par), _ptr(out) if out is not None else None,
                                             _ptr(pcm16) if pcm16 is not None else None, int(ch_fac), _ptr(err), n,
                                             ctypes.c_void_p(stream.cuda_stream))
    ctx.check(rc, "xaac_b200_esbr_dec_hbe_dev")
    return out, err


# ---- float parametric stereo (ixheaacd_esbr_apply_ps, decoder/ixheaacd_ps_dec_flt.c:381-505) ----
FPS_SIDE_WORDS, FPS_ST_WORDS = 1024, 4368


def esbr_apply_ps(ctx, low_re, low_im, high_re, high_im, rg_par, side, state, left=None, right=None, err=None, stream=None):
    """Batched drop-in for ixheaacd_esbr_apply_ps in the 20-band configuration, fused with the regrouping and the look-ahead
    slots of decoder/ixheaacd_sbr_dec.c:481-517.  low_re / low_im float32 [n, 40 | 72, 64] (qmf_buf_real / imag), high_re /
    high_im float32 [n, 40, 64] (sbr_qmf_out_real / imag), rg_par int32 [n, 4], side float32 [n, 1024] (XAAC_FPS_SIDE_*), state
    float32 [n, 4368] (XAAC_FPS_ST_*, updated in place).  Returns (left, right, err): float32 [n, 32, 128] per slot
    re[64] | im[64], err int32 [n] (0, or -2 for borders outside the supported subset)."""
    n = int(state.shape[0])
    rows = int(low_re.shape[1])
    for t, nm in ((low_re, "low_re"), (low_im, "low_im")):
        _chk(t, torch.float32, (n, rows, 64), nm, "cuda")
    for t, nm in ((high_re, "high_re"), (high_im, "high_im")):
        _chk(t, torch.float32, (n, 40, 64), nm, "cuda")
    _chk(rg_par, torch.int32, (n, 4), "rg_par", "cuda")
    _chk(side, torch.float32, (n, FPS_SIDE_WORDS), "side", "cuda")
    _chk(state, torch.float32, (n, FPS_ST_WORDS), "state", "cuda")
    dev = state.device
    if left is None:
        left = torch.empty((n, 32, 128), dtype=torch.float32, device=dev)
    if right is None:
        right = torch.empty((n, 32, 128), dtype=torch.float32, device=dev)
    _chk(left, torch.float32, (n, 32, 128), "left", "cuda")
    _chk(right, torch.float32, (n, 32, 128), "right", "cuda")
    if err is None:
        err = torch.empty((n,), dtype=torch.int32, device=dev)
    else:
        _chk(err, torch.int32, (n,), "err", "cuda")
    if stream is None:
        stream = torch.cuda.current_stream(dev)
    rc = ctx._lib.xaac_b200_esbr_ps_apply_dev(ctx.handle, _ptr(low_re), _ptr(low_im), rows, _ptr(high_re), _ptr(high_im),
                                              _ptr(rg_par), _ptr(side), _ptr(state), _ptr(left), _ptr(right), _ptr(err), n,
                                              ctypes.c_void_p(stream.cuda_stream))
    ctx.check(rc, "xaac_b200_esbr_ps_apply_dev")
    return left, right, err


class _EsbrPsView(ctypes.Structure):
    _fields_ = [("ps_state", ctypes.c_void_p), ("left", ctypes.c_void_p), ("right", ctypes.c_void_p),
                ("synth_states_r", ctypes.c_void_p), ("synth_pos_r", ctypes.c_void_p)]


class EsbrDecPsBatch(EsbrDecHbeBatch):
    """State of the eSBR stage of a mono + PS element: the harmonic-transposer stage's state (hbe=False drops the transposer
    and its 32-slot delay) plus the PS instance, the second channel's synthesis bank and the two QMF matrices in flight."""

    def __init__(self, n_units, device="cuda:0", hbe=True):
        self.hbe = bool(hbe)
        self.SHAPES = dict(EsbrDecHbeBatch.SHAPES if hbe else EsbrDecBatch.SHAPES, ps_state=((FPS_ST_WORDS,), torch.float32),
                           left=((32, 128), torch.float32), right=((32, 128), torch.float32),
                           synth_states_r=((1280,), torch.int32), synth_pos_r=((2,), torch.int32))
        EsbrDecBatch.__init__(self, n_units, device)

    def view(self):
        base = _EsbrStateView(**{k: ctypes.c_void_p(getattr(self, k).data_ptr()) for k in EsbrDecBatch.SHAPES})
        pv = [ctypes.c_void_p(getattr(self, k).data_ptr()) if self.hbe else None for k in ("pv_re", "pv_im", "hbe_state")]
        return _EsbrHbeStateView(base, *pv)

    def ps_view(self):
        return _EsbrPsView(*[ctypes.c_void_p(getattr(self, k).data_ptr())
                             for k in ("ps_state", "left", "right", "synth_states_r", "synth_pos_r")])


def esbr_dec_ps(ctx, state, core, hbe_cfg, hf_par, ec_ipar, ec_fpar, rg_par, ps_side, out_l=None, out_r=None, err=None,
                stream=None):
    """Batched drop-in for the eSBR branch of ixheaacd_sbr_dec for a mono + PS element (channel_mode == PS_STEREO or
    enh_sbr_ps, decoder/ixheaacd_sbr_dec.c:976-1001).  Arguments as esbr_dec_hbe (hbe_cfg = None without the transposer) plus
    ps_side float32 [n, 1024]; state is an EsbrDecPsBatch.  Returns (out_l, out_r, err[6, n]) with float32 [n, 2048] outputs."""
    n = state.n
    dev = hf_par.device
    _chk(core, core.dtype if core.dtype in (torch.float32, torch.int32) else torch.float32, (n, 1024), "core", "cuda")
    if state.hbe:
        _chk(hbe_cfg, torch.int32, (n, HBE_CFG_WORDS), "hbe_cfg", "cuda")
    _chk(hf_par, torch.int32, (n, EHF_PAR_WORDS), "hf_par", "cuda")
    _chk(ec_ipar, torch.int32, (n, EEC_IPAR_WORDS), "ec_ipar", "cuda")
    _chk(ec_fpar, torch.float32, (n, EEC_FPAR_WORDS), "ec_fpar", "cuda")
    _chk(rg_par, torch.int32, (n, 4), "rg_par", "cuda")
    _chk(ps_side, torch.float32, (n, FPS_SIDE_WORDS), "ps_side", "cuda")
    if out_l is None:
        out_l = torch.empty((n, 2048), dtype=torch.float32, device=dev)
    if out_r is None:
        out_r = torch.empty((n, 2048), dtype=torch.float32, device=dev)
    _chk(out_l, torch.float32, (n, 2048), "out_l", "cuda")
    _chk(out_r, torch.float32, (n, 2048), "out_r", "cuda")
    if err is None:
        err = torch.empty((6, n), dtype=torch.int32, device=dev)
        if not state.hbe:
            err[4].zero_()
    else:
        _chk(err, torch.int32, (6, n), "err", "cuda")
    if stream is None:
        stream = torch.cuda.current_stream(dev)
    v, pv = state.view(), state.ps_view()
    rc = ctx._lib.xaac_b200_esbr_dec_ps_dev(ctx.handle, ctypes.byref(v), ctypes.byref(pv),
                                            _ptr(core) if core.dtype == torch.float32 else None,
                                            _ptr(core) if core.dtype == torch.int32 else None,
                                            _ptr(hbe_cfg) if state.hbe else None, _ptr(hf_par), _ptr(ec_ipar), _ptr(ec_fpar),
                                            _ptr(rg_par), _ptr(ps_side), _ptr(out_l), _ptr(out_r), _ptr(err), n,
                                            ctypes.c_void_p(stream.cuda_stream))
    ctx.check(rc, "xaac_b200_esbr_dec_ps_dev")
    return out_l, out_r, err


# ---- the stage in two halves / pass-through (include/xaac_b200.h: xaac_b200_esbr_dec_front_dev / _back_dev / _bypass_dev) ----
def _hbe_view(state):
    """xaac_b200_esbr_hbe_state_view of an EsbrDecBatch / EsbrDecHbeBatch / EsbrDecPsBatch (pv_* NULL without the transposer)"""
    if isinstance(state, EsbrDecPsBatch):
        return state.view()
    if isinstance(state, EsbrDecHbeBatch):
        return state.view()
    return _EsbrHbeStateView(state.view(), None, None, None)


def esbr_dec_front(ctx, state, core, hbe_cfg, hf_par, err, stream=None):
    """First half of the eSBR stage: analysis bank, harmonic transposer (when the state carries one), HF generator.  err int32
    [6, n]: rows 0, 1, 4 are written.  Afterwards state.patch holds the patch table of the frame — what a host needs to rebuild
    the limiter tables (ixheaacd_createlimiterbands) on reset frames and on frames where sbr_patching_mode changes."""
    n = state.n
    dev = hf_par.device
    _chk(hf_par, torch.int32, (n, EHF_PAR_WORDS), "hf_par", "cuda")
    _chk(err, torch.int32, (6, n), "err", "cuda")
    if stream is None:
        stream = torch.cuda.current_stream(dev)
    v = _hbe_view(state)
    rc = ctx._lib.xaac_b200_esbr_dec_front_dev(ctx.handle, ctypes.byref(v), _ptr(core) if core.dtype == torch.float32 else None,
                                               _ptr(core) if core.dtype == torch.int32 else None,
                                               _ptr(hbe_cfg) if hbe_cfg is not None else None, _ptr(hf_par), _ptr(err), n,
                                               ctypes.c_void_p(stream.cuda_stream))
    ctx.check(rc, "xaac_b200_esbr_dec_front_dev")
    return err


def esbr_dec_back(ctx, state, ec_ipar, ec_fpar, rg_par, err, ps_side=None, out=None, out_r=None, pcm16=None, ch_fac=1, stream=None):
    """Second half: envelope adjuster, PS (state is an EsbrDecPsBatch and ps_side is given), synthesis bank(s).  err rows 2, 3, 5."""
    n = state.n
    dev = ec_ipar.device
    _chk(ec_ipar, torch.int32, (n, EEC_IPAR_WORDS), "ec_ipar", "cuda")
    _chk(ec_fpar, torch.float32, (n, EEC_FPAR_WORDS), "ec_fpar", "cuda")
    _chk(rg_par, torch.int32, (n, 4), "rg_par", "cuda")
    _chk(err, torch.int32, (6, n), "err", "cuda")
    with_ps = ps_side is not None
    if out is None and pcm16 is None:
        out = torch.empty((n, 2048), dtype=torch.float32, device=dev)
    if with_ps and out_r is None:
        out_r = torch.empty((n, 2048), dtype=torch.float32, device=dev)
    if stream is None:
        stream = torch.cuda.current_stream(dev)
    v = _hbe_view(state)
    pv = state.ps_view() if with_ps else None
    rc = ctx._lib.xaac_b200_esbr_dec_back_dev(ctx.handle, ctypes.byref(v), ctypes.byref(pv) if with_ps else None, _ptr(ec_ipar),
                                              _ptr(ec_fpar), _ptr(rg_par), _ptr(ps_side) if with_ps else None,
                                              _ptr(out) if out is not None else None, _ptr(out_r) if with_ps else None,
                                              _ptr(pcm16) if pcm16 is not None else None, int(ch_fac), _ptr(err), n,
                                              ctypes.c_void_p(stream.cuda_stream))
    ctx.check(rc, "xaac_b200_esbr_dec_back_dev")
    return out, out_r, err


def esbr_dec_bypass(ctx, state, core, rg_par, err, out=None, out_r=None, stream=None):
    """apply_processing = 0: analysis bank, sbr_qmf_out cleared, synthesis bank(s) over the regrouped core bands (rg_par =
    {x, sub_band_start, 0, 0}); with an EsbrDecPsBatch the same matrix also leaves through the second channel's bank."""
    n = state.n
    dev = rg_par.device
    _chk(rg_par, torch.int32, (n, 4), "rg_par", "cuda")
    _chk(err, torch.int32, (6, n), "err", "cuda")
    with_ps = isinstance(state, EsbrDecPsBatch)
    if out is None:
        out = torch.empty((n, 2048), dtype=torch.float32, device=dev)
    if with_ps and out_r is None:
        out_r = torch.empty((n, 2048), dtype=torch.float32, device=dev)
    if stream is None:
        stream = torch.cuda.current_stream(dev)
    v = _hbe_view(state)
    pv = state.ps_view() if with_ps else None
    rc = ctx._lib.xaac_b200_esbr_dec_bypass_dev(ctx.handle, ctypes.byref(v), ctypes.byref(pv) if with_ps else None,
                                                _ptr(core) if core.dtype == torch.float32 else None,
                                                _ptr(core) if core.dtype == torch.int32 else None, _ptr(rg_par), _ptr(out),
                                                _ptr(out_r) if with_ps else None, None, 1, _ptr(err), n,
                                                ctypes.c_void_p(stream.cuda_stream))
    ctx.check(rc, "xaac_b200_esbr_dec_bypass_dev")
    return out, out_r, err
