"""Host-side mirror of the reference's fixed-point SBR middle stages, batched.

hf_generator  <- ixheaacd_hf_generator(ia_sbr_hf_generator_struct*, ia_sbr_scale_fact_struct*, WORD32 **qmf_real,
                 WORD32 **qmf_imag, time_step, first_slot_offset, last_slot_offset, num_if_bands,
                 max_qmf_subband_aac, sbr_invf_mode, sbr_invf_mode_prev, ...)   (decoder/ixheaacd_lpp_tran.c:956)
"""
import ctypes

import torch

from .imdct import _chk, _ptr

HF_PARAM_WORDS = 80


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
