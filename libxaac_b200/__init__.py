"""libxaac_b200 — B200-native (sm_100a) kernels for the decode-side DSP hot path of libxaac.

Host-side mirror of the reference's stage interface over the C-ABI in include/xaac_b200.h.
PyTorch is used only for device memory and streams.
"""
from ._lib import XaacB200Error, load, LIB_PATH  # noqa: F401
from .context import Context  # noqa: F401
from .imdct import (  # noqa: F401
    ONLY_LONG_SEQUENCE,
    LONG_START_SEQUENCE,
    EIGHT_SHORT_SEQUENCE,
    LONG_STOP_SEQUENCE,
    ImdctBatch,
    ImdctHostState,
    imdct_process,
    imdct_process_host,
    imdct_out_to_pcm16,
)
from .qmf import (  # noqa: F401
    QmfAnalBatch,
    QmfSynthBatch,
    QmfSynthHostState,
    cplx_anal_qmffilt,
    cplx_synt_qmffilt,
    cplx_synt_qmffilt_host,
    synth_params,
)
from .output import PeakLimiterBatch, peak_limiter_process, peak_limiter_reset_state  # noqa: F401
from .esbr import (EsbrAnalBatch, EsbrDecBatch, EsbrDecHbeBatch, EsbrHbeBatch, EsbrSynthBatch, esbr_analysis_filt_block, esbr_dec,  # noqa: F401
                   esbr_dec_hbe, esbr_dec_ps, esbr_apply_ps, EsbrDecPsBatch, esbr_dec_front, esbr_dec_back, esbr_dec_bypass, esbr_env_calc, esbr_generate_hf, esbr_qmf_hbe_apply, esbr_synthesis_filt)
from .spectral import aac_channel_pair_process  # noqa: F401
from .sideinfo import PSD_WORDS, SD_WORDS, dec_sbrdata, decode_ps_data  # noqa: F401
from .usac import STOP_START_SEQUENCE, UsacFdBatch, usac_fd_frm_dec  # noqa: F401
from .sbr import SbrState, calc_sbrenvelope, heaac_frame_host, heaac_lp_frame_host, hf_generator, sbr_dec, sbr_dec_lp, sbr_dec_lp_w32, sbr_dec_w32  # noqa: F401

__all__ = [
    "aac_channel_pair_process",
    "dec_sbrdata",
    "decode_ps_data",
    "hf_generator",
    "calc_sbrenvelope",
    "SbrState",
    "sbr_dec",
    "sbr_dec_w32",
    "sbr_dec_lp_w32",
    "sbr_dec_lp",
    "PeakLimiterBatch",
    "peak_limiter_process",
    "peak_limiter_reset_state",
    "EsbrSynthBatch",
    "EsbrAnalBatch",
    "esbr_analysis_filt_block",
    "esbr_dec",
    "esbr_env_calc",
    "esbr_generate_hf",
    "esbr_qmf_hbe_apply",
    "esbr_dec_hbe",
    "esbr_dec_ps",
    "esbr_apply_ps",
    "EsbrDecPsBatch",
    "esbr_dec_front",
    "esbr_dec_back",
    "esbr_dec_bypass",
    "EsbrDecHbeBatch",
    "EsbrHbeBatch",
    "esbr_synthesis_filt",
    "UsacFdBatch",
    "usac_fd_frm_dec",
    "heaac_frame_host",
    "heaac_lp_frame_host",
    "QmfAnalBatch",
    "cplx_anal_qmffilt",
    "QmfSynthBatch",
    "QmfSynthHostState",
    "cplx_synt_qmffilt",
    "cplx_synt_qmffilt_host",
    "synth_params",
    "Context",
    "ImdctBatch",
    "ImdctHostState",
    "imdct_process",
    "imdct_process_host",
    "imdct_out_to_pcm16",
    "XaacB200Error",
    "load",
]
