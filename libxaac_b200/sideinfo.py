"""Host-side mirror of the SBR side-info dequantisation (ixheaacd_dec_sbrdata, decoder/ixheaacd_env_dec.c:628): delta decoding and
dequantisation of the envelope and noise-floor data of one or two SBR channels per element, batched over elements."""
import ctypes

import torch

from .sbr import _chk, _ptr

SD_WORDS = 1304  # XAAC_SD_WORDS


def dec_sbrdata(ctx, records, stream=None):
    """Batched drop-in for ixheaacd_dec_sbrdata on the fixed-point path (usac_flag = enh_sbr = 0, ec_flag = 0).
    records int16 [n, 1304] (XAAC_SD_* layout, include/xaac_b200.h), rewritten in place: the delta-coded Huffman indices of
    int_env_sf_arr / int_noise_floor become the (mantissa | exponent) words the envelope adjuster reads, sfb_nrg_prev and
    prev_noise_level carry to the next frame, word XAAC_SD_ERR holds what the reference function would have returned."""
    n = int(records.shape[0])
    _chk(records, torch.int16, (n, SD_WORDS), "records", "cuda")
    if stream is None:
        stream = torch.cuda.current_stream(records.device)
    rc = ctx._lib.xaac_b200_dec_sbrdata_dev(ctx.handle, _ptr(records), n, ctypes.c_void_p(stream.cuda_stream))
    ctx.check(rc, "xaac_b200_dec_sbrdata_dev")
    return records


PSD_WORDS = 576  # XAAC_PSD_WORDS


def decode_ps_data(ctx, records, stream=None):
    """Batched drop-in for ixheaacd_decode_ps_data (decoder/ixheaacd_ps_bitdec.c:98).  records int16 [n, 576] (XAAC_PSD_* layout),
    rewritten in place: IID / ICC indices delta-decoded and clamped, envelope borders fixed up, 34-band sets mapped to 20."""
    n = int(records.shape[0])
    _chk(records, torch.int16, (n, PSD_WORDS), "records", "cuda")
    if stream is None:
        stream = torch.cuda.current_stream(records.device)
    rc = ctx._lib.xaac_b200_decode_ps_data_dev(ctx.handle, _ptr(records), n, ctypes.c_void_p(stream.cuda_stream))
    ctx.check(rc, "xaac_b200_decode_ps_data_dev")
    return records
