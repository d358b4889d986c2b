"""Host-side mirror of the AAC pre-IMDCT spectral stage (ixheaacd_channel_pair_process, decoder/ixheaacd_channel.c:602-718):
M/S stereo, intensity stereo and TNS on the dequantised spectrum of AAC-LC elements, batched over elements."""
import ctypes

import torch

from .sbr import _chk, _ptr

SPS_BYTES = 3712


def aac_channel_pair_process(ctx, spec, side, pns_seed=None, err=None, stream=None):
    """Batched drop-in for ixheaacd_channel_pair_process (AAC-LC, 1 or 2 channels per element, frame length 1024).
    spec int32 [n, 2, 1024] (ptr_spec_coeff of LEFT / RIGHT, in place), side uint8 [n, 3712] (XAAC_SPS_* record), pns_seed int32
    [n] (current_seed of each element's stream, advanced in place; None: elements that use PNS are refused).
    Returns err int32 [n]: 0, or -2 for an element outside the supported subset (malformed side info) — left untouched."""
    n = int(spec.shape[0])
    _chk(spec, torch.int32, (n, 2, 1024), "spec", "cuda")
    _chk(side, torch.uint8, (n, SPS_BYTES), "side", "cuda")
    if pns_seed is not None:
        _chk(pns_seed, torch.int32, (n,), "pns_seed", "cuda")
    if err is None:
        err = torch.empty((n,), dtype=torch.int32, device=spec.device)
    else:
        _chk(err, torch.int32, (n,), "err", "cuda")
    if stream is None:
        stream = torch.cuda.current_stream(spec.device)
    rc = ctx._lib.xaac_b200_aac_spectral_dev(ctx.handle, _ptr(spec), _ptr(side), _ptr(pns_seed) if pns_seed is not None else None, _ptr(err), n, ctypes.c_void_p(stream.cuda_stream))
    ctx.check(rc, "xaac_b200_aac_spectral_dev")
    return err
