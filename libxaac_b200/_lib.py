"""ctypes binding of libxaac_b200.so (C-ABI declared in include/xaac_b200.h).

There is deliberately no fallback: if the CUDA library is missing or no device is usable, importing the
product API raises.  The CPU oracle under oracle/ is test infrastructure and is never touched from here.
"""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# XAAC_B200_LIB: another build of the same library (kernel A/B measurements); default is the in-tree one
LIB_PATH = os.environ.get("XAAC_B200_LIB") or os.path.join(_HERE, "libxaac_b200.so")
ROM_DIR = os.path.join(_HERE, "rom")

FATAL = -0x80000000

_c = ctypes
_vp, _i32, _i64, _sz = _c.c_void_p, _c.c_int32, _c.c_int64, _c.c_size_t

# name -> (restype, argtypes); kept in sync with include/xaac_b200.h (tests/test_abi.py checks both ways)
SIGNATURES = {
    "xaac_b200_create": (_i32, [_c.POINTER(_vp), _i32]),
    "xaac_b200_destroy": (None, [_vp]),
    "xaac_b200_last_error": (_c.c_char_p, [_vp]),
    "xaac_b200_num_sms": (_i32, [_vp]),
    "xaac_b200_launch_count": (_i64, [_vp]),
    "xaac_b200_sync": (_i32, [_vp]),
    "xaac_b200_kernel_timing": (_i32, [_vp, _i32]),
    "xaac_b200_kernel_times": (_i32, [_vp, _vp, _sz]),
    "xaac_b200_set_imdct_rom": (_i32, [_vp, _vp, _sz]),
    "xaac_b200_imdct_process_dev": (_i32, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _i64, _i32, _vp]),
    "xaac_b200_imdct_state_create": (_i32, [_vp, _i64, _c.POINTER(_vp)]),
    "xaac_b200_imdct_state_destroy": (None, [_vp, _vp]),
    "xaac_b200_imdct_state_upload": (_i32, [_vp, _vp, _vp, _vp]),
    "xaac_b200_imdct_state_download": (_i32, [_vp, _vp, _vp, _vp]),
    "xaac_b200_imdct_process_host": (_i32, [_vp, _vp, _vp, _vp, _vp, _vp, _i32]),
    "xaac_b200_set_qmf_rom": (_i32, [_vp, _vp, _sz]),
    "xaac_b200_qmf_synth_hq_dev": (_i32, [_vp, _vp, _vp, _vp, _vp, _vp, _i64, _i32, _vp]),
    "xaac_b200_qmf_synth_state_create": (_i32, [_vp, _i64, _c.POINTER(_vp)]),
    "xaac_b200_qmf_synth_state_destroy": (None, [_vp, _vp]),
    "xaac_b200_qmf_synth_state_upload": (_i32, [_vp, _vp, _vp, _vp]),
    "xaac_b200_qmf_synth_state_download": (_i32, [_vp, _vp, _vp, _vp]),
    "xaac_b200_qmf_synth_hq_host": (_i32, [_vp, _vp, _vp, _vp, _vp, _i32]),
    "xaac_b200_qmf_anal_hq_dev": (_i32, [_vp, _vp, _vp, _vp, _vp, _vp, _i64, _i32, _vp]),
    "xaac_b200_hf_generator_hq_dev": (_i32, [_vp, _vp, _vp, _vp, _vp, _vp, _i64, _vp]),
    "xaac_b200_set_env_rom": (_i32, [_vp, _vp, _sz, _vp, _sz]),
    "xaac_b200_calc_sbrenvelope_hq_dev": (_i32, [_vp, _vp, _vp, _vp, _vp, _vp, _i64, _vp]),
    "xaac_b200_set_ps_rom": (_i32, [_vp, _vp, _sz]),
    "xaac_b200_sbr_state_create": (_i32, [_vp, _i64, _i32, _c.POINTER(_vp)]),
    "xaac_b200_sbr_state_destroy": (None, [_vp, _vp]),
    "xaac_b200_sbr_state_upload": (_i32, [_vp, _vp, _vp, _vp]),
    "xaac_b200_sbr_state_download": (_i32, [_vp, _vp, _vp, _vp]),
    "xaac_b200_sbr_dec_hq_dev": (_i32, [_vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "xaac_b200_sbr_dec_hq_w32_dev": (_i32, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "xaac_b200_sbr_dec_lp_dev": (_i32, [_vp, _vp, _vp, _vp, _vp, _i32, _vp, _vp]),
    "xaac_b200_sbr_dec_lp_w32_dev": (_i32, [_vp, _vp, _vp, _vp, _vp, _vp, _i32, _vp, _vp]),
    "xaac_b200_heaac_frame_host": (_i32, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "xaac_b200_heaac_lp_frame_host": (_i32, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _i32, _vp]),
    "xaac_b200_peak_limiter_state_init": (_i32, [_vp, _i32, _i32]),
    "xaac_b200_peak_limiter_dev": (_i32, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _i64, _i32, _vp]),
    "xaac_b200_set_usac_rom": (_i32, [_vp, _vp, _sz]),
    "xaac_b200_usac_fd_frm_dec_dev": (_i32, [_vp, _vp, _vp, _vp, _vp, _vp, _i64, _vp]),
    "xaac_b200_set_esbr_rom": (_i32, [_vp, _vp, _sz]),
    "xaac_b200_esbr_synth64_dev": (_i32, [_vp, _vp, _vp, _vp, _vp, _vp, _i64, _vp]),
    "xaac_b200_esbr_anal32_dev": (_i32, [_vp, _vp, _vp, _vp, _vp, _vp, _i64, _vp]),
    "xaac_b200_esbr_synth64_pcm16_dev": (_i32, [_vp, _vp, _vp, _vp, _vp, _vp, _i32, _vp, _i64, _vp]),
    "xaac_b200_esbr_anal32_core_dev": (_i32, [_vp, _vp, _vp, _vp, _vp, _vp, _i64, _vp]),
    "xaac_b200_esbr_anal32_pcm16_dev": (_i32, [_vp, _vp, _i32, _vp, _vp, _vp, _vp, _i64, _vp]),
    "xaac_b200_esbr_dec_dev": (_i32, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i32, _vp, _i64, _vp]),
    "xaac_b200_esbr_generate_hf_dev": (_i32, [_vp] * 11 + [_i64, _vp]),
    "xaac_b200_set_esbr_envcalc_rom": (_i32, [_vp, _vp, _sz]),
    "xaac_b200_esbr_env_calc_dev": (_i32, [_vp] * 7 + [_i64, _vp]),
    "xaac_b200_esbr_env_calc_tes_dev": (_i32, [_vp] * 5 + [_i32] + [_vp] * 4 + [_i64, _vp]),
    "xaac_b200_dev_alloc": (_i32, [_vp, _sz, _vp]),
    "xaac_b200_dev_free": (_i32, [_vp, _vp]),
    "xaac_b200_h2d": (_i32, [_vp, _vp, _vp, _sz]),
    "xaac_b200_d2h": (_i32, [_vp, _vp, _vp, _sz]),
    "xaac_b200_dev_memset": (_i32, [_vp, _vp, _i32, _sz]),
    "xaac_b200_ipc_export": (_i32, [_vp, _vp, _vp]),
    "xaac_b200_ipc_import": (_i32, [_vp, _vp, _vp]),
    "xaac_b200_ipc_close": (_i32, [_vp, _vp]),
    "xaac_b200_set_hbe_rom": (_i32, [_vp, _vp, _sz]),
    "xaac_b200_esbr_hbe_apply_dev": (_i32, [_vp] * 8 + [_i64, _vp]),
    "xaac_b200_esbr_dec_hbe_dev": (_i32, [_vp] * 11 + [_i32, _vp, _i64, _vp]),
    "xaac_b200_set_fps_rom": (_i32, [_vp, _vp, _sz]),
    "xaac_b200_set_block_rom": (_i32, [_vp, _vp, _sz]),
    "xaac_b200_aac_spectral_dev": (_i32, [_vp] * 5 + [_i64, _vp]),
    "xaac_b200_dec_sbrdata_dev": (_i32, [_vp, _vp, _i64, _vp]),
    "xaac_b200_decode_ps_data_dev": (_i32, [_vp, _vp, _i64, _vp]),
    "xaac_b200_esbr_ps_apply_dev": (_i32, [_vp] * 3 + [_i32] + [_vp] * 8 + [_i64, _vp]),
    "xaac_b200_esbr_dec_ps_dev": (_i32, [_vp] * 14 + [_i64, _vp]),
    "xaac_b200_esbr_dec_front_dev": (_i32, [_vp] * 7 + [_i64, _vp]),
    "xaac_b200_esbr_dec_back_dev": (_i32, [_vp] * 10 + [_i32, _vp, _i64, _vp]),
    "xaac_b200_esbr_dec_bypass_dev": (_i32, [_vp] * 9 + [_i32, _vp, _i64, _vp]),
    "xaac_b200_imdct_out_to_pcm16_dev": (_i32, [_vp, _vp, _vp, _vp, _i64, _i32, _vp]),
}

_lib = None


class XaacB200Error(RuntimeError):
    pass


def load():
    """Load the shared library once; raise loudly when it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise XaacB200Error(
            f"{LIB_PATH} not found: build it with `make lib` (or __graft_entry__.build()). "
            "libxaac_b200 has no CPU fallback."
        )
    lib = ctypes.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def rom_blob(name):
    with open(os.path.join(ROM_DIR, name), "rb") as f:
        return f.read()
