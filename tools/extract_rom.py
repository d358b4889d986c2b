#!/usr/bin/env python
"""Extract the ISO/IEC 14496-3 constant tables the hot path consumes from the compiled reference.

The reference passes its ROM tables to every hot function as pointer arguments; a drop-in
deployment does the same through xaac_b200_set_*_rom().  For standalone use (tests, bench, GPU box — where
/root/reference does not exist) the same bytes are kept as small binary blobs under libxaac_b200/rom/.
This script regenerates them from oracle/_ref/libxaac_ref.so (built by `make ref`); it is the only producer
of those blobs.  Run:  python tools/extract_rom.py
"""
import ctypes
import hashlib
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.path.join(ROOT, "oracle", "_ref", "libxaac_ref.so")
OUT = os.path.join(ROOT, "libxaac_b200", "rom")


def dump(lib, getter, nbytes, name):
    fn = getattr(lib, getter)
    fn.restype = ctypes.c_void_p
    total = ctypes.c_int(0)
    p = fn(ctypes.byref(total))
    assert total.value >= nbytes, (getter, total.value, nbytes)
    data = ctypes.string_at(p, nbytes)
    path = os.path.join(OUT, name)
    with open(path, "wb") as f:
        f.write(data)
    print(f"{name}: {nbytes} bytes sha256={hashlib.sha256(data).hexdigest()[:16]}")


def main():
    lib = ctypes.CDLL(REF)
    os.makedirs(OUT, exist_ok=True)
    # leading part of ia_aac_dec_imdct_tables_struct (decoder/ixheaacd_aac_rom.h:112-121)
    dump(lib, "ref_rom_imdct_tables", 7500, "imdct_rom.bin")
    for getter, n, name in EXTRA:
        dump(lib, getter, n, name)


# leading part of ia_qmf_dec_tables_struct up to and including qmf_c (decoder/ixheaacd_sbr_rom.h:71-95)
# ia_env_calc_tables_struct (decoder/ixheaacd_sbr_rom.h:59-68) and the leading part of ixheaacd_misc_tables up to and
# including sqrt_table (decoder/ixheaacd_common_rom.h:27-37)
EXTRA = [("ref_rom_qmf_tables", 3464, "qmf_rom.bin"), ("ref_rom_env_tables", 2404, "env_rom.bin"),
         ("ref_rom_misc_tables", 2470, "misc_rom.bin"),
         # leading part of ia_ps_tables_struct through p8_13 (decoder/ixheaacd_sbr_rom.h:177-203)
         ("ref_rom_ps_tables", 1230, "ps_rom.bin"),
         # USAC frequency-domain core transform: FFT twiddles, pre / post twiddles (512, 64), sine / KBD windows (1024, 128)
         # concatenated by ref_rom_usac_tables (oracle/ref_shim_usac.c; layout XAAC_UROM_* in include/xaac_b200.h)
         ("ref_rom_usac_tables", 15880, "usac_rom.bin"),
         # eSBR banks: esbr_qmf_c[1280], esbr_w_32[60], esbr_sin_cos_twiddle_l64[64], esbr_alt_sin_twiddle_l64[32], esbr_w_16[24],
         # esbr_sin_cos_twiddle_l32[32], esbr_alt_sin_twiddle_l32[16], esbr_t_cos_sin_l32[64]
         # of ia_qmf_dec_tables_struct (decoder/ixheaacd_sbr_rom.h:96-105), concatenated by ref_rom_esbr_tables
         ("ref_rom_esbr_tables", 6288, "esbr_rom.bin"),
         # QMF harmonic transposer: the reference's global float tables (common/ixheaac_esbr_rom.c) concatenated by
         # ref_rom_hbe_tables (oracle/ref_shim_hbe.c; layout XAAC_HROM_* in include/xaac_b200.h)
         ("ref_rom_hbe_tables", 9324 * 4, "hbe_rom.bin"),
         # ref_rom_fps_tables (oracle/ref_shim_fps.c; layout XAAC_FPSROM_*)
         ("ref_rom_fps_tables", 1016 * 4, "fps_rom.bin"),
         # leading part of ia_aac_dec_block_tables_struct (decoder/ixheaacd_aac_rom.h:25-31) through scale_mant_tab
         ("ref_rom_block_tables", 620, "block_rom.bin")]

if __name__ == "__main__":
    main()
