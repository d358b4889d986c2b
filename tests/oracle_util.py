"""Loaders for the CPU oracle (oracle/liboracle.so) and the compiled reference (oracle/_ref/libxaac_ref.so).
Test infrastructure only — the product package never imports this module."""
import ctypes
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_SO = os.path.join(ROOT, "oracle", "liboracle.so")
REF_SO = os.path.join(ROOT, "oracle", "_ref", "libxaac_ref.so")
ROM_DIR = os.path.join(ROOT, "libxaac_b200", "rom")


def P(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def rom(name="imdct_rom.bin"):
    return np.fromfile(os.path.join(ROM_DIR, name), dtype=np.uint8)


def build_oracle():
    src_dir = os.path.join(ROOT, "oracle", "src")
    newest = max(os.path.getmtime(os.path.join(src_dir, f)) for f in os.listdir(src_dir))
    if not os.path.exists(ORACLE_SO) or os.path.getmtime(ORACLE_SO) < newest:
        subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "oracle"), "oracle"])


class Oracle:
    """Our plain-C restatement of the hot path."""

    def __init__(self):
        build_oracle()
        self.lib = ctypes.CDLL(ORACLE_SO)
        self.rom = rom()

    def imdct_process(self, spec, ovl, prev_shape, prev_seq, win_seq, win_shape, ch_fac=1):
        """Single unit. Returns (out[1024], ovl_out[512], prev_shape', prev_seq', qshift_adj)."""
        sp = np.ascontiguousarray(spec, np.int32).copy()
        ov = np.ascontiguousarray(ovl, np.int32).copy()
        out = np.zeros(1024 * ch_fac, np.int32)
        ps, pq = ctypes.c_int32(int(prev_shape)), ctypes.c_int32(int(prev_seq))
        adj = self.lib.xo_imdct_process(P(self.rom), P(sp), P(ov), ctypes.byref(ps), ctypes.byref(pq),
                                        int(win_seq), int(win_shape), P(out), int(ch_fac))
        return out[::ch_fac].copy(), ov, ps.value, pq.value, adj

    def imdct_batch(self, spec, ovl, wstate, ics):
        """spec [n,1024], ovl [n,512], wstate [n,2]=(shape,seq) prev, ics [n,2]=(seq,shape).
        Returns (out, ovl', wstate', qshift_adj)."""
        n = spec.shape[0]
        sp = np.ascontiguousarray(spec, np.int32).copy()
        ov = np.ascontiguousarray(ovl, np.int32).copy()
        ps = np.ascontiguousarray(wstate[:, 0], np.int32).copy()
        pq = np.ascontiguousarray(wstate[:, 1], np.int32).copy()
        ws = np.ascontiguousarray(ics[:, 0], np.int32).copy()
        wh = np.ascontiguousarray(ics[:, 1], np.int32).copy()
        out = np.zeros((n, 1024), np.int32)
        adj = np.zeros(n, np.int32)
        self.lib.xo_imdct_process_batch(P(self.rom), P(sp), P(ov), P(ps), P(pq), P(ws), P(wh), P(out), P(adj), n)
        return out, ov, np.stack([ps, pq], 1).astype(np.uint8), adj.astype(np.int8)


class Ref:
    """The unmodified reference, compiled from /root/reference by oracle/Makefile (target ref)."""

    def __init__(self):
        self.lib = ctypes.CDLL(REF_SO)

    @staticmethod
    def try_load():
        if not os.path.exists(REF_SO):
            return None
        return Ref()

    def rom_imdct(self, nbytes=7500):
        fn = self.lib.ref_rom_imdct_tables
        fn.restype = ctypes.c_void_p
        total = ctypes.c_int(0)
        p = fn(ctypes.byref(total))
        return np.frombuffer(ctypes.string_at(p, nbytes), dtype=np.uint8).copy()

    def imdct_process(self, spec, ovl, prev_shape, prev_seq, win_seq, win_shape, ch_fac=1):
        sp = np.ascontiguousarray(spec, np.int32).copy()
        ov = np.ascontiguousarray(ovl, np.int32).copy()
        out = np.zeros(1024 * ch_fac, np.int32)
        ps, pq = ctypes.c_int32(int(prev_shape)), ctypes.c_int32(int(prev_seq))
        adj = self.lib.ref_imdct_process(P(sp), P(ov), ctypes.byref(ps), ctypes.byref(pq), int(win_seq),
                                         int(win_shape), P(out), int(ch_fac))
        return out[::ch_fac].copy(), ov, ps.value, pq.value, adj


# ---- shared synthetic-input generators (seeded) -------------------------------------------
def synth_units(n, seed, seq_mix=True):
    """Random spectra with per-unit magnitude 2^12..2^27 plus fixed corner units, random overlap state and a
    window-sequence mix. Returns spec, ovl, wstate(prev shape,seq), ics(seq,shape)."""
    rng = np.random.default_rng(seed)
    s = rng.integers(12, 28, size=n)
    spec = (rng.random((n, 1024)) * 2.0 - 1.0) * (2.0 ** s)[:, None]
    spec = spec.astype(np.int64).astype(np.int32)
    so = rng.integers(0, 16, size=n)
    ovl = ((rng.random((n, 512)) * 2.0 - 1.0) * (2.0 ** so)[:, None]).astype(np.int64).astype(np.int32)
    if n >= 8:  # corner units
        spec[0] = 0
        spec[1] = np.where(np.arange(1024) % 2 == 0, 2 ** 31 - 1, -(2 ** 31)).astype(np.int64).astype(np.int32)
        spec[2] = 0
        spec[2, 17] = 2 ** 31 - 1
        spec[3] = 1 << 20
        spec[4] = -1
        ovl[5] = rng.integers(-2 ** 31, 2 ** 31, 512, dtype=np.int64).astype(np.int32)
        spec[6] = rng.integers(-2 ** 31, 2 ** 31, 1024, dtype=np.int64).astype(np.int32)
        spec[7] = rng.integers(-3, 4, 1024).astype(np.int32)
    wstate = np.zeros((n, 2), np.uint8)
    ics = np.zeros((n, 2), np.uint8)
    if seq_mix:
        wstate[:, 0] = rng.integers(0, 2, n)
        wstate[:, 1] = rng.integers(0, 4, n)
        ics[:, 0] = rng.integers(0, 4, n)
        ics[:, 1] = rng.integers(0, 2, n)
    return spec, ovl, wstate, ics
