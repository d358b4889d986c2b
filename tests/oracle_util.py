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


    # ---- QMF banks -------------------------------------------------------------------------------------
    @property
    def qrom(self):
        if not hasattr(self, "_qrom"):
            self._qrom = rom("qmf_rom.bin")
        return self._qrom

    def cos_sin_mod(self, subband, no_channels):
        sb = np.ascontiguousarray(subband, np.int32).copy()
        self.lib.xo_cos_sin_mod(P(self.qrom), P(sb), int(no_channels))
        return sb

    def synth_batch(self, matrix, filter_states, pos, params):
        """matrix [n,32,128] i32, filter_states [n,1280] i16, pos [n,2] i16, params [n,8] i16.
        Returns (pcm [n,2048] i16, filter_states', pos')."""
        n = matrix.shape[0]
        m = np.ascontiguousarray(matrix, np.int32).copy()
        fs = np.ascontiguousarray(filter_states, np.int16).copy()
        off = np.ascontiguousarray(pos[:, 0], np.int32).copy()
        fp = np.ascontiguousarray(pos[:, 1], np.int32).copy()
        sf = np.ascontiguousarray(params[:, 0:4], np.int32).copy()
        lsb = np.ascontiguousarray(params[:, 4], np.int32).copy()
        usb = np.ascontiguousarray(params[:, 5], np.int32).copy()
        assert (params[:, 6] == 6).all()
        pcm = np.zeros((n, 2048), np.int16)
        self.lib.xo_synt_qmffilt_hq_batch(P(self.qrom), P(m), P(fs), P(off), P(fp), P(sf), P(lsb), P(usb), P(pcm), n)
        return pcm, fs, np.stack([off, fp], 1).astype(np.int16)

    def anal_batch(self, time_in, states, pos, usb):
        """time_in [n,1024] i16, states [n,320] i16, pos [n,2] i16 {core_samples offset, filter_pos}, usb [n].
        Returns (matrix [n,32,128] i32, states', pos')."""
        n = time_in.shape[0]
        t = np.ascontiguousarray(time_in, np.int16)
        st = np.ascontiguousarray(states, np.int16).copy()
        po = np.ascontiguousarray(pos[:, 0], np.int32).copy()
        fp = np.ascontiguousarray(pos[:, 1], np.int32).copy()
        ub = np.ascontiguousarray(usb, np.int32).copy()
        m = np.zeros((n, 32, 128), np.int32)
        self.lib.xo_anal_qmffilt_hq_batch(P(self.qrom), P(t), P(st), P(po), P(fp), P(ub), P(m), n)
        return m, st, np.stack([po, fp], 1).astype(np.int16)


    # ---- HF generator ------------------------------------------------------------------------------------
    def hfgen_batch(self, lpc, matrix, prm, bw_prev):
        """lpc [n,2,128] i32, matrix [n,38,128] i32, prm [n,80] i16, bw_prev [n,6] i32.
        Returns (matrix', bw_prev', hb_scale [n] i16)."""
        n = matrix.shape[0]
        l = np.ascontiguousarray(lpc, np.int32)
        m = np.ascontiguousarray(matrix, np.int32).copy()
        pr = np.ascontiguousarray(prm, np.int16)
        bw = np.ascontiguousarray(bw_prev, np.int32).copy()
        hb = np.zeros(n, np.int32)
        self.lib.xo_hf_generator_hq_batch(P(l), P(m), P(pr), P(bw), P(hb), n)
        return m, bw, hb.astype(np.int16)


    # ---- envelope adjuster -------------------------------------------------------------------------------
    @property
    def erom(self):
        if not hasattr(self, "_erom"):
            self._erom = rom("env_rom.bin")
            self._mrom = rom("misc_rom.bin")
        return self._erom

    def envcalc_batch(self, prm, sf, state, matrix):
        """prm [n,656] i16, sf [n,8] i16, state [n,232] i16, matrix [n,38,128] i32.
        Returns (matrix', sf', state', err [n] i32)."""
        n = matrix.shape[0]
        m = np.ascontiguousarray(matrix, np.int32).copy()
        s = np.ascontiguousarray(sf, np.int16).copy()
        st = np.ascontiguousarray(state, np.int16).copy()
        err = np.zeros(n, np.int32)
        er = self.erom
        self.lib.xo_calc_sbrenvelope_hq_batch(P(er), P(self._mrom), P(np.ascontiguousarray(prm, np.int16)), P(s), P(st),
                                              P(m), P(err), n)
        return m, s, st, err


    # ---- whole SBR stage / PS ------------------------------------------------------------------------------
    @property
    def psrom(self):
        if not hasattr(self, "_psrom"):
            self._psrom = rom("ps_rom.bin")
        return self._psrom

    def sbr_dec(self, side, st, ps, time_in):
        """single unit: side [1232] i16, st [3920] i16, ps [3888] i16, time_in [1024] i16.
        Returns (st', ps', out_l [2048], out_r [2048], err)."""
        s = np.ascontiguousarray(st, np.int16).copy()
        p = np.ascontiguousarray(ps, np.int16).copy()
        ol = np.zeros(2048, np.int16)
        orr = np.zeros(2048, np.int16)
        scratch = np.zeros(40 * 128, np.int32)
        self.erom
        err = self.lib.xo_sbr_dec_hq(P(self.qrom), P(self._erom), P(self._mrom), P(self.psrom),
                                     P(np.ascontiguousarray(side, np.int16)), P(s), P(p),
                                     P(np.ascontiguousarray(time_in, np.int16)), 1, P(ol), P(orr), 1, P(scratch))
        return s, p, ol, orr, err

    def sbr_dec_batch(self, side, st, ps, time_in):
        outs = [self.sbr_dec(side[u], st[u], ps[u], time_in[u]) for u in range(len(side))]
        return tuple(np.stack([o[k] for o in outs]) for k in range(4)) + (np.array([o[4] for o in outs], np.int32),)

    def sbr_dec_lp(self, side, st, time_in):
        """low-power stage, single unit. Returns (st', out [2048], err)."""
        s = np.ascontiguousarray(st, np.int16).copy()
        out = np.zeros(2048, np.int16)
        scratch = np.zeros(40 * 64, np.int32)
        self.erom
        err = self.lib.xo_sbr_dec_lp(P(self.qrom), P(self._erom), P(self._mrom), P(np.ascontiguousarray(side, np.int16)),
                                     P(s), P(np.ascontiguousarray(time_in, np.int16)), 1, P(out), 1, P(scratch))
        return s, out, err

    # ---- USAC frequency-domain core transform ------------------------------------------------------------------
    @property
    def urom(self):
        if not hasattr(self, "_urom"):
            self._urom = rom("usac_rom.bin")
        return self._urom

    def usac_complex_fft(self, xr, xi, preshift):
        a = np.ascontiguousarray(xr, np.int32).copy()
        b = np.ascontiguousarray(xi, np.int32).copy()
        ps = self.lib.xo_usac_complex_fft(P(self.urom), P(a), P(b), len(a), int(preshift))
        return a, b, ps

    def usac_fd_batch(self, coef, ov, win_seq, win_shape, win_shape_prev):
        """coef [n,1024], ov [n,1024], sequences / shapes [n].  Returns (out [n,1024], ov', err [n])."""
        n = coef.shape[0]
        c = np.ascontiguousarray(coef, np.int32).copy()
        o = np.ascontiguousarray(ov, np.int32).copy()
        out = np.zeros((n, 1024), np.int32)
        err = np.zeros(n, np.int32)
        self.lib.xo_usac_fd_frm_dec_batch(P(self.urom), P(c), P(o), P(np.ascontiguousarray(win_seq, np.int32)),
                                          P(np.ascontiguousarray(win_shape, np.int32)),
                                          P(np.ascontiguousarray(win_shape_prev, np.int32)), P(out), P(err), n)
        return out, o, err

    @property
    def esrom(self):
        if not hasattr(self, "_esrom"):
            self._esrom = rom("esbr_rom.bin")
        return self._esrom

    def esbr_synth_batch(self, qmf, fs, pos):
        """qmf float32 [n,32,128], fs int32 [n,1280], pos int32 [n,2].  Returns (out float32 [n,2048], fs', pos')."""
        q = np.ascontiguousarray(qmf, np.float32)
        f = np.ascontiguousarray(fs, np.int32).copy()
        p = np.ascontiguousarray(pos, np.int32).copy()
        out = np.zeros((q.shape[0], 2048), np.float32)
        self.lib.xo_esbr_synth64_batch(P(self.esrom), P(q), P(f), P(p), P(out), q.shape[0])
        return out, f, p

    def esbr_anal_batch(self, time_in, states, pos):
        """time_in float32 [n,1024], states int32 [n,320], pos int32 [n,2].  Returns (qmf float32 [n,32,128], states', pos')."""
        t = np.ascontiguousarray(time_in, np.float32)
        f = np.ascontiguousarray(states, np.int32).copy()
        p = np.ascontiguousarray(pos, np.int32).copy()
        q = np.zeros((t.shape[0], 32, 128), np.float32)
        self.lib.xo_esbr_anal32_batch(P(self.esrom), P(t), P(f), P(p), P(q), t.shape[0])
        return q, f, p

    def peak_limiter_batch(self, st, samples, qshift_adj, ch):
        """st [n,1548] int32 (XO_PL_*), samples int32 [n,1024,ch], qshift_adj int8 [n,ch].
        Returns (st', samples', pcm16, err)."""
        s = np.ascontiguousarray(st, np.int32).copy()
        x = np.ascontiguousarray(samples, np.int32).copy()
        q = np.ascontiguousarray(qshift_adj, np.int8)
        n = x.shape[0]
        pcm = np.zeros(x.shape, np.int16)
        err = np.zeros(n, np.int32)
        self.lib.xo_peak_limiter_batch(P(s), P(x), P(q), P(pcm), P(err), int(ch), n)
        return s, x, pcm, err

    def imdct_out_to_pcm16(self, samples, qshift_adj, mode):
        x = np.ascontiguousarray(samples, np.int32)
        q = np.ascontiguousarray(qshift_adj, np.int8)
        out = np.zeros(x.shape, np.int16)
        self.lib.xo_imdct_out_to_pcm16(P(x), P(q), P(out), int(x.shape[0]), int(mode))
        return out

    def ps_apply_frame(self, ps_prm, ps, m, usb, shiftdelay_late, common_shift, as_built):
        p = np.ascontiguousarray(ps, np.int16).copy()
        mm = np.ascontiguousarray(m, np.int32).copy()
        right = np.zeros((32, 128), np.int32)
        self.erom
        self.lib.xo_ps_apply_frame(P(self._erom), P(self._mrom), P(self.psrom), P(np.ascontiguousarray(ps_prm, np.int16)),
                                   P(p), P(mm), P(right), int(usb), int(shiftdelay_late), int(common_shift), int(as_built))
        return mm, right, p


class Ref:
    """The unmodified reference, compiled from /root/reference by oracle/Makefile (target ref)."""

    def __init__(self, path=REF_SO):
        self.lib = ctypes.CDLL(path)

    @staticmethod
    def try_load(path=REF_SO):
        if not os.path.exists(path):
            return None
        return Ref(path)

    def sbr_dec(self, side, st, ps, time_in):
        """ixheaacd_sbr_dec driven from the flat records; ps may be None (no PS instance)."""
        s = np.ascontiguousarray(st, np.int16).copy()
        p = None if ps is None else np.ascontiguousarray(ps, np.int16).copy()
        ol = np.zeros(2048, np.int16)
        orr = np.zeros(2048, np.int16)
        err = self.lib.ref_sbr_dec_hq(P(np.ascontiguousarray(side, np.int16)), P(s), None if p is None else P(p),
                                      P(np.ascontiguousarray(time_in, np.int16)), P(ol), P(orr))
        return s, p, ol, orr, err

    def sbr_dec_lp(self, side, st, time_in):
        s = np.ascontiguousarray(st, np.int16).copy()
        out = np.zeros(2048, np.int16)
        err = self.lib.ref_sbr_dec_lp(P(np.ascontiguousarray(side, np.int16)), P(s),
                                      P(np.ascontiguousarray(time_in, np.int16)), P(out))
        return s, out, err

    def ps_apply_frame(self, side, st, ps, sf, m, usb, common_shift):
        p = np.ascontiguousarray(ps, np.int16).copy()
        mm = np.ascontiguousarray(m, np.int32).copy()
        right = np.zeros((32, 128), np.int32)
        self.lib.ref_ps_apply_frame(P(np.ascontiguousarray(side, np.int16)), P(np.ascontiguousarray(st, np.int16)), P(p),
                                    P(np.ascontiguousarray(sf, np.int16)), P(mm), P(right), int(usb), int(common_shift))
        return mm, right, p

    def usac_complex_fft(self, xr, xi, preshift):
        a = np.ascontiguousarray(xr, np.int32).copy()
        b = np.ascontiguousarray(xi, np.int32).copy()
        ps = self.lib.ref_usac_complex_fft(P(a), P(b), len(a), int(preshift))
        return a, b, ps

    def usac_fd_batch(self, coef, ov, win_seq, win_shape, win_shape_prev):
        n = coef.shape[0]
        c = np.ascontiguousarray(coef, np.int32).copy()
        o = np.ascontiguousarray(ov, np.int32).copy()
        out = np.zeros((n, 1024), np.int32)
        err = np.zeros(n, np.int32)
        self.lib.ref_usac_fd_frm_dec_batch(P(c), P(o), P(np.ascontiguousarray(win_seq, np.int32)),
                                           P(np.ascontiguousarray(win_shape, np.int32)),
                                           P(np.ascontiguousarray(win_shape_prev, np.int32)), P(out), P(err), n)
        return out, o, err

    def esbr_synth_batch(self, qmf, fs, pos):
        q = np.ascontiguousarray(qmf, np.float32)
        f = np.ascontiguousarray(fs, np.int32).copy()
        p = np.ascontiguousarray(pos, np.int32).copy()
        out = np.zeros((q.shape[0], 2048), np.float32)
        self.lib.ref_esbr_synth64_batch(P(q), P(f), P(p), P(out), q.shape[0])
        return out, f, p

    def esbr_anal_batch(self, time_in, states, pos):
        t = np.ascontiguousarray(time_in, np.float32)
        f = np.ascontiguousarray(states, np.int32).copy()
        p = np.ascontiguousarray(pos, np.int32).copy()
        q = np.zeros((t.shape[0], 32, 128), np.float32)
        self.lib.ref_esbr_anal32_batch(P(t), P(f), P(p), P(q), t.shape[0])
        return q, f, p

    def peak_limiter_init(self, ch, sample_rate):
        st = np.zeros(PL_WORDS, np.int32)
        delay = self.lib.ref_peak_limiter_init(P(st), int(ch), int(sample_rate))
        return st, delay

    def peak_limiter_batch(self, st, samples, qshift_adj, ch):
        s = np.ascontiguousarray(st, np.int32).copy()
        x = np.ascontiguousarray(samples, np.int32).copy()
        q = np.ascontiguousarray(qshift_adj, np.int8)
        pcm = np.zeros(x.shape, np.int16)
        self.lib.ref_peak_limiter_batch(P(s), P(x), P(q), P(pcm), int(ch), x.shape[0])
        return s, x, pcm

    def rom_imdct(self, nbytes=7500):
        fn = self.lib.ref_rom_imdct_tables
        fn.restype = ctypes.c_void_p
        total = ctypes.c_int(0)
        p = fn(ctypes.byref(total))
        return np.frombuffer(ctypes.string_at(p, nbytes), dtype=np.uint8).copy()

    def hfgen(self, lpc, matrix, prm, bw_prev):
        m = np.ascontiguousarray(matrix, np.int32).copy()
        bw = np.ascontiguousarray(bw_prev, np.int32).copy()
        hb = self.lib.ref_hf_generator_hq(P(np.ascontiguousarray(lpc, np.int32)), P(m),
                                          P(np.ascontiguousarray(prm, np.int16)), P(bw))
        return m, bw, hb

    def envcalc(self, prm, sf, state, matrix):
        m = np.ascontiguousarray(matrix, np.int32).copy()
        s = np.ascontiguousarray(sf, np.int16).copy()
        st = np.ascontiguousarray(state, np.int16).copy()
        err = self.lib.ref_calc_sbrenvelope_hq(P(np.ascontiguousarray(prm, np.int16)), P(s), P(st), P(m))
        return m, s, st, err

    def rom_blob(self, getter, nbytes):
        fn = getattr(self.lib, getter)
        fn.restype = ctypes.c_void_p
        total = ctypes.c_int(0)
        p = fn(ctypes.byref(total))
        assert total.value >= nbytes
        return np.frombuffer(ctypes.string_at(p, nbytes), dtype=np.uint8).copy()

    def rom_qmf(self, nbytes=3464):
        fn = self.lib.ref_rom_qmf_tables
        fn.restype = ctypes.c_void_p
        total = ctypes.c_int(0)
        p = fn(ctypes.byref(total))
        return np.frombuffer(ctypes.string_at(p, nbytes), dtype=np.uint8).copy()

    def cos_sin_mod(self, subband, no_channels):
        sb = np.ascontiguousarray(subband, np.int32).copy()
        self.lib.ref_cos_sin_mod(P(sb), int(no_channels))
        return sb

    def dec_sbrdata_batch(self, records):
        """XAAC_SD_* records through the compiled ixheaacd_dec_sbrdata (oracle/ref_shim_sd.c); returns the rewritten records"""
        r = np.ascontiguousarray(records, np.int16).copy()
        self.lib.ref_dec_sbrdata_batch(ctypes.c_int64(r.shape[0]), P(r))
        return r

    def decode_ps_data_batch(self, records):
        """XAAC_PSD_* records through the compiled ixheaacd_decode_ps_data (oracle/ref_shim_sd.c)"""
        r = np.ascontiguousarray(records, np.int16).copy()
        self.lib.ref_decode_ps_data_batch(ctypes.c_int64(r.shape[0]), P(r))
        return r

    def synth(self, matrix, filter_states, pos, params, ch_fac=1):
        """single unit through ixheaacd_cplx_synt_qmffilt"""
        m = np.ascontiguousarray(matrix, np.int32).copy()
        fs = np.ascontiguousarray(filter_states, np.int16).copy()
        off, fp = ctypes.c_int32(int(pos[0])), ctypes.c_int32(int(pos[1]))
        sf = np.ascontiguousarray(params[0:4], np.int32).copy()
        pcm = np.zeros(2048 * ch_fac, np.int16)
        self.lib.ref_synt_qmffilt_hq(P(m), P(fs), ctypes.byref(off), ctypes.byref(fp), P(sf), int(params[4]),
                                     int(params[5]), int(params[6]), P(pcm), int(ch_fac))
        return pcm[::ch_fac].copy(), fs, np.array([off.value, fp.value], np.int16)

    def anal(self, time_in, states, pos, usb, ch_fac=1):
        st = np.ascontiguousarray(states, np.int16).copy()
        po, fp = ctypes.c_int32(int(pos[0])), ctypes.c_int32(int(pos[1]))
        m = np.zeros((32, 128), np.int32)
        t = np.ascontiguousarray(time_in, np.int16)
        lb = self.lib.ref_anal_qmffilt_hq(P(t), int(ch_fac), P(st), ctypes.byref(po), ctypes.byref(fp), int(usb), P(m))
        return m, st, np.array([po.value, fp.value], np.int16), lb

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


def synth_qmf_units(n, seed):
    """Synthetic synthesis-bank inputs: QMF matrices with per-unit magnitude, random filter state,
    ring/coefficient offsets in every phase, scale factors spanning left and right block shifts, lsb/usb spread.
    Returns matrix [n,32,128] i32, filter_states [n,1280] i16, pos [n,2] i16, params [n,8] i16."""
    rng = np.random.default_rng(seed)
    s = rng.integers(8, 30, size=(n, 1, 1))
    matrix = ((rng.random((n, 32, 128)) * 2 - 1) * (2.0 ** s)).astype(np.int64).astype(np.int32)
    fs = rng.integers(-32768, 32768, (n, 1280)).astype(np.int16)
    pos = np.stack([rng.integers(0, 10, n) * 128, rng.integers(0, 10, n) * 64], 1).astype(np.int16)
    params = np.zeros((n, 8), np.int16)
    params[:, 0] = rng.integers(-16, 8, n)   # ov_lb_scale
    params[:, 1] = rng.integers(-16, 8, n)   # lb_scale
    params[:, 2] = rng.integers(-16, 8, n)   # hb_scale
    params[:, 3] = -6                        # st_syn_scale (reference constant)
    lsb = rng.integers(0, 41, n)
    params[:, 4] = lsb
    params[:, 5] = np.minimum(64, lsb + rng.integers(0, 33, n))
    params[:, 6] = 6
    if n >= 8:
        matrix[0] = 0
        fs[0] = 0
        matrix[1] = rng.integers(-2 ** 31, 2 ** 31, (32, 128), dtype=np.int64).astype(np.int32)  # saturating everywhere
        params[1, 0:3] = (-20, -20, -20)
        matrix[2] = 2 ** 31 - 1
        matrix[3] = -(2 ** 31)
        params[4, 0:3] = (31, 31, 31)       # extreme right shifts
        params[5, 0] = params[5, 1]         # ov_lb_shift == lb_shift branch
        params[6, 4:6] = (0, 0)             # no bands scaled
        params[7, 4:6] = (64, 64)           # everything low band
    return matrix, fs, pos, params


def synth_hfgen_units(n, seed, golden_prm):
    """HF-generator inputs: parameter rows drawn from the tapped real ones (transposer settings are table-driven) with
    randomised inverse-filter modes, envelope borders and start band; random QMF matrices / LPC states with per-unit
    magnitude. Returns lpc [n,2,128], matrix [n,38,128], prm [n,80], bw_prev [n,6]."""
    rng = np.random.default_rng(seed)
    prm = golden_prm[rng.integers(0, len(golden_prm), n)].copy()
    prm[:, 54:64] = rng.integers(0, 4, (n, 10))
    prm[:, 64:74] = rng.integers(0, 4, (n, 10))
    prm[:, 52] = rng.integers(0, 3, n)
    prm[:, 53] = rng.integers(-1, 3, n)
    prm[:, 76] = prm[:, 76] + rng.integers(-3, 6, n)
    prm[:, 74] = rng.integers(-4, 24, n)
    prm[:, 75] = rng.integers(-4, 24, n)
    s = rng.integers(8, 31, size=(n, 1, 1))
    matrix = ((rng.random((n, 38, 128)) * 2 - 1) * (2.0 ** s)).astype(np.int64).astype(np.int32)
    lpc = ((rng.random((n, 2, 128)) * 2 - 1) * (2.0 ** s)).astype(np.int64).astype(np.int32)
    bw_prev = rng.integers(0, 0x7f800000, (n, 6)).astype(np.int32)
    if n >= 4:
        matrix[0] = 0
        lpc[0] = 0
        matrix[1] = rng.integers(-2 ** 31, 2 ** 31, (38, 128), dtype=np.int64).astype(np.int32)
        matrix[2, :, :] = 1 << 29
        bw_prev[3] = 0
    return lpc, matrix, prm, bw_prev


def synth_env_units(n, seed, golden):
    """Envelope-adjuster inputs: side-info rows and states drawn from the tapped real ones (frequency tables, grids and
    envelope data obey many invariants) with randomised limiter gains / interpolation / smoothing / channel mode /
    harmonics / transient position / envelope exponents / scale factors / adjuster state; random QMF matrices with
    per-unit magnitude.  Returns prm [n,656] i16, sf [n,8] i16, state [n,232] i16, matrix [n,38,128] i32."""
    rng = np.random.default_rng(seed)
    idx = rng.integers(0, len(golden["prm"]), n)
    prm = golden["prm"][idx].copy()
    sf = golden["sf_in"][idx].copy()
    st = golden["st_in"][idx].copy()
    s = rng.integers(6, 31, size=(n, 1, 1))
    matrix = ((rng.random((n, 38, 128)) * 2 - 1) * (2.0 ** s)).astype(np.int64).astype(np.int32)
    for u in range(n):
        if u % 7 == 0:
            matrix[u] = golden["m_in"][idx[u]]
        prm[u, 3] = rng.integers(0, 4)
        prm[u, 4] = rng.integers(0, 2)
        prm[u, 5] = rng.integers(0, 2)
        prm[u, 2] = rng.choice([1, 3])
        nhi = prm[u, 7]
        if u % 3 == 0:
            prm[u, 151:151 + nhi] = rng.random(nhi) < 0.3
        if u % 5 == 0:
            prm[u, 13] = rng.integers(-1, prm[u, 12] + 1)
        if u % 2 == 0:
            v = prm[u, 207:207 + 448].astype(np.int32)
            v = (v & 0xFFC0) | np.clip((v & 0x3F) + rng.integers(-6, 7, 448), 0, 63)
            prm[u, 207:207 + 448] = v.astype(np.uint16).view(np.int16)
        sf[u, 3] = rng.integers(-4, 20)
        sf[u, 4] = rng.integers(-4, 20)
        sf[u, 0] = rng.integers(-4, 20)
        if u % 4 == 0:
            st[u, 0:112:2] = rng.integers(0, 32767, 56)
            st[u, 1:112:2] = rng.integers(-10, 20, 56)
            st[u, 112:168] = rng.integers(0, 32767, 56)
            st[u, 168] = rng.integers(-5, 25)
            st[u, 169] = rng.integers(0, 2)
            st[u, 170] = rng.integers(0, 512)
            st[u, 171] = rng.integers(-1, 1)
            st[u, 172] = rng.integers(0, 4)
            st[u, 173:229] = rng.integers(0, 2, 56)
    if n >= 4:
        matrix[0] = 0
        matrix[1] = rng.integers(-2 ** 31, 2 ** 31, (38, 128), dtype=np.int64).astype(np.int32)
        matrix[2] = 2 ** 31 - 1
        matrix[3] = -(2 ** 31)
    return prm, sf, st, matrix


REF_NSA_SO = os.path.join(ROOT, "oracle", "_ref", "libxaac_ref_nsa.so")
PS_ST_DSP_WORDS = 2596  # PS state words before the right synthesis bank


def synth_sbr_units(n, seed, golden):
    """Whole-stage inputs: side info and state drawn from tapped real frames (their invariants intact), core PCM replaced
    by scaled noise / tones per unit, analysis / synthesis / overlap / LPC state perturbed.  Every third unit has
    apply_processing = 0 (upsampling only); PS units alternate between the two rotation modes.
    Returns side [n,1232], st [n,3920], ps [n,3888], tin [n,1024] (all int16)."""
    rng = np.random.default_rng(seed)
    idx = rng.integers(0, len(golden["side"]), n)
    side = golden["side"][idx].copy()
    st = golden["st_in"][idx].copy()
    ps = golden["ps_in"][idx].copy()
    tin = np.zeros((n, 1024), np.int16)
    for u in range(n):
        amp = 2.0 ** rng.integers(2, 16)
        kind = u % 4
        if kind == 0:
            x = golden["tin"][idx[u]].astype(np.float64)
        elif kind == 1:
            x = rng.standard_normal(1024) * amp / 3
        elif kind == 2:
            x = amp * np.sin(2 * np.pi * rng.uniform(0.001, 0.49) * np.arange(1024) + rng.uniform(0, 6))
        else:
            x = rng.integers(-32768, 32768, 1024).astype(np.float64)
        tin[u] = np.clip(x, -32768, 32767).astype(np.int16)
        if u % 3 == 2 and not side[u, 737]:
            side[u, 736] = 0
        if side[u, 737] and u % 2:
            side[u, 737] = 2
        if u % 5 == 1:
            st[u, 0:320] = rng.integers(-20000, 20000, 320)
            st[u, 580:1860] = rng.integers(-20000, 20000, 1280)
            sc = rng.integers(10, 30)
            ov = ((rng.random(768) * 2 - 1) * 2.0 ** sc).astype(np.int64).astype(np.int32)
            ov[384:] = 0
            st[u, 2384:3920] = ov.view(np.int16)
            st[u, 326] = rng.integers(-6, 16)      # ov_lb_scale
            st[u, 328] = rng.integers(0, 20)       # ov_hb_scale
    return side, st, ps, tin


def synth_sbr_lp_units(n, seed, golden):
    """Low-power whole-stage inputs: side info / state from tapped HE-AACv1 stereo frames, core PCM and state perturbed
    like synth_sbr_units.  Returns side [n,1232], st [n,3920], tin [n,1024]."""
    rng = np.random.default_rng(seed)
    idx = rng.integers(0, len(golden["side"]), n)
    side = golden["side"][idx].copy()
    st = golden["st_in"][idx].copy()
    tin = np.zeros((n, 1024), np.int16)
    for u in range(n):
        amp = 2.0 ** rng.integers(2, 16)
        kind = u % 4
        if kind == 0:
            x = golden["tin"][idx[u]].astype(np.float64)
        elif kind == 1:
            x = rng.standard_normal(1024) * amp / 3
        elif kind == 2:
            x = amp * np.sin(2 * np.pi * rng.uniform(0.001, 0.49) * np.arange(1024) + rng.uniform(0, 6))
        else:
            x = rng.integers(-32768, 32768, 1024).astype(np.float64)
        tin[u] = np.clip(x, -32768, 32767).astype(np.int16)
        if u % 3 == 2:
            side[u, 736] = 0
        if u % 5 == 1:
            st[u, 0:320] = rng.integers(-20000, 20000, 320)
            st[u, 580:1860] = rng.integers(-20000, 20000, 1280)
            ov = np.zeros(768, np.int32)
            ov[:384] = ((rng.random(384) * 2 - 1) * 2.0 ** rng.integers(10, 30)).astype(np.int64).astype(np.int32)
            st[u, 2384:3920] = ov.view(np.int16)
            st[u, 326] = rng.integers(-6, 16)
            st[u, 328] = rng.integers(0, 20)
        if u % 4 == 1:
            nhi = side[u, 7]
            side[u, 151:151 + nhi] = rng.random(nhi) < 0.3
            side[u, 3] = rng.integers(0, 4)
            side[u, 4] = rng.integers(0, 2)
    return side, st, tin


PL_WORDS = 1548


def peak_limiter_reset_state(ch, sample_rate):
    """state of ixheaacd_peak_limiter_init restated (decoder/ixheaacd_peak_limiter.c:45-75): attack = 5 ms, release = 50 ms"""
    st = np.zeros(PL_WORDS, np.int32)
    attack = int(np.float32(5.0) * np.float32(sample_rate) / np.float32(1000))
    f = st.view(np.float32)
    f[0] = np.float32(0.1 ** (1.0 / (attack + 1)))
    f[1] = np.float32(0.1 ** (1.0 / (float(np.float32(50.0) * np.float32(sample_rate) / np.float32(1000)) + 1)))
    f[2] = 1.0
    f[3] = 1.0
    st[4:6] = np.array([1.0], np.float64).view(np.int32)
    st[6] = attack
    st[10] = 1
    st[11] = ch
    return st


def synth_peaklim_units(n, ch, seed, loud_fraction=0.5):
    """WORD32 IMDCT-domain samples [n,1024,ch] + qshift_adj [n,ch]: quiet units (limiter at rest) and loud ones whose
    scaled peaks exceed 2^31 (limiter engages), bursts and silence"""
    rng = np.random.default_rng(seed)
    x = np.zeros((n, 1024, ch), np.int32)
    q = rng.integers(1, 3, (n, ch)).astype(np.int8)
    for u in range(n):
        loud = rng.random() < loud_fraction
        amp = 2.0 ** (rng.uniform(29.0, 30.9) if loud else rng.uniform(10, 28))
        kind = u % 4
        tt = np.arange(1024)
        if kind == 0:
            s = np.sin(2 * np.pi * rng.uniform(0.001, 0.3) * tt + rng.uniform(0, 6))[:, None] * np.ones((1, ch))
        elif kind == 1:
            s = rng.standard_normal((1024, ch)) / 3
        elif kind == 2:
            s = np.zeros((1024, ch))
            p = rng.integers(0, 900)
            s[p:p + 100] = rng.standard_normal((100, ch))
        else:
            s = rng.uniform(-1, 1, (1024, ch))
        x[u] = np.clip(s * amp, -2 ** 31, 2 ** 31 - 1).astype(np.int64).astype(np.int32)
    return x, q


def usac_seq_walk(n_units, n_frames, seed):
    """legal USAC FD window-sequence walk per unit: seq [frames, n] (0 ONLY_LONG, 1 LONG_START, 2 EIGHT_SHORT,
    3 LONG_STOP, 4 STOP_START), shape [frames, n] (0 sine, 1 KBD)"""
    rng = np.random.default_rng(seed)
    seq = np.zeros((n_frames, n_units), np.int32)
    shape = rng.integers(0, 2, (n_frames, n_units)).astype(np.int32)
    prev = np.zeros(n_units, np.int32)
    for f in range(n_frames):
        r = rng.random(n_units)
        nxt = np.zeros(n_units, np.int32)
        longish = (prev == 0) | (prev == 3)          # after ONLY_LONG / LONG_STOP: long or start
        nxt[longish] = np.where(r[longish] < 0.3, 1, 0)
        startish = (prev == 1) | (prev == 2) | (prev == 4)  # after a start-type / short frame: short, stop or stop-start
        nxt[startish] = np.where(r[startish] < 0.4, 2, np.where(r[startish] < 0.8, 3, 4))
        seq[f] = nxt
        prev = nxt
    return seq, shape


def synth_usac_units(n, seed):
    """dequantised USAC spectra [n,1024] with per-unit magnitude 2^4 .. 2^30 plus corner units, overlap [n,1024]"""
    rng = np.random.default_rng(seed)
    s = rng.integers(4, 31, (n, 1))
    coef = ((rng.random((n, 1024)) * 2 - 1) * 2.0 ** s).astype(np.int64).clip(-2 ** 31, 2 ** 31 - 1).astype(np.int32)
    ov = ((rng.random((n, 1024)) * 2 - 1) * 2.0 ** rng.integers(0, 29, (n, 1))).astype(np.int64).astype(np.int32)
    if n > 4:
        coef[0] = 0
        coef[1] = np.where(np.arange(1024) % 2 == 0, 2 ** 31 - 1, -2 ** 31)
        coef[2] = 0
        coef[2, 5] = 2 ** 31 - 1
        coef[3] = 1 << 20
        ov[0] = 0
    return coef, ov


def synth_esbr_units(n, seed):
    """float QMF matrices [n,32,128] (re 64 | im 64 per slot) with per-unit magnitude 2^-6 .. 2^22 (the float eSBR path works
    on PCM-scaled floats), filter states [n,1280] and lock-step (drc_offset, filter_pos) pairs; a few corner units"""
    rng = np.random.default_rng(seed)
    mag = 2.0 ** rng.uniform(-6, 22, (n, 1, 1))
    qmf = ((rng.random((n, 32, 128)) * 2 - 1) * mag).astype(np.float32)
    usb = rng.integers(20, 65, n)
    for u in range(n):
        qmf[u, :, usb[u]:64] = 0
        qmf[u, :, 64 + usb[u]:] = 0
    fs = ((rng.random((n, 1280)) * 2 - 1) * 2.0 ** rng.integers(8, 30, (n, 1))).astype(np.int64).astype(np.int32)
    ph = rng.integers(0, 5, n)
    pos = np.stack([(256 * ph) % 1280, (128 * ((5 - ph) % 5)) % 640], 1).astype(np.int32)
    if n > 3:
        qmf[0] = 0
        qmf[1] = 3.0e7
        fs[0] = 0
    return qmf, fs, pos


def synth_esbr_anal_units(n, seed):
    """float core samples [n,1024] in the float eSBR path's scale (+-1.0 = full scale after the 2^-15 hand-over), ring
    states [n,320] and lock-step (position, coefficient phase) pairs"""
    rng = np.random.default_rng(seed)
    amp = 2.0 ** rng.uniform(-14, 0.5, (n, 1))
    x = np.zeros((n, 1024), np.float32)
    t = np.arange(1024)
    for u in range(n):
        k = u % 3
        if k == 0:
            s = np.sin(2 * np.pi * rng.uniform(0.001, 0.45) * t + rng.uniform(0, 6))
        elif k == 1:
            s = rng.standard_normal(1024) / 3
        else:
            s = rng.uniform(-1, 1, 1024)
        x[u] = (s * amp[u]).astype(np.float32)
    st = ((rng.random((n, 320)) * 2 - 1) * 2.0 ** rng.integers(2, 16, (n, 1))).astype(np.int64).astype(np.int32)
    ph = rng.integers(0, 5, n)
    pos = np.stack([(64 * ph) % 320, (128 * ((5 - ph) % 5)) % 640], 1).astype(np.int32)
    if n > 2:
        x[0] = 0
        st[0] = 0
    return x, st, pos


# ---- eSBR float HF generator (ixheaacd_generate_hf) ---------------------------------------------------------------------
EHF_ROWS, EHF_PAR_WORDS = 40, 96
EHF = dict(NUM_MF=0, NUM_IF=1, SB_START=2, BORDER_FIRST=3, BORDER_LAST=4, HBE_FLAG=5, PATCHING_MODE=6, FS=7, PRE_PROC=8,
           USF4=9, MPS_SBR=10, COV_COUNT=11, INVF=16, INVF_PREV=21, INVF_TBL=26, FMASTER=32)


def synth_esbr_hfgen_units(n, seed, hbe=True):
    """Random but well-formed eSBR HF-generator units: master / noise band tables, inverse-filtering modes, frame borders,
    low-band QMF history (noise, tones, silent bands) and, for HBE units, a phase-vocoder buffer.  Every 16th unit carries
    a malformed noise table (the reference returns -1)."""
    rng = np.random.default_rng(seed)
    par = np.zeros((n, EHF_PAR_WORDS), np.int32)
    src = np.zeros((2, n, EHF_ROWS, 64), np.float32)
    pv = np.zeros((2, n, EHF_ROWS, 64), np.float32)
    dst = (rng.standard_normal((2, n, EHF_ROWS, 64)) * 3).astype(np.float32)   # in/out: untouched cells must survive
    bw_prev = np.zeros((n, 6), np.float32)
    t = np.arange(EHF_ROWS)[:, None]
    for u in range(n):
        lsb = int(rng.integers(6, 33))
        fm = [lsb]
        while fm[-1] < 64 and len(fm) < 57:
            w = int(rng.integers(1, 5)) if fm[-1] < 40 else int(rng.integers(2, 7))
            if fm[-1] + w > 64 or (len(fm) > 6 and rng.random() < 0.04):
                break
            fm.append(fm[-1] + w)
        if len(fm) < 3:
            fm = [lsb, min(lsb + 4, 63), 64]
        num_mf = len(fm) - 1
        usb = fm[-1]
        xi = int(rng.integers(0, min(4, num_mf)))
        sb_start = fm[xi]
        num_if = int(rng.integers(1, 6))
        cuts = sorted(set(rng.choice(np.arange(sb_start + 1, usb + 1), size=min(num_if - 1, max(usb - sb_start - 1, 0)),
                                     replace=False).tolist())) if usb - sb_start > 1 else []
        cuts = [c for c in cuts if c < usb][: num_if - 1] + [usb]
        num_if = len(cuts)
        tbl = cuts + [usb] * (5 - len(cuts))
        if u % 16 == 15:
            tbl = [sb_start] * 5
        p = par[u]
        p[EHF["NUM_MF"]], p[EHF["NUM_IF"]], p[EHF["SB_START"]] = num_mf, num_if, sb_start
        p[EHF["BORDER_FIRST"]] = int(rng.integers(0, 4))
        p[EHF["BORDER_LAST"]] = int(rng.integers(14, 20))
        p[EHF["HBE_FLAG"]] = int(hbe and rng.random() < 0.5)
        p[EHF["PATCHING_MODE"]] = int(rng.random() < 0.5)
        p[EHF["FS"]] = int(rng.choice([24000, 32000, 44100, 48000, 64000, 88200]))
        if rng.random() < 0.15:
            p[EHF["MPS_SBR"]] = 1
            p[EHF["COV_COUNT"]] = int(rng.integers(0, 40))
        p[EHF["INVF"]:EHF["INVF"] + 5] = rng.integers(0, 4, 5)
        p[EHF["INVF_PREV"]:EHF["INVF_PREV"] + 5] = rng.integers(0, 4, 5)
        p[EHF["INVF_TBL"]:EHF["INVF_TBL"] + 5] = tbl
        p[EHF["FMASTER"]:EHF["FMASTER"] + len(fm)] = fm
        bw_prev[u] = rng.choice([0.0, 0.6, 0.75, 0.9, 0.98, 0.3, 0.01], 6).astype(np.float32)
        amp = 2.0 ** rng.uniform(-10, 10)
        for buf in (src, pv):
            kind = rng.integers(0, 4, 64)
            for b in range(64):
                if kind[b] == 0:      # noise
                    buf[0, u, :, b] = rng.standard_normal(EHF_ROWS) * amp
                    buf[1, u, :, b] = rng.standard_normal(EHF_ROWS) * amp
                elif kind[b] == 1:    # tone (nearly singular covariance, large alpha)
                    w, ph = rng.uniform(0, np.pi), rng.uniform(0, 6)
                    buf[0, u, :, b] = np.cos(w * t[:, 0] + ph) * amp
                    buf[1, u, :, b] = np.sin(w * t[:, 0] + ph) * amp
                elif kind[b] == 2:    # tone + weak noise
                    w, ph = rng.uniform(0, np.pi), rng.uniform(0, 6)
                    buf[0, u, :, b] = (np.cos(w * t[:, 0] + ph) + 1e-3 * rng.standard_normal(EHF_ROWS)) * amp
                    buf[1, u, :, b] = (np.sin(w * t[:, 0] + ph) + 1e-3 * rng.standard_normal(EHF_ROWS)) * amp
                # kind 3: silent band
    return dict(par=par, src_re=src[0], src_im=src[1], pv_re=pv[0], pv_im=pv[1], dst_re=dst[0], dst_im=dst[1],
                bw_prev=bw_prev)


def _esbr_hfgen_batch(fn, d, with_pv=True):
    n = d["par"].shape[0]
    dr, di = d["dst_re"].copy(), d["dst_im"].copy()
    bw = d["bw_prev"].copy()
    patch = d["patch_in"].copy() if "patch_in" in d else np.zeros((n, 8), np.int32)
    err = np.zeros(n, np.int32)
    c = lambda a: np.ascontiguousarray(a)
    fn(P(c(d["src_re"])), P(c(d["src_im"])), P(c(d["pv_re"])) if with_pv else None, P(c(d["pv_im"])) if with_pv else None,
       P(dr), P(di), P(c(d["par"])), P(bw), P(patch), P(err), n)
    return dr, di, bw, patch, err


def oracle_esbr_hfgen_batch(orc, d, with_pv=True):
    return _esbr_hfgen_batch(orc.lib.xo_esbr_generate_hf_batch, d, with_pv)


def ref_esbr_hfgen_batch(ref, d, with_pv=True):
    return _esbr_hfgen_batch(ref.lib.ref_esbr_generate_hf_batch, d, with_pv)


# ---- eSBR float envelope adjuster (ixheaacd_sbr_env_calc, ORIG_SBR) ------------------------------------------------------
EEC_IPAR_WORDS, EEC_FPAR_WORDS, EEC_STATE_WORDS = 288, 464, 640
EEC = dict(SB_START=0, SB_END=1, NUM_ENV=2, TRANS_ENV=3, SHORT_PREV=4, NUM_NOISE_ENV=5, NUM_SF_LO=6, NUM_SF_HI=7, NUM_NF=8,
           SMOOTHING_MODE=9, INTERPOL_FREQ=10, LIMITER_BANDS=11, LIMITER_GAINS=12, HARM_INDEX=13, PHASE_INDEX=14, START_UP=15,
           RESET=16, SBR_MODE=17, USF4=18, PATCHING_CHANGED=19, BORDER=24, FREQ_RES=33, NOISE_BORDER=41, INTER_TES=44,
           GATE_MODE=52, LIM_TABLE=56, TBL_NOISE=108, TBL_LO=116, TBL_HI=148, ADD_HARM=208, HARM_PREV=264,
           SFB_NRG=0, NOISE_FLOOR=448)


def esbr_random_phase(ref=None):
    """ixheaac_random_phase[512][2] (common/ixheaac_esbr_rom.c:437) — from the compiled reference when present, else the
    committed blob libxaac_b200/rom/esbr_random_phase.bin"""
    path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "libxaac_b200", "rom", "esbr_random_phase.bin")
    if ref is not None:
        out = np.zeros(1024, np.float32)
        ref.lib.ref_rom_esbr_random_phase(P(out))
        return out
    return np.fromfile(path, np.float32)


def synth_esbr_envcalc_units(n, seed):
    """Well-formed eSBR envelope-adjuster units: nested hi / lo / noise band tables, 1..5 envelopes with mixed resolution,
    transient / short flags, limiter tables cut from the lo table, sinusoids, smoothing history; every 32nd unit carries an
    invalid noise-envelope count (the reference returns IA_FATAL_ERROR after processing)."""
    rng = np.random.default_rng(seed)
    ipar = np.zeros((n, EEC_IPAR_WORDS), np.int32)
    fpar = np.zeros((n, EEC_FPAR_WORDS), np.float32)
    state = (2.0 ** rng.uniform(-6, 6, (n, EEC_STATE_WORDS))).astype(np.float32)
    re = np.zeros((n, 40, 64), np.float32)
    im = np.zeros((n, 40, 64), np.float32)
    for u in range(n):
        p = ipar[u]
        sbs = int(rng.integers(6, 33))
        hi = [sbs]
        while hi[-1] < 64 and len(hi) < 49:
            w = int(rng.integers(1, 4)) if hi[-1] < 40 else int(rng.integers(2, 6))
            if hi[-1] + w > 64 or (len(hi) > 4 and rng.random() < 0.05):
                break
            hi.append(hi[-1] + w)
        if len(hi) < 3:
            hi = [sbs, sbs + 2, sbs + 4]
        num_hi = len(hi) - 1
        lo = hi[::2] if num_hi % 2 == 0 else [hi[0]] + hi[1::2]
        num_lo = len(lo) - 1
        sbe = hi[-1]
        num_nf = int(rng.integers(1, min(5, num_lo) + 1))
        inner = sorted(rng.choice(lo[1:-1], size=num_nf - 1, replace=False).tolist()) if num_nf > 1 else []
        tn = [lo[0]] + inner + [sbe]
        tn = tn + [sbe] * (6 - len(tn))
        num_env = int(rng.integers(1, 6))
        b0, bl = int(rng.integers(0, 3)), int(rng.integers(14, 20))
        cuts = sorted(rng.choice(np.arange(b0 + 1, bl), size=num_env - 1, replace=False).tolist()) if num_env > 1 else []
        border = [b0] + cuts + [bl]
        if rng.random() < 0.05 and num_env > 1:
            border[1] = border[0]            # an empty envelope
        nne = 1 if num_env == 1 else 2
        nb = [b0, bl, 0] if nne == 1 else [b0, border[int(rng.integers(1, num_env))], bl]
        p[EEC["SB_START"]], p[EEC["SB_END"]], p[EEC["NUM_ENV"]] = sbs, sbe, num_env
        p[EEC["TRANS_ENV"]] = int(rng.integers(-1, num_env + 1))
        p[EEC["SHORT_PREV"]] = int(rng.choice([0, -1]))
        p[EEC["NUM_NOISE_ENV"]] = nne if u % 32 != 31 else 3
        p[EEC["NUM_SF_LO"]], p[EEC["NUM_SF_HI"]], p[EEC["NUM_NF"]] = num_lo, num_hi, num_nf
        p[EEC["SMOOTHING_MODE"]] = int(rng.random() < 0.4)
        p[EEC["INTERPOL_FREQ"]] = int(rng.random() < 0.6)
        p[EEC["LIMITER_BANDS"]] = int(rng.integers(0, 4))
        p[EEC["LIMITER_GAINS"]] = int(rng.integers(0, 4))
        p[EEC["HARM_INDEX"]] = int(rng.integers(0, 4))
        p[EEC["PHASE_INDEX"]] = int(rng.integers(0, 512))
        p[EEC["START_UP"]] = int(rng.random() < 0.15)
        p[EEC["SBR_MODE"]] = 1
        p[EEC["BORDER"]:EEC["BORDER"] + len(border)] = border
        p[EEC["FREQ_RES"]:EEC["FREQ_RES"] + num_env] = rng.integers(0, 2, num_env)
        p[EEC["NOISE_BORDER"]:EEC["NOISE_BORDER"] + 3] = nb
        for r in range(4):
            g = 1 if r == 0 else int(rng.integers(1, min(12, num_lo) + 1))
            ins = sorted(rng.choice(lo[1:-1], size=g - 1, replace=False).tolist()) if g > 1 else []
            lt = [0] + [x - sbs for x in ins] + [sbe - sbs]
            p[EEC["GATE_MODE"] + r] = g
            p[EEC["LIM_TABLE"] + 13 * r:EEC["LIM_TABLE"] + 13 * r + len(lt)] = lt
        p[EEC["TBL_NOISE"]:EEC["TBL_NOISE"] + 6] = tn
        p[EEC["TBL_LO"]:EEC["TBL_LO"] + len(lo)] = lo
        p[EEC["TBL_HI"]:EEC["TBL_HI"] + len(hi)] = hi
        p[EEC["ADD_HARM"]:EEC["ADD_HARM"] + num_hi] = rng.random(num_hi) < 0.12
        hp = (rng.random(64) < 0.2).astype(np.int8)
        p[EEC["HARM_PREV"]:EEC["HARM_PREV"] + 16] = hp.view(np.int32)
        fpar[u, :448] = 2.0 ** rng.uniform(-4, 26, 448)
        fpar[u, 448:458] = 2.0 ** rng.uniform(-8, 8, 10)
        if rng.random() < 0.1:
            fpar[u, 448 + int(rng.integers(0, 10))] = 0.0
        amp = 2.0 ** rng.uniform(-6, 12)
        re[u] = rng.standard_normal((40, 64)) * amp
        im[u] = rng.standard_normal((40, 64)) * amp
        if rng.random() < 0.1:
            re[u, :, sbs:sbs + 3] = 0
            im[u, :, sbs:sbs + 3] = 0
    return dict(re=re, im=im, ipar=ipar, fpar=fpar, state=state)


def _esbr_envcalc_batch(fn, d, pre=()):
    n = d["ipar"].shape[0]
    re, im, ipar, state = d["re"].copy(), d["im"].copy(), d["ipar"].copy(), d["state"].copy()
    err = np.zeros(n, np.int32)
    fn(*pre, P(re), P(im), P(ipar), P(np.ascontiguousarray(d["fpar"])), P(state), P(err), n)
    return re, im, ipar, state, err


def oracle_esbr_envcalc_batch(orc, d, rphase):
    return _esbr_envcalc_batch(orc.lib.xo_esbr_env_calc_batch, d, (P(rphase),))


def ref_esbr_envcalc_batch(ref, d):
    return _esbr_envcalc_batch(ref.lib.ref_esbr_env_calc_batch, d)


# ---- whole float eSBR stage (eSBR branch of ixheaacd_sbr_dec): composition of the oracle pieces --------------------------
ESD_KEYS = ("qmf_re", "qmf_im", "out_re", "out_im", "anal_states", "anal_pos", "synth_states", "synth_pos", "bw_prev", "patch",
            "ec_state")


def oracle_esbr_stage(orc, rphase, st, time_in, hf_par, ec_ipar, ec_fpar, rg_par):
    """One frame of n channels.  st: dict of ESD_KEYS arrays (copied, not modified).  Follows decoder/ixheaacd_sbr_dec.c:
    836-856 (history shift), :877 (analysis into rows 8..39), :921 (generate_hf), :953 (sbr_env_calc), :297-397 (regrouping),
    :583-654 (synthesis).  Returns (time_out [n,2048], st', ec_ipar', err [4,n])."""
    n = time_in.shape[0]
    s = {k: np.ascontiguousarray(v).copy() for k, v in st.items()}
    for k in ("qmf_re", "qmf_im", "out_re", "out_im"):
        s[k][:, 0:8] = s[k][:, 32:40]
    q, s["anal_states"], s["anal_pos"] = orc.esbr_anal_batch(time_in, s["anal_states"], s["anal_pos"])
    s["qmf_re"][:, 8:40, 0:32] = q[:, :, 0:32]
    s["qmf_im"][:, 8:40, 0:32] = q[:, :, 64:96]
    d = dict(par=hf_par, src_re=s["qmf_re"], src_im=s["qmf_im"], pv_re=s["qmf_re"], pv_im=s["qmf_im"], dst_re=s["out_re"],
             dst_im=s["out_im"], bw_prev=s["bw_prev"], patch_in=s["patch"])
    s["out_re"], s["out_im"], s["bw_prev"], s["patch"], e1 = oracle_esbr_hfgen_batch(orc, d, with_pv=False)
    d = dict(re=s["out_re"], im=s["out_im"], ipar=ec_ipar, fpar=ec_fpar, state=s["ec_state"])
    s["out_re"], s["out_im"], ipar2, s["ec_state"], e2 = oracle_esbr_envcalc_batch(orc, d, rphase)
    m = np.zeros((n, 32, 128), np.float32)
    k = np.arange(64)[None, :]
    for u in range(n):
        xo = np.where(np.arange(32) < rg_par[u, 2], rg_par[u, 0], rg_par[u, 1])[:, None]
        m[u, :, :64] = np.where(k < xo, s["qmf_re"][u, 2:34], s["out_re"][u, 2:34])
        m[u, :, 64:] = np.where(k < xo, s["qmf_im"][u, 2:34], s["out_im"][u, 2:34])
    out, s["synth_states"], s["synth_pos"] = orc.esbr_synth_batch(m, s["synth_states"], s["synth_pos"])
    return out, s, ipar2, np.stack([np.zeros(n, np.int32), e1, e2, np.zeros(n, np.int32)])


# ---- eSBR QMF harmonic transposer (ixheaacd_qmf_hbe_apply) ---------------------------------------------------------------
HBE_CFG_WORDS, HBE_ST_WORDS = 16, 3616
HBE_ST_TAIL, HBE_ST_SYNTH, HBE_ST_ANAL, HBE_ST_QIN, HBE_ST_QOUT = 0, 32, 416, 800, 2336


def hbe_rom():
    return np.fromfile(os.path.join(ROM_DIR, "hbe_rom.bin"), dtype=np.float32)


def oracle_hbe_batch(orc, cfg, state, qre, qim, pv_re=None, pv_im=None):
    """Returns (pv_re, pv_im, state_out, err); pv_* start from the given arrays (only bands start..end-1 are written)."""
    n = len(cfg)
    st = np.ascontiguousarray(state, np.float32).copy()
    pr = np.zeros((n, 32, 64), np.float32) if pv_re is None else np.ascontiguousarray(pv_re, np.float32).copy()
    pi = np.zeros((n, 32, 64), np.float32) if pv_im is None else np.ascontiguousarray(pv_im, np.float32).copy()
    err = np.zeros(n, np.int32)
    r = hbe_rom()
    orc.lib.xo_esbr_hbe_apply_batch(P(r), P(np.ascontiguousarray(cfg, np.int32)), P(st), P(np.ascontiguousarray(qre, np.float32)),
                                    P(np.ascontiguousarray(qim, np.float32)), P(pr), P(pi), P(err), n)
    return pr, pi, st, err


def ref_hbe_batch(ref, cfg, state, qre, qim, tbl=None, pv_re=None, pv_im=None):
    """The compiled ixheaacd_qmf_hbe_apply; tbl int16 [n,128] = {num_lo, num_hi, lo[0..num_lo], hi[0..num_hi]} (needed for
    synth_size 20, where the reference re-initialises from the frequency tables inside the call)."""
    n = len(cfg)
    st = np.ascontiguousarray(state, np.float32).copy()
    pr = np.zeros((n, 32, 64), np.float32) if pv_re is None else np.ascontiguousarray(pv_re, np.float32).copy()
    pi = np.zeros((n, 32, 64), np.float32) if pv_im is None else np.ascontiguousarray(pv_im, np.float32).copy()
    err = np.zeros(n, np.int32)
    t = None if tbl is None else np.ascontiguousarray(tbl, np.int16)
    ref.lib.ref_esbr_hbe_apply_batch(P(np.ascontiguousarray(cfg, np.int32)), P(st), P(np.ascontiguousarray(qre, np.float32)),
                                     P(np.ascontiguousarray(qim, np.float32)), P(pr), P(pi), None if t is None else P(t), P(err), n)
    return pr, pi, st, err


def hbe_tables(rng, kx, top):
    """A plausible pair of SBR frequency-band tables starting at QMF band kx and ending at `top`: HIGH with steps 1..3,
    LOW = every other HIGH border (as ixheaacd_calc_frq_bnd_tbls derives it)."""
    hi = [kx]
    while hi[-1] < top:
        hi.append(min(top, hi[-1] + int(rng.integers(1, 4))))
    lo = hi[::2] if (len(hi) - 1) % 2 == 0 else [hi[0]] + hi[1::2]
    t = np.zeros(128, np.int16)
    t[0], t[1] = len(lo) - 1, len(hi) - 1
    t[2:2 + len(lo)] = lo
    t[2 + len(lo):2 + len(lo) + len(hi)] = hi
    return t


def ref_hbe_reinit(ref, tbl):
    """ixheaacd_qmf_hbe_data_reinit on a table row of hbe_tables(): returns (ret, cfg[16])."""
    cfg = np.zeros(HBE_CFG_WORDS, np.int32)
    nlo, nhi = int(tbl[0]), int(tbl[1])
    lo = np.ascontiguousarray(tbl[2:2 + nlo + 1])
    hi = np.ascontiguousarray(tbl[2 + nlo + 1:2 + nlo + 1 + nhi + 1])
    ret = ref.lib.ref_esbr_hbe_reinit(P(lo), nlo, P(hi), nhi, P(cfg))
    return ret, cfg


def synth_hbe_units(n, seed, ref, pitch_mode="mixed", sizes=(4, 8, 12, 16, 20)):
    """n transposer units over a spread of configurations derived by the reference's own re-initialisation from random
    frequency tables (so synth_size 20 is drivable through the shim too), with random state and QMF input.
    pitch_mode: "zero" (plain products), "pitch" (cross products, pitch_in_bins >= 12) or "mixed"."""
    rng = np.random.default_rng(seed)
    cfg = np.zeros((n, HBE_CFG_WORDS), np.int32)
    tbl = np.zeros((n, 128), np.int16)
    for u in range(n):
        while True:
            s = sizes[u % len(sizes)]
            lo_kx = {4: 1, 8: 4, 12: 12, 16: 20, 20: 28}[s]
            kx = int(rng.integers(lo_kx, lo_kx + 8 if s != 4 else 4))
            if s == 20:
                kx = int(rng.integers(28, 36))
            top = int(rng.integers(min(63, kx + 6), 65))
            t = hbe_tables(rng, kx, top)
            ret, c = ref_hbe_reinit(ref, t)
            if ret == 0 and c[0] == s and c[4] >= 2 and c[1] >= 0 and c[1] + s <= 32 and not (c[4] >= 4 and c[10] <= 1):
                break
        if pitch_mode == "pitch" or (pitch_mode == "mixed" and u % 3 == 1):
            c[5] = int(rng.integers(12, 128))
        elif pitch_mode == "mixed" and u % 3 == 2:
            c[5] = int(rng.integers(0, 12))
        cfg[u], tbl[u] = c, t
    amp = 10.0 ** rng.uniform(-2, 3.5, size=(n, 1, 1))
    qre = (rng.standard_normal((n, 32, 64)) * amp).astype(np.float32)
    qim = (rng.standard_normal((n, 32, 64)) * amp).astype(np.float32)
    qre[::7, 5:9] = 0
    qim[::7, 5:9] = 0
    state = np.zeros((n, HBE_ST_WORDS), np.float32)
    a2 = amp[:, 0]
    state[:, :HBE_ST_QIN] = (rng.standard_normal((n, HBE_ST_QIN)) * a2).astype(np.float32)
    for u in range(n):  # qmf_in history is only ever non-zero inside the analysis bank's band range
        s, k = int(cfg[u, 0]), int(cfg[u, 1])
        q = np.zeros((12, 128), np.float32)
        q[:, 4 * k:4 * k + 4 * s] = rng.standard_normal((12, 4 * s)) * a2[u]
        state[u, HBE_ST_QIN:HBE_ST_QOUT] = q.ravel()
        o = np.zeros((10, 128), np.float32)
        o[:9, 2 * cfg[u, 2]:2 * cfg[u, 3]] = rng.standard_normal((9, 2 * (cfg[u, 3] - cfg[u, 2]))) * a2[u]
        state[u, HBE_ST_QOUT:] = o.ravel()
        state[u, HBE_ST_TAIL + s:HBE_ST_SYNTH] = 0          # words past the instance's sizes are not part of the state
        state[u, HBE_ST_SYNTH + 18 * s:HBE_ST_ANAL] = 0
        state[u, HBE_ST_ANAL + 18 * s:HBE_ST_QIN] = 0
    state[::5] = 0
    return cfg, tbl, state, qre, qim


ESH_KEYS = ESD_KEYS + ("pv_re", "pv_im", "hbe_state")


def oracle_esbr_hbe_stage(orc, rphase, st, time_in, hbe_cfg, hf_par, ec_ipar, ec_fpar, rg_par):
    """The eSBR stage with the harmonic transposer (hbe_flag = 1): as oracle_esbr_stage, but the core QMF arrays have 72 rows
    (delayed by ESBR_HBE_DELAY_OFFSET = 32 slots: rows 32..71 -> 0..39, analysis into rows 40..71, decoder/ixheaacd_sbr_dec.c:
    821-846, 877-880), the transposer turns the new slots into ph_vocod_qmf rows 8..39 (:896-907) and the HF generator reads them.
    Returns (time_out, st', ec_ipar', err [5, n])."""
    n = time_in.shape[0]
    s = {k: np.ascontiguousarray(v).copy() for k, v in st.items()}
    for k in ("qmf_re", "qmf_im"):
        s[k][:, 0:40] = s[k][:, 32:72].copy()
    for k in ("out_re", "out_im", "pv_re", "pv_im"):
        s[k][:, 0:8] = s[k][:, 32:40]
    q, s["anal_states"], s["anal_pos"] = orc.esbr_anal_batch(time_in, s["anal_states"], s["anal_pos"])
    s["qmf_re"][:, 40:72, 0:32] = q[:, :, 0:32]
    s["qmf_im"][:, 40:72, 0:32] = q[:, :, 64:96]
    pr, pi, s["hbe_state"], e4 = oracle_hbe_batch(orc, hbe_cfg, s["hbe_state"], s["qmf_re"][:, 40:72], s["qmf_im"][:, 40:72],
                                                  s["pv_re"][:, 8:40], s["pv_im"][:, 8:40])
    s["pv_re"][:, 8:40], s["pv_im"][:, 8:40] = pr, pi
    d = dict(par=hf_par, src_re=np.ascontiguousarray(s["qmf_re"][:, :40]), src_im=np.ascontiguousarray(s["qmf_im"][:, :40]),
             pv_re=s["pv_re"], pv_im=s["pv_im"], dst_re=s["out_re"], dst_im=s["out_im"], bw_prev=s["bw_prev"], patch_in=s["patch"])
    s["out_re"], s["out_im"], s["bw_prev"], s["patch"], e1 = oracle_esbr_hfgen_batch(orc, d, with_pv=True)
    d = dict(re=s["out_re"], im=s["out_im"], ipar=ec_ipar, fpar=ec_fpar, state=s["ec_state"])
    s["out_re"], s["out_im"], ipar2, s["ec_state"], e2 = oracle_esbr_envcalc_batch(orc, d, rphase)
    m = np.zeros((n, 32, 128), np.float32)
    k = np.arange(64)[None, :]
    for u in range(n):
        xo = np.where(np.arange(32) < rg_par[u, 2], rg_par[u, 0], rg_par[u, 1])[:, None]
        m[u, :, :64] = np.where(k < xo, s["qmf_re"][u, 2:34], s["out_re"][u, 2:34])
        m[u, :, 64:] = np.where(k < xo, s["qmf_im"][u, 2:34], s["out_im"][u, 2:34])
    out, s["synth_states"], s["synth_pos"] = orc.esbr_synth_batch(m, s["synth_states"], s["synth_pos"])
    z = np.zeros(n, np.int32)
    return out, s, ipar2, np.stack([z, e1, e2, z, e4])


# ---- float parametric stereo (ixheaacd_esbr_apply_ps) ----
FPS_PAR_WORDS, FPS_HST_WORDS, FPS_SIDE_WORDS, FPS_ST_WORDS = 386, 228, 1024, 4368


def fps_rom():
    return np.fromfile(os.path.join(ROOT, "libxaac_b200", "rom", "fps_rom.bin"), np.float32)


def fps_fresh_state(n):
    """The instance right after ixheaacd_create_ps_esbr_dec (ps_dec_flt.c:297-379): everything zero, h11 / h12 real parts one."""
    st = np.zeros((n, FPS_ST_WORDS), np.float32)
    hst = np.zeros((n, FPS_HST_WORDS), np.float32)
    hst[:, 0:40] = 1.0
    return st, hst


def synth_fps_frame(n, rng, good_borders=True):
    """One frame of inputs for n independent mono + PS channels: the low-band QMF arrays (left slot i = row 2 + i, rows 34..39
    the look-ahead) and the frame's PS parameters in the shim's par layout."""
    gain = (10.0 ** rng.uniform(0.0, 3.5, (n, 1, 1))).astype(np.float32)
    tilt = np.exp(-np.arange(64) / rng.uniform(6, 40, (n, 1, 1))).astype(np.float32)
    low_re = (rng.standard_normal((n, 40, 64)).astype(np.float32) * gain * tilt).astype(np.float32)
    low_im = (rng.standard_normal((n, 40, 64)).astype(np.float32) * gain * tilt).astype(np.float32)
    par = np.zeros((n, FPS_PAR_WORDS), np.int32)
    for u in range(n):
        ne = int(rng.integers(1, 6))
        inner = np.sort(rng.choice(np.arange(1, 32), ne - 1, replace=False)) if ne > 1 else np.zeros(0, np.int64)
        b = np.concatenate([[0], inner, [32]])
        if not good_borders:
            b[-1] = 30
        par[u, 0] = ne
        par[u, 1:1 + len(b)] = b
        usb = int(rng.integers(24, 65))
        par[u, 7] = usb
        fine = int(rng.integers(0, 2))
        par[u, 8] = fine
        par[u, 9] = int(rng.integers(0, 3))
        lim = 15 if fine else 7
        par[u, 16:116] = rng.integers(-lim, lim + 1, 100)
        par[u, 116:216] = rng.integers(0, 8, 100)
        if rng.random() < 0.75:
            par[u, 216:301] = rng.integers(0, 8, 85)
            par[u, 301:386] = rng.integers(0, 8, 85)
        low_re[u, :, usb:] = 0
        low_im[u, :, usb:] = 0
    return low_re, low_im, par


def ref_fps_batch(ref, low_re, low_im, par, state, hst):
    """The compiled ixheaacd_esbr_apply_ps through oracle/ref_shim_fps.c.  Returns dict(side, commit, left, right, state, hst, rc):
    side / commit are what the drop-in's host code (b200_fps_side) derives from the same parameters before the call."""
    n = len(par)
    st = np.ascontiguousarray(state, np.float32).copy()
    h = np.ascontiguousarray(hst, np.float32).copy()
    side = np.zeros((n, FPS_SIDE_WORDS), np.float32)
    commit = np.zeros((n, FPS_HST_WORDS), np.float32)
    left = np.zeros((n, 32, 128), np.float32)
    right = np.zeros((n, 32, 128), np.float32)
    ref.lib.ref_fps_apply_batch.argtypes = [ctypes.c_int64] + [ctypes.c_void_p] * 9
    rc = ref.lib.ref_fps_apply_batch(n, P(np.ascontiguousarray(low_re, np.float32)), P(np.ascontiguousarray(low_im, np.float32)),
                                     P(np.ascontiguousarray(par, np.int32)), P(st), P(h), P(side), P(commit), P(left), P(right))
    return dict(side=side, commit=commit, left=left, right=right, state=st, hst=h, rc=rc)


def ref_fps_side_batch(ref, par, hst):
    """The drop-in's host-side preparation alone (b200_fps_side + b200_fps_commit through oracle/ref_shim_fps.c): returns
    (side [n, 1024], hst advanced by the frame)."""
    n = len(par)
    h = np.ascontiguousarray(hst, np.float32).copy()
    side = np.zeros((n, FPS_SIDE_WORDS), np.float32)
    ref.lib.ref_fps_side_batch.argtypes = [ctypes.c_int64] + [ctypes.c_void_p] * 3
    rc = ref.lib.ref_fps_side_batch(n, P(np.ascontiguousarray(par, np.int32)), P(h), P(side))
    assert rc == 0
    return side, h


ESP_KEYS = ("qmf_re", "qmf_im", "out_re", "out_im", "pv_re", "pv_im", "anal_states", "anal_pos", "synth_states", "synth_pos",
            "bw_prev", "patch", "ec_state", "hbe_state")


def heaacv2_esbr_units(g, n):
    """Initial state of n mono + PS elements for the HE-AACv2 default-mode stage: the eSBR part tiled from the channels of the
    tapped -harmonic_sbr:1 stream (tests/golden/esbr_hbe_stage_tapped.npz), PS instance and second synthesis bank fresh."""
    ch = np.arange(n) % 2
    st = {k: np.ascontiguousarray(g["in0_" + k][ch]).copy() for k in ESP_KEYS}
    st["ps_state"], st["ps_hst"] = fps_fresh_state(n)
    st["synth_states_r"] = np.zeros((n, 1280), np.int32)
    st["synth_pos_r"] = np.zeros((n, 2), np.int32)
    return st


def ref_heaacv2_esbr_chain(ref, st, time_in, hbe_cfg, hbe_tbl, hf_par, ec_ipar, ec_fpar, rg, ps_par, a=0, b=None, q6=None):
    """ref_heaacv2_esbr_chain_batch (oracle/ref_shim_fps.c) on the state dict of heaacv2_esbr_units, updated in place; ec_ipar is
    in/out.  Returns (out_l, out_r, err)."""
    n = len(time_in)
    b = n if b is None else b
    own = q6 is None
    if own:
        q6 = np.ascontiguousarray(np.concatenate([st[k].reshape(n, -1) for k in ESP_KEYS[:6]], 1))
    out_l = np.zeros((n, 2048), np.float32)
    out_r = np.zeros((n, 2048), np.float32)
    err = np.zeros(n, np.int32)
    ref.lib.ref_heaacv2_esbr_chain_batch(P(np.ascontiguousarray(time_in, np.float32)), P(q6), P(st["anal_states"]), P(st["anal_pos"]),
                                         P(st["synth_states"]), P(st["synth_pos"]), P(st["bw_prev"]), P(st["patch"]), P(st["ec_state"]),
                                         P(st["hbe_state"]), P(np.ascontiguousarray(hbe_cfg, np.int32)), P(hbe_tbl),
                                         P(np.ascontiguousarray(hf_par, np.int32)), P(ec_ipar), P(np.ascontiguousarray(ec_fpar, np.float32)),
                                         P(np.ascontiguousarray(rg, np.int32)), P(np.ascontiguousarray(ps_par, np.int32)),
                                         P(st["ps_state"]), P(st["ps_hst"]), P(st["synth_states_r"]), P(st["synth_pos_r"]),
                                         P(out_l), P(out_r), P(err), int(a), int(b))
    if own:
        o = 0
        for k in ESP_KEYS[:6]:
            w = st[k][0].size
            st[k][...] = q6[:, o:o + w].reshape(st[k].shape)
            o += w
    return out_l, out_r, err


def synth_esbr_envcalc_tes_units(n, seed):
    """synth_esbr_envcalc_units with inter-TES modes 0..3 on the envelopes and a low band (qmf_buf rows) for them to read"""
    d = synth_esbr_envcalc_units(n, seed)
    rng = np.random.default_rng(seed + 1000)
    for u in range(n):
        ne = int(d["ipar"][u, EEC["NUM_ENV"]])
        d["ipar"][u, 44:44 + ne] = rng.integers(0, 4, ne)
    amp = (2.0 ** rng.uniform(-6, 12, (n, 1, 1))).astype(np.float32)
    d["low_re"] = (rng.standard_normal((n, 40, 64)).astype(np.float32) * amp).astype(np.float32)
    d["low_im"] = (rng.standard_normal((n, 40, 64)).astype(np.float32) * amp).astype(np.float32)
    return d


def ref_esbr_envcalc_tes_batch(ref, d):
    n = d["ipar"].shape[0]
    re, im, ipar, state = d["re"].copy(), d["im"].copy(), d["ipar"].copy(), d["state"].copy()
    err = np.zeros(n, np.int32)
    ref.lib.ref_esbr_env_calc_tes_batch(P(re), P(im), P(np.ascontiguousarray(d["low_re"])), P(np.ascontiguousarray(d["low_im"])),
                                        P(ipar), P(np.ascontiguousarray(d["fpar"])), P(state), P(err), n)
    return re, im, ipar, state, err


# ---- AAC pre-IMDCT spectral stage (ixheaacd_channel_pair_process) ----
SPS_BYTES, SPS_CH, SPS_CH_BYTES = 3712, 544, 1584
SFB_LONG_44 = [0, 4, 8, 12, 16, 20, 24, 28, 32, 36, 40, 48, 56, 64, 72, 80, 88, 96, 108, 120, 132, 144, 160, 176, 196, 216, 240, 264,
               292, 320, 352, 384, 416, 448, 480, 512, 544, 576, 608, 640, 672, 704, 736, 768, 800, 832, 864, 896, 928, 1024]
SFB_SHORT_44 = [0, 4, 8, 12, 16, 20, 28, 36, 44, 56, 68, 80, 96, 112, 128]


def synth_sps_units(n, seed, tns=True, stereo_tools=True, pns=False, tns_prob=0.7, pns_prob=0.6):
    """Elements for the AAC-LC spectral stage at 44.1 kHz: single channels and pairs, long / start / stop / eight-short sequences
    with random grouping, random M/S masks and intensity bands, up to three TNS filters per window (orders up to 12 / 7, both
    directions and resolutions), spectra of mixed magnitude up to full scale."""
    rng = np.random.default_rng(seed)
    rec = np.zeros((n, SPS_BYTES), np.uint8)
    spec = np.zeros((n, 2, 1024), np.int32)
    for u in range(n):
        r = rec[u]
        hdr = r[:32].view(np.int32)
        num_ch = 2 if rng.random() < 0.8 else 1
        common = int(num_ch == 2 and rng.random() < 0.8)
        hdr[0], hdr[1] = num_ch, common
        ws0 = int(rng.choice([0, 0, 1, 2, 2, 3]))
        groups0 = None
        for c in range(num_ch):
            b = r[SPS_CH + c * SPS_CH_BYTES: SPS_CH + (c + 1) * SPS_CH_BYTES]
            w = b[:32].view(np.int32)
            ws = ws0 if (common or c == 0) else int(rng.choice([0, 1, 2, 3]))
            short = ws == 2
            tbl = SFB_SHORT_44 if short else SFB_LONG_44
            nsfb = len(tbl) - 1
            if common and c == 1:
                max_sfb, glens = int(r[SPS_CH:][:32].view(np.int32)[1]), groups0
            else:
                max_sfb = int(rng.integers(1, nsfb + 1))
                if short:
                    cuts = sorted(rng.choice(np.arange(1, 8), int(rng.integers(0, 5)), replace=False).tolist())
                    edges = [0] + cuts + [8]
                    glens = [edges[i + 1] - edges[i] for i in range(len(edges) - 1)]
                else:
                    glens = [1]
                groups0 = glens
            pa = int(pns and rng.random() < pns_prob)
            w[0], w[1], w[2], w[3], w[4], w[5] = ws, max_sfb, len(glens), pa, 14 if short else 42, 4
            b[32:32 + len(glens)] = np.array(glens, np.uint8)
            cbk = rng.integers(1, 12, 128).astype(np.int8)
            if stereo_tools and c == 1:
                m = rng.random(128) < 0.25
                cbk[m] = rng.choice([14, 15], int(m.sum()))
            if pa:
                m = rng.random(128) < 0.3
                cbk[m] = 13
                b[1456:1584] = m.astype(np.uint8)
            b[40:168] = cbk.view(np.uint8)
            b[168:424].view(np.int16)[:] = rng.integers(-40, 60, 128)
            ti = b[424:424 + 924]
            if tns and rng.random() < tns_prob:
                ti[:4].view(np.int32)[0] = 1
                for win in range(8 if short else 1):
                    nf = int(rng.integers(0, 2 if short else 4))
                    ti[4 + win] = nf
                    top = nsfb
                    for f in range(nf):
                        fb = ti[12 + (win * 3 + f) * 38: 12 + (win * 3 + f + 1) * 38]
                        length = int(rng.integers(0, top + 1))
                        fb[:4].view(np.int16)[:] = [max(top - length, 0), top]
                        top = max(top - length, 0)
                        order = int(rng.integers(0, (7 if short else 12) + 1))
                        res = int(rng.integers(0, 2))
                        fb[4] = np.uint8(255 if rng.random() < 0.5 else 1)   # direction -1 / 1
                        fb[5], fb[6] = res, order
                        lim = 8 if res else 4
                        fb[7:7 + order] = rng.integers(-lim, lim, order).astype(np.int8).view(np.uint8)
            b[1348:1348 + 2 * len(tbl)].view(np.int16)[:] = tbl
            mag = int(rng.choice([8, 14, 20, 26, 30, 31]))
            x = rng.integers(-(1 << mag), 1 << mag, 1024, dtype=np.int64)
            if rng.random() < 0.3:
                x[rng.random(1024) < 0.02] = rng.choice([-(1 << 31), (1 << 31) - 1])
            if short:
                for wdw in range(8):
                    x[128 * wdw + tbl[max_sfb]:128 * (wdw + 1)] = 0
            else:
                x[tbl[max_sfb]:] = 0
            spec[u, c] = np.clip(x, -(1 << 31), (1 << 31) - 1).astype(np.int32)
        if stereo_tools:
            r[32:544] = (rng.random(512) < 0.4).astype(np.uint8)
        if pns and rng.random() < 0.3:
            r[16:32] = rng.integers(0, 256, 16).astype(np.uint8)   # correlation flags the parser may have left
    return spec, rec


def ref_channel_pair_process(ref, spec, rec, seed=None):
    s = np.ascontiguousarray(spec, np.int32).copy()
    err = np.zeros(len(rec), np.int32)
    sd = np.zeros(len(rec), np.int32) if seed is None else np.ascontiguousarray(seed, np.int32).copy()
    ref.lib.ref_channel_pair_process_batch.argtypes = [ctypes.c_int64] + [ctypes.c_void_p] * 4
    ref.lib.ref_channel_pair_process_batch(len(rec), P(s), P(np.ascontiguousarray(rec, np.uint8)), P(sd), P(err))
    return (s, err) if seed is None else (s, err, sd)


# ---- SBR side-info dequantisation (ixheaacd_dec_sbrdata) ---------------------------------------------------------------
SD_WORDS, SD_CH, SD_CH_WORDS = 1304, 8, 648
SDC = dict(NUM_SF_LO=0, NUM_SF_HI=1, NUM_NF=2, NUM_TIME_SLOTS=3, ERR_FLAG=4, ERR_FLAG_PREV=5, HDR_AMP_RES=6, NUM_NOISE_SFAC=7,
           NUM_ENV=8, NUM_NOISE_ENV=9, TRANSIENT_ENV=10, AMP_RES=11, COUPLING=12, NUM_ENV_SFAC=13, MAX_QMF_SB=14, FREQ_RES=16,
           BORDER=24, NOISE_BORDER=33, DIR=36, DIR_NOISE=44, INVF=46, ADD_HARM=56, ENV=112, NOISE=560, PREV_NRG=570,
           PREV_NOISE=626, PREV_AMP_RES=631, PREV_END_POS=632, PREV_MAX_QMF=633, PREV_COUPLING=634, PREV_INVF=635)


def synth_sbrdata_records(n, seed):
    """Seeded XAAC_SD_* records (include/xaac_b200.h) as the SBR payload parser leaves them before ixheaacd_dec_sbrdata: valid
    frame grids and band counts, delta-coded envelope / noise indices in both directions and resolutions, mono elements, plain
    and coupled pairs with own or shared headers; a fraction of the elements has a timing mismatch against the previous frame
    (concealment / timing compensation / coupling-mode change), raised error flags, or values that fail the range check (retry
    through the concealment)."""
    rng = np.random.default_rng(seed)
    rec = np.zeros((n, SD_WORDS), np.int16)
    for u in range(n):
        two = rng.random() < 0.6
        shared = two and rng.random() < 0.5
        coupled = two and rng.random() < 0.5
        rec[u, 0] = 2 if two else 1
        rec[u, 1] = 1 if shared else 0
        nlo = int(rng.integers(3, 25))
        kind = rng.random()
        nhi = 2 * nlo - int(rng.integers(0, 2)) if kind < 0.6 else int(rng.integers(nlo, min(3 * nlo, 56) + 1))
        nhi = max(min(nhi, 56), nlo)
        nnf = int(rng.integers(1, 6))
        hamp = int(rng.integers(0, 2))
        mismatch = rng.random()
        for c in range(2 if two else 1):
            b = rec[u, SD_CH + c * SD_CH_WORDS: SD_CH + (c + 1) * SD_CH_WORDS]
            b[SDC["NUM_SF_LO"]], b[SDC["NUM_SF_HI"]], b[SDC["NUM_NF"]], b[SDC["NUM_TIME_SLOTS"]] = nlo, nhi, nnf, 16
            b[SDC["HDR_AMP_RES"]] = hamp
            if rng.random() < 0.08:
                b[SDC["ERR_FLAG"]] = 1
            if rng.random() < 0.08:
                b[SDC["ERR_FLAG_PREV"]] = 1
            nenv = int(rng.integers(1, 6))
            cuts = np.sort(rng.choice(np.arange(1, 16), nenv - 1, replace=False)) if nenv > 1 else np.array([], int)
            start = int(rng.integers(0, 3))
            border = np.concatenate([[start], np.maximum(cuts, start + 1), [16]])
            border = np.maximum.accumulate(border)
            b[SDC["NUM_ENV"]] = nenv
            b[SDC["BORDER"]: SDC["BORDER"] + nenv + 1] = border
            nne = 2 if (nenv > 1 and rng.random() < 0.6) else 1
            b[SDC["NUM_NOISE_ENV"]] = nne
            b[SDC["NOISE_BORDER"]] = start
            b[SDC["NOISE_BORDER"] + 1] = border[nenv // 2] if nne == 2 else 16
            b[SDC["NOISE_BORDER"] + 2] = 16 if nne == 2 else 0
            b[SDC["TRANSIENT_ENV"]] = int(rng.integers(-1, nenv))
            amp = int(rng.integers(0, 2))
            b[SDC["AMP_RES"]] = amp
            b[SDC["COUPLING"]] = (1 if c == 0 else 2) if coupled else 0
            fres = rng.integers(0, 2, 8)
            b[SDC["FREQ_RES"]: SDC["FREQ_RES"] + 8] = fres
            dirs = rng.integers(0, 2, 8)
            b[SDC["DIR"]: SDC["DIR"] + 8] = dirs
            b[SDC["DIR_NOISE"]: SDC["DIR_NOISE"] + 2] = rng.integers(0, 2, 2)
            b[SDC["INVF"]: SDC["INVF"] + 10] = rng.integers(0, 4, 10)
            b[SDC["ADD_HARM"]: SDC["ADD_HARM"] + 56] = rng.integers(0, 2, 56)
            b[SDC["MAX_QMF_SB"]] = int(rng.integers(8, 33))
            off = 0
            big = rng.random() < 0.06
            for i in range(nenv):
                nsb = nhi if fres[i] else nlo
                if dirs[i] == 0:
                    d = rng.integers(-5, 6, nsb)
                    d[0] = rng.integers(0, 45 >> amp)
                else:
                    d = rng.integers(-3, 4, nsb)
                if big and i == nenv - 1:
                    d[int(rng.integers(0, nsb))] += 90
                b[SDC["ENV"] + off: SDC["ENV"] + off + nsb] = d
                off += nsb
            b[SDC["NUM_ENV_SFAC"]] = off
            nf = rng.integers(-3, 4, 10)
            nf[0] = rng.integers(0, 30)
            nf[nnf] = rng.integers(0, 30)
            b[SDC["NOISE"]: SDC["NOISE"] + 10] = nf
            b[SDC["NUM_NOISE_SFAC"]] = int(rng.integers(0, 11))
            b[SDC["PREV_NRG"]: SDC["PREV_NRG"] + 56] = rng.integers(-2, 60 >> amp, 56)
            b[SDC["PREV_NOISE"]: SDC["PREV_NOISE"] + 5] = rng.integers(0, 31, 5)
            b[SDC["PREV_AMP_RES"]] = int(rng.integers(0, 2))
            b[SDC["PREV_END_POS"]] = 16 + start + (0 if mismatch > 0.25 else int(rng.integers(-1, 3)))
            b[SDC["PREV_MAX_QMF"]] = int(rng.integers(8, 33))
            # a mono element has no second channel to take energies from: its previous coupling mode is always off
            b[SDC["PREV_COUPLING"]] = b[SDC["COUPLING"]] if (rng.random() < 0.8 or not two) else int(rng.integers(0, 3))
            b[SDC["PREV_INVF"]: SDC["PREV_INVF"] + 10] = rng.integers(0, 4, 10)
    return rec


PSD_WORDS = 576
PSD = dict(DATA_PRESENT=0, ENABLE_IID=1, ENABLE_ICC=2, IID_MODE=3, ICC_MODE=4, IID_QUANT=5, FRAME_CLASS=6, NUM_ENV=7, FRAME_SIZE=8,
           BORDER=9, IID_DT=16, ICC_DT=21, IID_TABLE=32, ICC_TABLE=270, IID_PREV=508, ICC_PREV=542)


def synth_psdata_records(n, seed):
    """Seeded XAAC_PSD_* records as ixheaacd_read_ps_data leaves them: 10 / 20 / 34-band IID and ICC sets, coarse and fine IID
    quantisation, time- and frequency-direction deltas (with values that hit the clamps), fixed and variable envelope borders
    (some short of the frame end, some out of order), frames without PS data, 960-sample frames."""
    rng = np.random.default_rng(seed)
    rec = np.zeros((n, PSD_WORDS), np.int16)
    for u in range(n):
        r = rec[u]
        r[PSD["DATA_PRESENT"]] = rng.random() < 0.85
        r[PSD["ENABLE_IID"]] = rng.random() < 0.85
        r[PSD["ENABLE_ICC"]] = rng.random() < 0.85
        r[PSD["IID_MODE"]] = rng.integers(0, 3)
        r[PSD["ICC_MODE"]] = rng.integers(0, 3)
        r[PSD["IID_QUANT"]] = rng.integers(0, 2)
        fc = int(rng.random() < 0.4)
        r[PSD["FRAME_CLASS"]] = fc
        cols = 30 if rng.random() < 0.1 else 32
        r[PSD["FRAME_SIZE"]] = 960 if cols == 30 else 1024
        ne = int(rng.choice([0, 1, 2, 4])) if fc == 0 else int(rng.integers(1, 5))
        r[PSD["NUM_ENV"]] = ne
        if fc:
            b = np.sort(rng.integers(1, cols + 1, ne))
            if rng.random() < 0.2 and ne > 1:
                b = rng.permutation(b)
            r[PSD["BORDER"] + 1: PSD["BORDER"] + 1 + ne] = b
            r[PSD["BORDER"]] = rng.integers(0, 3)
        else:
            r[PSD["BORDER"]: PSD["BORDER"] + 7] = rng.integers(0, 33, 7)
        r[PSD["IID_DT"]: PSD["IID_DT"] + 5] = rng.integers(0, 2, 5)
        r[PSD["ICC_DT"]: PSD["ICC_DT"] + 5] = rng.integers(0, 2, 5)
        r[PSD["IID_TABLE"]: PSD["IID_TABLE"] + 238] = rng.integers(-6, 7, 238)
        r[PSD["ICC_TABLE"]: PSD["ICC_TABLE"] + 238] = rng.integers(-3, 4, 238)
        r[PSD["IID_PREV"]: PSD["IID_PREV"] + 34] = rng.integers(-15, 16, 34)
        r[PSD["ICC_PREV"]: PSD["ICC_PREV"] + 34] = rng.integers(0, 8, 34)
    return rec
