"""GPU parity tests for the float parametric stereo kernel (xaac_b200_esbr_ps_apply_dev = batched ixheaacd_esbr_apply_ps,
decoder/ixheaacd_ps_dec_flt.c:381-505, fused with ixheaacd_esbr_synthesis_regrp and the look-ahead slots) against the COMPILED
reference function (oracle/_ref, through oracle/ref_shim_fps.c) and against the golden records it produced
(tests/golden/esbr_ps_ref.npz).  Float results are compared as bit patterns: the device evaluates no libm function on this path
(the mixing matrices come from the host), every sum keeps the reference's association and the file is built without FMA
contraction."""
import os

import numpy as np
import pytest

from tests import oracle_util as ou
from tests.test_ps_flt_host import GOLD, bits, golden_inputs

pytestmark = pytest.mark.gpu


def run_gpu(ctx, low_re, low_im, side, state, high_re=None, high_im=None, rg=None):
    import torch
    import libxaac_b200 as xb
    n = len(side)
    t = lambda a, dt: torch.from_numpy(np.ascontiguousarray(a, dt)).cuda()
    hr = np.zeros((n, 40, 64), np.float32) if high_re is None else high_re
    hi = np.zeros((n, 40, 64), np.float32) if high_im is None else high_im
    rgp = np.tile(np.array([64, 64, 0, 0], np.int32), (n, 1)) if rg is None else rg
    st = t(state, np.float32)
    left, right, err = xb.esbr_apply_ps(ctx, t(low_re, np.float32), t(low_im, np.float32), t(hr, np.float32), t(hi, np.float32),
                                        t(rgp, np.int32), t(side, np.float32), st)
    torch.cuda.synchronize()
    return left.cpu().numpy(), right.cpu().numpy(), st.cpu().numpy(), err.cpu().numpy()


def compare(got, r, what):
    left, right, st, err = got
    assert (err == 0).all(), (what, err[:8])
    for a, b, nm in ((left, r["left"], "left"), (right, r["right"], "right"), (st, r["state"], "state")):
        d = bits(a) != bits(b)
        assert not d.any(), f"{what}: {nm} differs in {int(d.sum())} cells, first {np.argwhere(d)[:5].tolist()}"


def test_ps_kernel_matches_golden_records(ctx):
    g = np.load(GOLD)
    n = g["par"].shape[1]
    st, _ = ou.fps_fresh_state(n)
    for f in range(g["par"].shape[0]):
        lr, li = golden_inputs(g, f)
        got = run_gpu(ctx, lr, li, g["side"][f], st)
        compare(got, dict(left=g["left"][f], right=g["right"][f], state=g["state_out"][f]), f"golden frame {f}")
        st = got[2]


def test_ps_kernel_matches_compiled_reference_over_frames(ctx, ref):
    rng = np.random.default_rng(11)
    n = 600  # more units than one wave of warps (148 SMs x 2 CTAs x 4 warps would be 1184; the tail CTA is partial)
    st, hst = ou.fps_fresh_state(n)
    st_gpu = st.copy()
    for f in range(4):
        lr, li, par = ou.synth_fps_frame(n, rng)
        r = ou.ref_fps_batch(ref, lr, li, par, st, hst)
        assert r["rc"] == 0
        got = run_gpu(ctx, lr, li, r["side"], st_gpu)
        compare(got, r, f"frame {f}")
        st, hst, st_gpu = r["state"], r["hst"], got[2]
    assert np.abs(r["right"]).max() > 10.0


def test_ps_kernel_regroups_like_the_synthesis_bank(ctx, ref):
    """72-row low-band arrays (harmonic-transposer delay) and a real cross-over: the kernel's own regrouping against the
    reference fed with host-regrouped arrays."""
    rng = np.random.default_rng(12)
    n = 64
    st, hst = ou.fps_fresh_state(n)
    lr, li, par = ou.synth_fps_frame(n, rng)
    low_re = (rng.standard_normal((n, 72, 64)) * 300).astype(np.float32)
    low_im = (rng.standard_normal((n, 72, 64)) * 300).astype(np.float32)
    high_re = (rng.standard_normal((n, 40, 64)) * 300).astype(np.float32)
    high_im = (rng.standard_normal((n, 40, 64)) * 300).astype(np.float32)
    rg = np.zeros((n, 4), np.int32)
    rg[:, 0] = rng.integers(8, 40, n)
    rg[:, 1] = rng.integers(8, 40, n)
    rg[:, 2] = rng.integers(0, 12, n)
    mre, mim = low_re[:, :40].copy(), low_im[:, :40].copy()
    for u in range(n):
        for s in range(32):
            xo = rg[u, 0] if s < rg[u, 2] else rg[u, 1]
            mre[u, 2 + s, xo:] = high_re[u, 2 + s, xo:]
            mim[u, 2 + s, xo:] = high_im[u, 2 + s, xo:]
    r = ou.ref_fps_batch(ref, mre, mim, par, st, hst)
    got = run_gpu(ctx, low_re, low_im, r["side"], st, high_re, high_im, rg)
    compare(got, r, "regrouped")


def test_ps_kernel_refuses_unsupported_borders(ctx):
    g = np.load(GOLD)
    n = g["par"].shape[1]
    st = np.random.default_rng(3).standard_normal((n, ou.FPS_ST_WORDS)).astype(np.float32)
    st[:, 4300:] = 0
    side = g["side"][0].copy()
    iv = side.view(np.int32)
    iv[0, 1 + iv[0, 0]] = 30      # last border short of the frame
    iv[1, 0] = 6                  # too many envelopes
    lr, li = golden_inputs(g, 0)
    left, right, st_out, err = run_gpu(ctx, lr, li, side, st)
    assert err[0] == -2 and err[1] == -2 and (err[2:] == 0).all()
    assert np.array_equal(bits(st_out[:2]), bits(st[:2]))


def _stage_frame(g, f, n, rng):
    """records of frame f of the tapped -harmonic_sbr:1 stream tiled over n mono + PS elements, plus random PS parameters"""
    r = (np.arange(n) % 2) + 2 * (f % 6)
    h = g["head"][r]
    rg = np.stack([h[:, 7], h[:, 8], 2 * h[:, 9], 0 * h[:, 9]], 1).astype(np.int32)
    gain = np.ldexp(1.0, -((np.arange(n) // 2) % 3)).astype(np.float32)[:, None]
    time_in = (g["time_in"][r] * gain).astype(np.float32)
    _, _, par = ou.synth_fps_frame(n, rng)
    ipar = np.ascontiguousarray(g["ec_ipar_in"][r]).copy()
    par[:, 7] = ipar[:, 1]  # usb = sub_band_end
    return time_in, g["hbe_cfg"][r], g["hf_par"][r], ipar, g["ec_fpar"][r], rg, par


def test_mono_ps_stage_matches_reference_chain(ctx, ref):
    """xaac_b200_esbr_dec_ps_dev (analysis -> transposer -> HF generator -> envelope adjuster -> PS -> two synthesis banks, state
    resident) against the same chain assembled from the compiled reference's stage functions, 4 frames, bit for bit"""
    import torch
    import libxaac_b200 as xb
    g = np.load(os.path.join(os.path.dirname(GOLD), "esbr_hbe_stage_tapped.npz"))
    rng = np.random.default_rng(21)
    n = 96
    st = ou.heaacv2_esbr_units(g, n)
    c0 = g["hbe_cfg"][0]
    tbl = np.zeros(128, np.int16)
    tbl[:6] = [1, 1, c0[2], c0[3], c0[2], c0[3]]
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
    s = xb.EsbrDecPsBatch(n)
    for k in ou.ESP_KEYS:
        getattr(s, k).copy_(t(st[k]))
    s.ps_state.copy_(t(st["ps_state"]))
    hst = st["ps_hst"].copy()
    for f in range(4):
        time_in, hc, hf, ipar, fp, rg, par = _stage_frame(g, f, n, rng)
        side, hst = ou.ref_fps_side_batch(ref, par, hst)
        ipar_ref = ipar.copy()
        wl, wr, err = ou.ref_heaacv2_esbr_chain(ref, st, time_in, hc, tbl, hf, ipar_ref, fp, rg, par)
        assert not err.any(), err
        d_ipar = t(ipar)
        ol, orr, e = xb.esbr_dec_ps(ctx, s, t(time_in), t(hc), t(hf), d_ipar, t(fp), t(rg), t(side))
        torch.cuda.synchronize()
        assert int(e.abs().max()) == 0, e.cpu().numpy()
        assert np.array_equal(bits(ol.cpu().numpy()), bits(wl)), f"frame {f}: left"
        assert np.array_equal(bits(orr.cpu().numpy()), bits(wr)), f"frame {f}: right"
        assert np.array_equal(d_ipar.cpu().numpy(), ipar_ref), f"frame {f}: in/out parameter words"
        for k in ("anal_states", "synth_states", "synth_pos", "ec_state", "bw_prev", "hbe_state", "ps_state", "synth_states_r", "synth_pos_r"):
            assert np.array_equal(getattr(s, k).cpu().numpy().view(np.int32), st[k].view(np.int32)), f"frame {f}: {k}"
        assert np.array_equal(bits(hst), bits(st["ps_hst"])), f"frame {f}: smoothing history"
    assert np.abs(wr).max() > 100 and np.abs(wl - wr).max() > 10
