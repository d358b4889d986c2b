"""CPU tests for the float parametric stereo hand-over (ixheaacd_esbr_apply_ps): the drop-in's host-side parameter
preparation (libxaac_b200/dropin/ixheaacd_b200_pack_ps_flt.h — the mixing matrices h11..h22 with their IPD / OPD smoothing, and
the history they advance) against what the COMPILED reference function leaves in its own instance, and the committed golden
records (tests/golden/esbr_ps_ref.npz, made by tools/make_golden.py fps from the compiled reference) against a fresh run."""
import os

import numpy as np

from tests import oracle_util as ou

GOLD = os.path.join(os.path.dirname(__file__), "golden", "esbr_ps_ref.npz")


def bits(a):
    return np.ascontiguousarray(a, np.float32).view(np.int32)


def golden_inputs(g, f):
    sh = g["shift"][:, None, None]
    return (np.ldexp(g["low_re_i16"][f].astype(np.float32), sh).astype(np.float32),
            np.ldexp(g["low_im_i16"][f].astype(np.float32), sh).astype(np.float32))


def test_host_side_matches_reference_instance_over_frames(ref):
    rng = np.random.default_rng(5)
    n = 48
    st, hst = ou.fps_fresh_state(n)
    for f in range(5):
        lr, li, par = ou.synth_fps_frame(n, rng)
        r = ou.ref_fps_batch(ref, lr, li, par, st, hst)
        assert r["rc"] == 0
        # the history the host code would commit == what ixheaacd_esbr_ps_apply_rotation left in the struct
        assert np.array_equal(bits(r["commit"]), bits(r["hst"])), f
        # set 0 of the side record = the h*_prev the frame started from; the last envelope's set = the new h*_prev
        side = r["side"]
        assert np.array_equal(bits(side[:, 16:176]), bits(hst[:, :160]))
        for u in range(n):
            ne = int(par[u, 0])
            assert np.array_equal(bits(side[u, 16 + 160 * ne:176 + 160 * ne]), bits(r["hst"][u, :160]))
            assert side[u].view(np.int32)[0] == ne and side[u].view(np.int32)[7] == par[u, 7]
        st, hst = r["state"], r["hst"]
    assert np.abs(hst[:, 80:160]).max() > 0.1  # IPD / OPD rotation exercised (imaginary parts)


def test_host_side_refuses_frames_outside_the_subset(ref):
    rng = np.random.default_rng(6)
    st, hst = ou.fps_fresh_state(4)
    lr, li, par = ou.synth_fps_frame(4, rng, good_borders=False)
    assert ou.ref_fps_batch(ref, lr, li, par, st, hst)["rc"] == -1


def test_golden_records_reproduce(ref):
    g = np.load(GOLD)
    n = g["par"].shape[1]
    st, hst = ou.fps_fresh_state(n)
    for f in range(g["par"].shape[0]):
        lr, li = golden_inputs(g, f)
        r = ou.ref_fps_batch(ref, lr, li, g["par"][f], st, hst)
        for k, key in (("left", "left"), ("right", "right"), ("state", "state_out"), ("hst", "hst_out"), ("side", "side")):
            assert np.array_equal(bits(r[k]), bits(g[key][f])), (f, k)
        st, hst = r["state"], r["hst"]


def test_golden_is_not_trivial():
    g = np.load(GOLD)
    assert np.abs(g["right"]).max() > 10.0 and np.abs(g["left"]).max() > 10.0
    assert set(g["par"][:, :, 0].ravel().tolist()) >= {1, 2}
    assert (g["state_out"][-1][:, 4300:4304].view(np.int32) >= 0).all()
