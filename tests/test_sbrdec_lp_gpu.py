"""GPU parity tests for the fused low-power SBR stage (xaac_b200_sbr_dec_lp_dev: ixheaacd_sbr_dec with low_pow_flag = 1,
the fixed-point path of stereo HE-AACv1) against whole-stage records tapped from the compiled reference decoding a
real HE-AACv1 stereo stream, against the CPU oracle on perturbed units, and over multi-frame stereo streams with the
state resident in HBM."""
import os

import numpy as np
import pytest

from tests import oracle_util

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden", "sbrdec_lp_tapped.npz")


def run_gpu(ctx, side, st, tin, out_ch=1, frames=1, low_power=True):
    import torch
    import libxaac_b200 as xb
    n = st.shape[0]
    state = xb.SbrState(ctx, n, low_power=low_power)
    state.upload(st, None)
    outs = []
    for f in range(frames):
        s = side if frames == 1 else side[f]
        t = tin if frames == 1 else tin[f]
        out, err = xb.sbr_dec_lp(ctx, state, torch.from_numpy(np.ascontiguousarray(s)).cuda(),
                                 torch.from_numpy(np.ascontiguousarray(t)).cuda(), out_ch=out_ch)
        torch.cuda.synchronize()
        outs.append((out.cpu().numpy(), err.cpu().numpy()))
    st2, _ = state.download()
    state.close()
    return outs, st2


def check(u, got_st, got_out, got_err, exp_st, exp_out, exp_err, what):
    assert got_err == exp_err, f"{what} unit {u}: err {got_err} != {exp_err}"
    if exp_err:
        return
    if not np.array_equal(got_out, exp_out):
        raise AssertionError(f"{what} unit {u}: PCM differs at {np.argwhere(got_out != exp_out).ravel()[:10]}")
    if not np.array_equal(got_st, exp_st):
        raise AssertionError(f"{what} unit {u}: state differs at {np.argwhere(got_st != exp_st).ravel()[:10]}")


def test_golden_tapped_records(ctx):
    g = np.load(GOLD)
    outs, st2 = run_gpu(ctx, g["side"], g["st_in"], g["tin"])
    out, err = outs[0]
    for u in range(len(g["side"])):
        check(u, st2[u], out[u], err[u], g["st_out"][u], g["out_l"][u], g["hdr"][u][4], "golden")


def test_golden_with_plain_state_object(ctx):
    """the low-power stage also runs on a state created without XAAC_B200_SBR_STATE_LP"""
    g = np.load(GOLD)
    outs, st2 = run_gpu(ctx, g["side"][:8], g["st_in"][:8], g["tin"][:8], low_power=False)
    out, err = outs[0]
    for u in range(8):
        check(u, st2[u], out[u], err[u], g["st_out"][u], g["out_l"][u], g["hdr"][u][4], "plain state")


@pytest.mark.parametrize("seed,n", [(1, 5), (2, 320), (3, 2500)])
def test_random_units(ctx, oracle, seed, n):
    g = np.load(GOLD)
    side, st, tin = oracle_util.synth_sbr_lp_units(n, seed, g)
    outs, st2 = run_gpu(ctx, side, st, tin)
    out, err = outs[0]
    for u in range(n):
        es, eo, ee = oracle.sbr_dec_lp(side[u], st[u], tin[u])
        check(u, st2[u], out[u], err[u], es, eo, ee, f"seed {seed}")


def test_stereo_streams_state_resident(ctx, oracle):
    """6 frames of stereo streams (units 2k / 2k+1 = L / R of stream k, interleaved PCM out like the reference's time
    buffer), state staying on the device between frames."""
    g = np.load(GOLD)
    frames, n = 6, 48
    rng = np.random.default_rng(5)
    side = np.zeros((frames, n, 1232), np.int16)
    tin = np.zeros((frames, n, 1024), np.int16)
    st = np.zeros((n, 3920), np.int16)
    for u in range(n):
        ch = u % 2
        st[u] = g["st_in"][2 + ch]
        for f in range(frames):
            r = 2 + ch + 2 * f
            side[f, u] = g["side"][r]
            tin[f, u] = g["tin"][r] if u < 2 else np.clip(
                g["tin"][r].astype(np.int32) * rng.integers(1, 5) + rng.integers(-300, 300, 1024), -32768, 32767)
    outs, st2 = run_gpu(ctx, side, st, tin, out_ch=2, frames=frames)
    for u in range(n):
        s = st[u].copy()
        for f in range(frames):
            s, o, e = oracle.sbr_dec_lp(side[f, u], s, tin[f, u])
            out, err = outs[f]
            assert err[u] == e == 0
            assert np.array_equal(out[u // 2, :, u % 2], o), f"unit {u} frame {f}: PCM"
            if u < 2:
                assert np.array_equal(o, g["out_l"][2 + (u % 2) + 2 * f])
        assert np.array_equal(st2[u], s), f"unit {u}: final state differs at {np.argwhere(st2[u] != s).ravel()[:10]}"
