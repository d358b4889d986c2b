"""GPU parity tests for the whole fixed-point HQ SBR stage (xaac_b200_sbr_dec_hq_dev: 9 kernels incl. parametric stereo)
against whole-stage records tapped from the compiled reference, and against the CPU oracle on perturbed units and on
multi-frame streams (state kept resident in HBM between frames)."""
import os

import numpy as np
import pytest

from tests import oracle_util

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden", "sbrdec_tapped.npz")
NDSP = oracle_util.PS_ST_DSP_WORDS


def run_gpu(ctx, side, st, ps, tin, with_ps=True, frames=1):
    """side/tin may be [frames, n, ...] for a stream test; returns per-frame outputs and the final state"""
    import torch
    import libxaac_b200 as xb
    n = st.shape[0]
    state = xb.SbrState(ctx, n, with_ps=with_ps)
    state.upload(st, ps if with_ps else None)
    outs = []
    for f in range(frames):
        s = side if frames == 1 else side[f]
        t = tin if frames == 1 else tin[f]
        out, err = xb.sbr_dec(ctx, state, torch.from_numpy(np.ascontiguousarray(s)).cuda(),
                              torch.from_numpy(np.ascontiguousarray(t)).cuda())
        torch.cuda.synchronize()
        outs.append((out.cpu().numpy(), err.cpu().numpy()))
    st2, ps2 = state.download()
    state.close()
    return outs, st2, ps2


def compare_unit(u, out, err, st2, ps2, exp, ps_active, with_ps, what):
    est, eps, eol, eorr, eerr = exp
    assert err[u] == eerr, f"{what} unit {u}: err"
    if not np.array_equal(st2[u], est):
        raise AssertionError(f"{what} unit {u}: state differs at {np.argwhere(st2[u] != est).ravel()[:10]}")
    left = out[u][:, 0] if with_ps else out[u]
    if not np.array_equal(left, eol):
        raise AssertionError(f"{what} unit {u}: left PCM differs at {np.argwhere(left != eol).ravel()[:10]}")
    if ps_active:
        if not np.array_equal(ps2[u], eps):
            raise AssertionError(f"{what} unit {u}: PS state differs at {np.argwhere(ps2[u] != eps).ravel()[:10]}")
        assert np.array_equal(out[u][:, 1], eorr), f"{what} unit {u}: right PCM differs"
    elif with_ps:
        assert np.array_equal(ps2[u], eps), f"{what} unit {u}: PS state of a non-PS frame was modified"


def test_golden_tapped_records(ctx):
    g = np.load(GOLD)
    outs, st2, ps2 = run_gpu(ctx, g["side"], g["st_in"], g["ps_in"], g["tin"])
    out, err = outs[0]
    for u in range(len(g["side"])):
        psa = g["side"][u, 737] != 0
        exp = (g["st_out"][u], g["ps_out"][u] if psa else g["ps_in"][u], g["out_l"][u], g["out_r"][u], g["hdr"][u][4])
        compare_unit(u, out, err, st2, ps2, exp, psa, True, "golden")


def test_golden_mono_state_without_ps(ctx):
    g = np.load(GOLD)
    sel = np.nonzero(g["side"][:, 737] == 0)[0]
    outs, st2, _ = run_gpu(ctx, g["side"][sel], g["st_in"][sel], None, g["tin"][sel], with_ps=False)
    out, err = outs[0]
    for k, u in enumerate(sel):
        compare_unit(k, out, err, st2, None, (g["st_out"][u], None, g["out_l"][u], None, 0), False, False, "mono")


@pytest.mark.parametrize("seed,n", [(1, 3), (2, 200), (3, 1500)])
def test_random_units(ctx, oracle, seed, n):
    g = np.load(GOLD)
    side, st, ps, tin = oracle_util.synth_sbr_units(n, seed, g)
    outs, st2, ps2 = run_gpu(ctx, side, st, ps, tin)
    out, err = outs[0]
    est, eps, eol, eorr, eerr = oracle.sbr_dec_batch(side, st, ps, tin)
    for u in range(n):
        compare_unit(u, out, err, st2, ps2, (est[u], eps[u], eol[u], eorr[u], eerr[u]), side[u, 737] != 0, True,
                     f"seed {seed}")


def test_streams_state_resident(ctx, oracle):
    """8 frames per stream with the state staying on the device: every frame's PCM and the final state must match the
    oracle fed with its own state (and, for the tapped streams, the reference's PCM)."""
    g = np.load(GOLD)
    frames = 8
    bases = [1, int(np.argmin(g["side"][:, 737])) + 1]
    n = 64
    rng = np.random.default_rng(11)
    side = np.zeros((frames, n, 1232), np.int16)
    tin = np.zeros((frames, n, 1024), np.int16)
    st = np.zeros((n, 3920), np.int16)
    ps = np.zeros((n, 3888), np.int16)
    for u in range(n):
        b = bases[u % 2]
        st[u], ps[u] = g["st_in"][b], g["ps_in"][b]
        for f in range(frames):
            side[f, u] = g["side"][b + f]
            tin[f, u] = g["tin"][b + f] if u < 2 else np.clip(
                g["tin"][b + f].astype(np.int32) * rng.integers(1, 6) + rng.integers(-200, 200, 1024), -32768, 32767)
    outs, st2, ps2 = run_gpu(ctx, side, st, ps, tin, frames=frames)
    for u in range(n):
        s, p = st[u].copy(), ps[u].copy()
        for f in range(frames):
            s, p, ol, orr, e = oracle.sbr_dec(side[f, u], s, p, tin[f, u])
            out, err = outs[f]
            assert err[u] == e
            assert np.array_equal(out[u][:, 0], ol), f"stream {u} frame {f}: left PCM"
            if side[f, u, 737]:
                assert np.array_equal(out[u][:, 1], orr), f"stream {u} frame {f}: right PCM"
            if u < 2:
                assert np.array_equal(ol, g["out_l"][bases[u % 2] + f])
        assert np.array_equal(st2[u], s) and np.array_equal(ps2[u], p), f"stream {u}: final state"
