"""GPU parity tests for the fixed-point HQ 32-band QMF analysis kernel (C-ABI) against the CPU oracle. Bit-exact."""
import numpy as np
import pytest

from tests import oracle_util

pytestmark = pytest.mark.gpu


def synth_inputs(n, seed):
    rng = np.random.default_rng(seed)
    s = rng.integers(2, 16, size=(n, 1))
    tin = ((rng.random((n, 1024)) * 2 - 1) * (2.0 ** s)).astype(np.int16)
    st = rng.integers(-32768, 32768, (n, 320)).astype(np.int16)
    pos = np.stack([rng.integers(0, 10, n) * 32, rng.integers(0, 5, n) * 128], 1).astype(np.int16)
    usb = rng.integers(0, 33, n).astype(np.int16)
    if n >= 6:
        tin[0] = 32767
        tin[1] = -32768
        tin[2] = np.where(np.arange(1024) % 2 == 0, 32767, -32768)
        st[0] = 32767
        st[1] = -32768
        tin[3] = 0
        st[3] = 0
        usb[4] = 32
        usb[5] = 0
    return tin, st, pos, usb


def run_gpu(ctx, tin, st, pos, usb, ch_fac=1):
    import torch
    import libxaac_b200 as xb
    n = st.shape[0]
    state = xb.QmfAnalBatch(n)
    state.states.copy_(torch.from_numpy(st))
    state.pos.copy_(torch.from_numpy(pos))
    m = xb.cplx_anal_qmffilt(ctx, state, torch.from_numpy(tin).cuda(), torch.from_numpy(usb).cuda(), ch_fac=ch_fac)
    torch.cuda.synchronize()
    return m.cpu().numpy(), state.states.cpu().numpy(), state.pos.cpu().numpy()


@pytest.mark.parametrize("seed,n", [(1, 1), (2, 9), (3, 500), (4, 5000)])
def test_random_units(ctx, oracle, seed, n):
    tin, st, pos, usb = synth_inputs(n, seed)
    g = run_gpu(ctx, tin, st, pos, usb)
    e = oracle.anal_batch(tin, st, pos, usb)
    for a, b, nm in zip(g, e, ("matrix", "states", "pos")):
        assert np.array_equal(a, b), f"{nm}: first mismatch at {np.argwhere(a != b)[:3]}"


def test_interleaved_input(ctx, oracle):
    tin, st, pos, usb = synth_inputs(16, 7)
    inter = np.ascontiguousarray(np.stack([tin[0::2], tin[1::2]], axis=2))  # [8,1024,2]
    g = run_gpu(ctx, inter, st, pos, usb, ch_fac=2)
    e = oracle.anal_batch(tin, st, pos, usb)
    assert np.array_equal(g[0], e[0]) and np.array_equal(g[1], e[1]) and np.array_equal(g[2], e[2])


def test_analysis_synthesis_chain(ctx, oracle):
    """analysis -> synthesis on the device over 5 frames, every frame equal to the oracle chain"""
    import torch
    import libxaac_b200 as xb
    n = 40
    rng = np.random.default_rng(12)
    a_state = xb.QmfAnalBatch(n)
    s_state = xb.QmfSynthBatch(n)
    a_st = np.zeros((n, 320), np.int16); a_pos = np.zeros((n, 2), np.int16)
    s_fs = np.zeros((n, 1280), np.int16); s_pos = np.zeros((n, 2), np.int16)
    params = np.tile(np.array([[-8, -8, -8, -6, 32, 32, 6, 0]], np.int16), (n, 1))
    usb = np.full(n, 32, np.int16)
    for f in range(5):
        t = np.arange(1024)[None, :] + 1024 * f
        freq = 0.002 + 0.004 * np.arange(n)[:, None]
        tin = (6000 * np.sin(2 * np.pi * freq * t) + 300 * rng.standard_normal((n, 1024))).astype(np.int16)
        m = xb.cplx_anal_qmffilt(ctx, a_state, torch.from_numpy(tin).cuda(), torch.from_numpy(usb).cuda())
        pcm = xb.cplx_synt_qmffilt(ctx, s_state, m, torch.from_numpy(params).cuda())
        e_m, a_st, a_pos = oracle.anal_batch(tin, a_st, a_pos, usb)
        e_pcm, s_fs, s_pos = oracle.synth_batch(e_m, s_fs, s_pos, params)
        assert np.array_equal(m.cpu().numpy(), e_m), f"frame {f} matrix"
        assert np.array_equal(pcm.cpu().numpy(), e_pcm), f"frame {f} pcm"
    assert np.abs(e_pcm).max() > 500
