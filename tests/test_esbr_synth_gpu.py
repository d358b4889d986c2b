"""GPU parity tests for the eSBR 64-band synthesis bank (xaac_b200_esbr_synth64_dev) against the CPU oracle (pinned to
the compiled reference's leaf functions by tests/test_oracle_esbr.py): float output compared bit for bit, state and
positions included; lock-step and arbitrary ring phases; streams with the state resident in HBM."""
import numpy as np
import pytest

from tests import oracle_util

pytestmark = pytest.mark.gpu


def run_gpu(ctx, qmf, fs, pos, frames=1):
    import torch
    import libxaac_b200 as xb
    n = fs.shape[0]
    st = xb.EsbrSynthBatch(n)
    st.states.copy_(torch.from_numpy(fs))
    st.pos.copy_(torch.from_numpy(pos))
    outs = []
    for f in range(frames):
        q = qmf if frames == 1 else qmf[f]
        out, err = xb.esbr_synthesis_filt(ctx, st, torch.from_numpy(np.ascontiguousarray(q)).cuda())
        torch.cuda.synchronize()
        assert int(err.abs().max().item()) == 0
        outs.append(out.cpu().numpy())
    return outs, st.states.cpu().numpy(), st.pos.cpu().numpy()


@pytest.mark.parametrize("seed,n", [(1, 5), (2, 300), (3, 1500)])
def test_units_vs_oracle(ctx, oracle, seed, n):
    qmf, fs, pos = oracle_util.synth_esbr_units(n, seed)
    outs, fs2, pos2 = run_gpu(ctx, qmf, fs, pos)
    eo, ef, ep = oracle.esbr_synth_batch(qmf, fs, pos)
    assert np.array_equal(pos2, ep)
    for u in range(n):
        if not np.array_equal(outs[0][u].view(np.int32), eo[u].view(np.int32)):
            raise AssertionError(f"unit {u}: output differs at {np.argwhere(outs[0][u] != eo[u]).ravel()[:8]}")
        if not np.array_equal(fs2[u], ef[u]):
            raise AssertionError(f"unit {u}: state differs at {np.argwhere(fs2[u] != ef[u]).ravel()[:8]}")


def test_non_lockstep_positions(ctx, oracle):
    n = 200
    qmf, fs, pos = oracle_util.synth_esbr_units(n, 7)
    rng = np.random.default_rng(7)
    pos[:, 0] = 128 * rng.integers(0, 10, n)
    pos[:, 1] = 64 * rng.integers(0, 10, n)
    outs, fs2, pos2 = run_gpu(ctx, qmf, fs, pos)
    eo, ef, ep = oracle.esbr_synth_batch(qmf, fs, pos)
    assert np.array_equal(outs[0].view(np.int32), eo.view(np.int32)) and np.array_equal(fs2, ef) and np.array_equal(pos2, ep)


def test_streams_state_resident(ctx, oracle):
    n, frames = 64, 5
    qs = np.stack([oracle_util.synth_esbr_units(n, 40 + f)[0] for f in range(frames)])
    fs = np.zeros((n, 1280), np.int32)
    pos = np.zeros((n, 2), np.int32)
    outs, fs2, pos2 = run_gpu(ctx, qs, fs, pos, frames=frames)
    for f in range(frames):
        eo, fs, pos = oracle.esbr_synth_batch(qs[f], fs, pos)
        assert np.array_equal(outs[f].view(np.int32), eo.view(np.int32)), f"frame {f}"
    assert np.array_equal(fs2, fs) and np.array_equal(pos2, pos)


def run_anal(ctx, x, st, pos):
    import torch
    import libxaac_b200 as xb
    n = x.shape[0]
    a = xb.EsbrAnalBatch(n)
    a.states.copy_(torch.from_numpy(st))
    a.pos.copy_(torch.from_numpy(pos))
    qmf, err = xb.esbr_analysis_filt_block(ctx, a, torch.from_numpy(x).cuda())
    torch.cuda.synchronize()
    assert int(err.abs().max().item()) == 0
    return qmf.cpu().numpy(), a.states.cpu().numpy(), a.pos.cpu().numpy()


@pytest.mark.parametrize("seed,n,lock", [(1, 6, True), (2, 700, True), (3, 300, False)])
def test_analysis_vs_oracle(ctx, oracle, seed, n, lock):
    x, st, pos = oracle_util.synth_esbr_anal_units(n, seed)
    if not lock:
        rng = np.random.default_rng(seed)
        pos[:, 0] = 32 * rng.integers(0, 10, n)
        pos[:, 1] = 64 * rng.integers(0, 10, n)
    q, s2, p2 = run_anal(ctx, x, st, pos)
    eq, es, ep = oracle.esbr_anal_batch(x, st, pos)
    assert np.array_equal(p2, ep)
    for u in range(n):
        if not np.array_equal(q[u].view(np.int32), eq[u].view(np.int32)):
            raise AssertionError(f"unit {u}: QMF output differs at {np.argwhere(q[u] != eq[u])[:4].tolist()}")
        assert np.array_equal(s2[u], es[u]), f"unit {u}: ring state"


def test_analysis_synthesis_stream(ctx, oracle):
    """analysis -> synthesis on the device over 5 frames, both states resident; the low band passes straight through"""
    import torch
    import libxaac_b200 as xb
    n, frames = 48, 5
    a = xb.EsbrAnalBatch(n)
    s = xb.EsbrSynthBatch(n)
    st = np.zeros((n, 320), np.int32)
    pos = np.zeros((n, 2), np.int32)
    fs = np.zeros((n, 1280), np.int32)
    sp = np.zeros((n, 2), np.int32)
    for f in range(frames):
        x, _, _ = oracle_util.synth_esbr_anal_units(n, 60 + f)
        qmf, _ = xb.esbr_analysis_filt_block(ctx, a, torch.from_numpy(x).cuda())
        out, _ = xb.esbr_synthesis_filt(ctx, s, qmf)
        torch.cuda.synchronize()
        eq, st, pos = oracle.esbr_anal_batch(x, st, pos)
        eo, fs, sp = oracle.esbr_synth_batch(eq, fs, sp)
        assert np.array_equal(out.cpu().numpy().view(np.int32), eo.view(np.int32)), f"frame {f}"
    assert np.array_equal(a.states.cpu().numpy(), st) and np.array_equal(s.states.cpu().numpy(), fs)
