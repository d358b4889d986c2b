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


def test_fused_handovers(ctx, oracle):
    """SURVEY 8a-F around the eSBR stage: USAC core WORD32 x 2^-15 and legacy interleaved PCM16 hand-overs fused into the
    analysis bank's load, ixheaacd_samples_sat fused into the synthesis bank's store — against oracle glue + oracle banks"""
    import torch
    import libxaac_b200 as xb
    P = oracle_util.P
    rng = np.random.default_rng(12)
    n = 96
    _, st, pos = oracle_util.synth_esbr_anal_units(n, 21)
    # USAC core output
    core = (rng.standard_normal((n, 1024)) * 2.0 ** rng.integers(4, 28, (n, 1))).clip(-2**31, 2**31 - 1).astype(np.int64).astype(np.int32)
    core[0, :4] = [2**31 - 1, -2**31, 16777217, -33554435]
    xf = np.zeros((n, 1024), np.float32)
    oracle.lib.xo_esbr_core_to_float(P(core), P(xf), n * 1024)
    want = oracle.esbr_anal_batch(xf, st, pos)
    ea = xb.EsbrAnalBatch(n)
    ea.states.copy_(torch.from_numpy(st)); ea.pos.copy_(torch.from_numpy(pos))
    q, err = xb.esbr_analysis_filt_block(ctx, ea, torch.from_numpy(core).cuda())
    torch.cuda.synchronize()
    assert int(err.abs().max()) == 0
    assert np.array_equal(q.cpu().numpy().view(np.int32), want[0].view(np.int32))
    assert np.array_equal(ea.states.cpu().numpy(), want[1])
    # legacy stereo PCM16, interleaved
    pcm = rng.integers(-32768, 32768, (n // 2, 1024, 2)).astype(np.int16)
    for u in range(n):
        oracle.lib.xo_esbr_pcm16_to_float(P(np.ascontiguousarray(pcm[u // 2])), 2, u % 2, P(xf[u]), 1024)
    want = oracle.esbr_anal_batch(xf, st, pos)
    ea.states.copy_(torch.from_numpy(st)); ea.pos.copy_(torch.from_numpy(pos))
    q, err = xb.esbr_analysis_filt_block(ctx, ea, torch.from_numpy(pcm).cuda(), ch_fac=2)
    torch.cuda.synchronize()
    assert np.array_equal(q.cpu().numpy().view(np.int32), want[0].view(np.int32))
    # synthesis + samples_sat (amplitudes around and beyond full scale)
    qmf, fs, sp = oracle_util.synth_esbr_units(n, 31)
    qmf *= (2.0 ** rng.uniform(-2, 13, (n, 1, 1))).astype(np.float32)
    o, fs2, sp2 = oracle.esbr_synth_batch(qmf, fs, sp)
    wpcm = np.zeros((n // 2, 2048, 2), np.int16)
    for u in range(n):
        oracle.lib.xo_samples_sat16(P(np.ascontiguousarray(o[u])), 2, u % 2, P(wpcm[u // 2]), 2048)
    es = xb.EsbrSynthBatch(n)
    es.states.copy_(torch.from_numpy(fs)); es.pos.copy_(torch.from_numpy(sp))
    gp = torch.zeros((n // 2, 2048, 2), dtype=torch.int16, device="cuda")
    out, err = xb.esbr_synthesis_filt(ctx, es, torch.from_numpy(qmf).cuda(), pcm16=gp, ch_fac=2)
    torch.cuda.synchronize()
    assert np.array_equal(out.cpu().numpy().view(np.int32), o.view(np.int32))
    assert np.array_equal(gp.cpu().numpy(), wpcm)
    assert np.abs(wpcm.astype(np.int32)).max() > 30000   # the bank saturates in WORD32 first: |out| < 32768 always
    es.states.copy_(torch.from_numpy(fs)); es.pos.copy_(torch.from_numpy(sp))
    gp.zero_()
    out, err = xb.esbr_synthesis_filt(ctx, es, torch.from_numpy(qmf).cuda(), pcm16=gp, ch_fac=2, want_float=False)
    torch.cuda.synchronize()
    assert out is None and np.array_equal(gp.cpu().numpy(), wpcm) and np.array_equal(es.states.cpu().numpy(), fs2)
