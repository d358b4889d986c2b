"""GPU parity tests for the QMF harmonic transposer (xaac_b200_esbr_hbe_apply_dev = batched ixheaacd_qmf_hbe_apply) against
records tapped from a real USAC decode with -harmonic_sbr:1, against the CPU oracle (pinned on the compiled reference by
tests/test_oracle_hbe.py) and, where oracle/_ref is present, against the compiled reference function itself.  Float results
are compared as bit patterns.  The only libm call on the path is cbrt (stretch-3 bands): both sides round a <= 1 ulp double
result to float, so a difference needs the exact value within ~2^-29 of a float rounding boundary — the stretch-3 comparisons
therefore allow a 1e-5 fraction of cells to differ by one float ulp (none observed)."""
import os

import numpy as np
import pytest

from tests import oracle_util

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden", "esbr_hbe_tapped.npz")


def bits(a):
    return np.ascontiguousarray(a, np.float32).view(np.int32)


def run_gpu(ctx, cfg, state, qre, qim, pv_re=None, pv_im=None):
    import torch
    import libxaac_b200 as xb
    n = len(cfg)
    hb = xb.EsbrHbeBatch(n)
    hb.state.copy_(torch.from_numpy(np.ascontiguousarray(state, np.float32)))
    pr = torch.zeros((n, 32, 64), dtype=torch.float32, device="cuda") if pv_re is None else torch.from_numpy(np.ascontiguousarray(pv_re)).cuda()
    pi = torch.zeros((n, 32, 64), dtype=torch.float32, device="cuda") if pv_im is None else torch.from_numpy(np.ascontiguousarray(pv_im)).cuda()
    err = xb.esbr_qmf_hbe_apply(ctx, hb, torch.from_numpy(np.ascontiguousarray(qre, np.float32)).cuda(),
                                torch.from_numpy(np.ascontiguousarray(qim, np.float32)).cuda(), pr, pi,
                                torch.from_numpy(np.ascontiguousarray(cfg, np.int32)).cuda())
    torch.cuda.synchronize()
    return pr.cpu().numpy(), pi.cpu().numpy(), hb.state.cpu().numpy(), err.cpu().numpy()


def check(cfg, got, want, what, exact=True):
    pr, pi, st, err = got
    er, ei, es, ee = want
    assert np.array_equal(err, ee), (what, err[:8], ee[:8])
    bad = 0
    cells = 0
    for u in range(len(cfg)):
        b0, b1 = cfg[u, 2], cfg[u, 3]
        for a, b, nm in ((pr[u][:, b0:b1], er[u][:, b0:b1], "pv_re"), (pi[u][:, b0:b1], ei[u][:, b0:b1], "pv_im"), (st[u], es[u], "state")):
            d = bits(a) != bits(b)
            cells += d.size
            if d.any():
                if exact:
                    raise AssertionError(f"{what}: unit {u} cfg {cfg[u].tolist()}: {nm} differs at {np.argwhere(d)[:6].tolist()}")
                assert np.abs(bits(a).astype(np.int64) - bits(b).astype(np.int64))[d].max() <= 1, f"{what}: unit {u} {nm} differs by more than one ulp"
                bad += int(d.sum())
    assert bad <= 1e-5 * cells, (what, bad, cells)


def test_tapped_records(ctx):
    g = np.load(GOLD)
    got = run_gpu(ctx, g["cfg"], g["state_in"], g["qmf_re"], g["qmf_im"], g["pv_re"], g["pv_im"])
    check(g["cfg"], got, (g["pv_re"], g["pv_im"], g["state_out"], g["ret"]), "tapped")


def test_tapped_stream_state_resident(ctx):
    """4 consecutive frames of both channels with the transposer state kept on the device"""
    import torch
    import libxaac_b200 as xb
    g = np.load(GOLD)
    r0, cnt = g["run"]
    hb = xb.EsbrHbeBatch(2)
    hb.state.copy_(torch.from_numpy(g["state_in"][r0:r0 + 2]))
    for f in range(cnt // 2):
        i = r0 + 2 * f
        pr = torch.zeros((2, 32, 64), dtype=torch.float32, device="cuda")
        pi = torch.zeros_like(pr)
        err = xb.esbr_qmf_hbe_apply(ctx, hb, torch.from_numpy(g["qmf_re"][i:i + 2]).cuda(), torch.from_numpy(g["qmf_im"][i:i + 2]).cuda(),
                                    pr, pi, torch.from_numpy(g["cfg"][i:i + 2]).cuda())
        torch.cuda.synchronize()
        check(g["cfg"][i:i + 2], (pr.cpu().numpy(), pi.cpu().numpy(), hb.state.cpu().numpy(), err.cpu().numpy()),
              (g["pv_re"][i:i + 2], g["pv_im"][i:i + 2], g["state_out"][i:i + 2], g["ret"][i:i + 2]), f"frame {f}")


@pytest.mark.parametrize("mode,seed,n", [("zero", 1, 400), ("pitch", 2, 400), ("mixed", 3, 1000)])
def test_units_vs_oracle_and_reference(ctx, oracle, ref, mode, seed, n):
    cfg, tbl, state, qre, qim = oracle_util.synth_hbe_units(n, seed, ref, mode)
    got = run_gpu(ctx, cfg, state, qre, qim)
    want = oracle_util.oracle_hbe_batch(oracle, cfg, state, qre, qim)
    exact = cfg[:, 4] < 3
    check(cfg[exact], tuple(x[exact] for x in got), tuple(x[exact] for x in want), f"{mode}/oracle/stretch2")
    check(cfg[~exact], tuple(x[~exact] for x in got), tuple(x[~exact] for x in want), f"{mode}/oracle/stretch3+", exact=False)
    wr = oracle_util.ref_hbe_batch(ref, cfg, state, qre, qim, tbl)
    check(cfg[exact], tuple(x[exact] for x in got), tuple(x[exact] for x in wr), f"{mode}/reference/stretch2")
    check(cfg[~exact], tuple(x[~exact] for x in got), tuple(x[~exact] for x in wr), f"{mode}/reference/stretch3+", exact=False)


def test_unsupported_and_failing_configurations(ctx, ref):
    cfg, tbl, state, qre, qim = oracle_util.synth_hbe_units(6, 5, ref, "zero")
    cfg[0, 6] = 1          # 4:1 system
    cfg[1, 0] = 24         # bank size the reference has no tables for
    cfg[2, 1] = -1         # k_start < 0: the reference's own -1
    cfg[3, 4], cfg[3, 10] = 4, 1   # x_over_qmf[2] <= 1 with stretch 4: IA_FATAL_ERROR
    _, _, st, err = run_gpu(ctx, cfg, state, qre, qim)
    assert err[0] == -2 and err[1] == -2 and err[2] == -1 and err[3] == np.int32(-2147483648) and err[4] == 0 and err[5] == 0
    assert np.array_equal(bits(st[:4]), bits(state[:4])), "refused units must leave their state untouched"


def test_baseline_batch_tiled_from_tapped_records(ctx):
    """131 072 channel units (BASELINE configs[4]: 65 536 stereo streams): every tiled copy reproduces its tapped record"""
    import torch
    import libxaac_b200 as xb
    g = np.load(GOLD)
    k = len(g["cfg"])
    n = 131072
    idx = np.arange(n) % k
    hb = xb.EsbrHbeBatch(n)
    hb.state.copy_(torch.from_numpy(g["state_in"])[torch.from_numpy(idx)])
    sel = torch.from_numpy(idx).cuda()
    qre, qim = torch.from_numpy(g["qmf_re"]).cuda()[sel], torch.from_numpy(g["qmf_im"]).cuda()[sel]
    pr = torch.zeros((n, 32, 64), dtype=torch.float32, device="cuda")
    pi = torch.zeros_like(pr)
    err = xb.esbr_qmf_hbe_apply(ctx, hb, qre, qim, pr, pi, torch.from_numpy(g["cfg"]).cuda()[sel])
    torch.cuda.synchronize()
    assert int(err.abs().max().item()) == 0
    want_r, want_i, want_s = torch.from_numpy(g["pv_re"]).cuda()[sel], torch.from_numpy(g["pv_im"]).cuda()[sel], torch.from_numpy(g["state_out"]).cuda()[sel]
    assert torch.equal(pr[:, :, 32:].view(torch.int32), want_r[:, :, 32:].view(torch.int32))
    assert torch.equal(pi[:, :, 32:].view(torch.int32), want_i[:, :, 32:].view(torch.int32))
    assert torch.equal(hb.state.view(torch.int32), want_s.view(torch.int32))


def test_stage_with_hbe_golden_stream(ctx):
    """the whole eSBR stage with the transposer (xaac_b200_esbr_dec_hbe_dev, five launches, state resident) against 6
    consecutive frames x 2 channels tapped around ixheaacd_sbr_dec in a real -harmonic_sbr:1 decode, bit for bit"""
    import torch
    import libxaac_b200 as xb
    from tests.test_oracle_esbr import esbr_stage_golden_frames
    g = np.load(os.path.join(os.path.dirname(GOLD), "esbr_hbe_stage_tapped.npz"))
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
    s = xb.EsbrDecHbeBatch(2)
    for k in oracle_util.ESH_KEYS:
        getattr(s, k).copy_(t(g["in0_" + k]))
    for f, r, rg in esbr_stage_golden_frames(g):
        ipar = t(g["ec_ipar_in"][r])
        pcm = torch.zeros((1, 2048, 2), dtype=torch.int16, device="cuda")
        out, err = xb.esbr_dec_hbe(ctx, s, t(g["time_in"][r]), t(g["hbe_cfg"][r]), t(g["hf_par"][r]), ipar, t(g["ec_fpar"][r]), t(rg),
                                   pcm16=pcm, ch_fac=2)
        torch.cuda.synchronize()
        assert int(err.abs().max()) == 0, f"frame {f}: {err.cpu().numpy()}"
        want = g["time_out"][r]
        assert np.array_equal(bits(out.cpu().numpy()), bits(want)), f"frame {f}: time output"
        assert np.array_equal(pcm.cpu().numpy()[0], np.trunc(np.clip(want, -32768, 32767)).astype(np.int16).T), f"frame {f}: PCM16"
        assert np.array_equal(ipar.cpu().numpy(), g["ec_ipar_out"][r]), f"frame {f}: in/out parameter words"
        for k in ("anal_states", "anal_pos", "synth_states", "synth_pos", "bw_prev", "patch", "ec_state", "hbe_state"):
            assert np.array_equal(getattr(s, k).cpu().numpy().view(np.int32), g["out_" + k][r].view(np.int32)), f"frame {f}: {k}"
    for k in ("qmf_re", "qmf_im", "out_re", "out_im", "pv_re", "pv_im"):
        assert np.array_equal(getattr(s, k).cpu().numpy().view(np.int32), g["out_" + k].view(np.int32)), k


def test_stage_with_hbe_full_batch_tiling(ctx):
    """BASELINE configs[4] batch (131 072 channel units = 65 536 stereo streams), two frames with the state carried: every
    tiled copy reproduces the tapped records"""
    import torch
    import libxaac_b200 as xb
    from tests.test_oracle_esbr import esbr_stage_golden_frames
    g = np.load(os.path.join(os.path.dirname(GOLD), "esbr_hbe_stage_tapped.npz"))
    n = 131072
    rep = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda().repeat((n // 2,) + (1,) * (a.ndim - 1)).contiguous()
    s = xb.EsbrDecHbeBatch(n)
    for k in oracle_util.ESH_KEYS:
        getattr(s, k).copy_(rep(g["in0_" + k]))
    pcm = torch.zeros((n // 2, 2048, 2), dtype=torch.int16, device="cuda")
    for f, r, rg in esbr_stage_golden_frames(g):
        if f >= 2:
            break
        ipar = rep(g["ec_ipar_in"][r])
        out, err = xb.esbr_dec_hbe(ctx, s, rep(g["time_in"][r]), rep(g["hbe_cfg"][r]), rep(g["hf_par"][r]), ipar, rep(g["ec_fpar"][r]),
                                   rep(rg), pcm16=pcm, ch_fac=2)
        torch.cuda.synchronize()
        assert int(err.abs().max()) == 0
        want = torch.from_numpy(np.ascontiguousarray(g["time_out"][r])).cuda()
        assert torch.equal(out.view(torch.int32).view(n // 2, 2, 2048), want.view(torch.int32).unsqueeze(0).expand(n // 2, 2, 2048)), f"frame {f}"
        assert torch.equal(pcm, pcm[:1].expand_as(pcm))
        for k in ("anal_states", "synth_states", "ec_state", "bw_prev", "hbe_state"):
            w = torch.from_numpy(np.ascontiguousarray(g["out_" + k][r])).cuda()
            tt = getattr(s, k)
            assert torch.equal(tt.view(torch.int32).view(n // 2, 2, -1), w.view(torch.int32).unsqueeze(0).expand(n // 2, 2, w.shape[-1])), f"frame {f}: {k}"


def test_stage_in_two_halves_equals_the_single_call(ctx):
    """xaac_b200_esbr_dec_front_dev + _back_dev (the split a host uses to rebuild the limiter tables between the HF generator and
    the envelope adjuster) against the tapped stream, bit for bit; and a frame flagged as 'patching mode changed' is refused
    without XAAC_EEC_LIM_REBUILT and accepted with it"""
    import torch
    import libxaac_b200 as xb
    from tests.test_oracle_esbr import esbr_stage_golden_frames
    g = np.load(os.path.join(os.path.dirname(GOLD), "esbr_hbe_stage_tapped.npz"))
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
    s = xb.EsbrDecHbeBatch(2)
    for k in oracle_util.ESH_KEYS:
        getattr(s, k).copy_(t(g["in0_" + k]))
    for f, r, rg in esbr_stage_golden_frames(g):
        ip = np.ascontiguousarray(g["ec_ipar_in"][r]).copy()
        err = torch.zeros((6, 2), dtype=torch.int32, device="cuda")
        xb.esbr_dec_front(ctx, s, t(g["time_in"][r]), t(g["hbe_cfg"][r]), t(g["hf_par"][r]), err)
        if f == 2:  # what the drop-in does on such a frame: same tables, flagged as rebuilt
            probe = ip.copy()
            probe[:, 19] = 1
            saved = {k: getattr(s, k).clone() for k in ("synth_states", "synth_pos", "ec_state")}
            _, _, e2 = xb.esbr_dec_back(ctx, s, t(probe), t(g["ec_fpar"][r]), t(rg), err.clone())
            torch.cuda.synchronize()
            assert (e2.cpu().numpy()[2] == -2).all()
            for k, v in saved.items():
                getattr(s, k).copy_(v)
            ip[:, 19] = 1
            ip[:, 20] = 1
        ipar = t(ip)
        out, _, err = xb.esbr_dec_back(ctx, s, ipar, t(g["ec_fpar"][r]), t(rg), err)
        torch.cuda.synchronize()
        assert int(err.abs().max()) == 0, f"frame {f}: {err.cpu().numpy()}"
        assert np.array_equal(bits(out.cpu().numpy()), bits(g["time_out"][r])), f"frame {f}: time output"
        for k in ("anal_states", "synth_states", "bw_prev", "patch", "ec_state", "hbe_state"):
            assert np.array_equal(getattr(s, k).cpu().numpy().view(np.int32), g["out_" + k][r].view(np.int32)), f"frame {f}: {k}"
