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


def test_heaac_lp_frame_host_matches_chained_oracles(ctx, oracle):
    """stereo HE-AACv1 streams through the host-buffer entry point (IMDCT -> PCM16 hand-over -> fused LP stage), 3 frames,
    n = 4400 channel units (> one 4096-unit chunk), against the chained CPU oracles; block-switching walk included"""
    import torch
    import libxaac_b200 as xb
    g = np.load(GOLD)
    n, frames = 4400, 3
    rng = np.random.default_rng(33)
    ist = xb.ImdctHostState(ctx, n)
    sst = xb.SbrState(ctx, n, low_power=True)
    u = np.arange(n)
    st = np.ascontiguousarray(g["st_in"][2 + (u & 1)])
    sst.upload(st, None)
    ovl = np.zeros((n, 512), np.int32)
    wstate = np.zeros((n, 2), np.uint8)
    check = np.concatenate([np.arange(0, 40), np.arange(4090, 4130), np.arange(n - 20, n)])
    for f in range(frames):
        s = rng.integers(10, 22, (n, 1))
        spec = ((rng.random((n, 1024)) * 2 - 1) * 2.0 ** s).astype(np.int64).astype(np.int32)
        ics = np.zeros((n, 2), np.uint8)
        ics[:, 1] = rng.integers(0, 2, n)
        if f == 1:
            ics[::7, 0] = 1  # long -> start
        if f == 2:
            ics[::7, 0] = 3  # start -> stop
        side = np.ascontiguousarray(g["side"][2 + (u & 1) + 2 * (((u >> 1) + f) % 12)])
        pcm = torch.zeros((n // 2, 2048, 2), dtype=torch.int16)
        err = torch.zeros((n,), dtype=torch.int32)
        xb.heaac_lp_frame_host(ctx, ist, sst, torch.from_numpy(spec), torch.from_numpy(ics), torch.from_numpy(side), pcm, 2,
                               err)
        out32, ovl, wstate, adj = oracle.imdct_batch(spec, ovl, wstate, ics)
        p16 = oracle.imdct_out_to_pcm16(out32, adj, 0)
        assert int(np.abs(err.numpy()).max()) == 0
        got = pcm.numpy()
        for k in check:
            st[k], o, e = oracle.sbr_dec_lp(side[k], st[k], p16[k])
            assert e == 0
            assert np.array_equal(got[k // 2, :, k % 2], o), f"frame {f} unit {k}: PCM"
    st2, _ = sst.download()
    for k in check:
        assert np.array_equal(st2[k], st[k]), f"unit {k}: final state"
    ist.close()
    sst.close()


def test_full_batch_tiling_property(ctx, oracle):
    """BASELINE configs[2] size (65536 stereo frames = 131072 channel units): 64 distinct units tiled 2048x through the
    device path; tile 0 bit-exact against the oracle, every other tile equal to tile 0; two consecutive frames."""
    import torch
    import libxaac_b200 as xb
    g = np.load(GOLD)
    base_n, tiles = 64, 2048
    n = base_n * tiles
    rng = np.random.default_rng(9)
    u = np.arange(base_n)
    st = np.ascontiguousarray(g["st_in"][2 + (u & 1)])
    state = xb.SbrState(ctx, n, low_power=True)
    state.upload(np.tile(st, (tiles, 1)), None)
    imdct_state = xb.ImdctBatch(n)
    ovl = np.zeros((base_n, 512), np.int32)
    wstate = np.zeros((base_n, 2), np.uint8)
    for f in range(2):
        s = rng.integers(10, 22, (base_n, 1))
        spec = ((rng.random((base_n, 1024)) * 2 - 1) * 2.0 ** s).astype(np.int64).astype(np.int32)
        ics = np.zeros((base_n, 2), np.uint8)
        ics[:, 1] = rng.integers(0, 2, base_n)
        side = np.ascontiguousarray(g["side"][2 + (u & 1) + 2 * (((u >> 1) + f) % 12)])
        d_spec = torch.from_numpy(spec).cuda().repeat(tiles, 1)
        d_ics = torch.from_numpy(ics).cuda().repeat(tiles, 1)
        d_side = torch.from_numpy(side).cuda().repeat(tiles, 1)
        w32, adj = xb.imdct_process(ctx, imdct_state, d_spec, d_ics)
        p16 = xb.imdct_out_to_pcm16(ctx, w32, adj, 0)
        out, err = xb.sbr_dec_lp(ctx, state, d_side, p16, out_ch=2)
        torch.cuda.synchronize()
        assert int(err.abs().max().item()) == 0
        o = out.view(tiles, base_n // 2, 2048, 2)
        assert bool((o == o[0:1]).all().item()), f"frame {f}: tiles differ"
        out32, ovl, wstate, eadj = oracle.imdct_batch(spec, ovl, wstate, ics)
        e16 = oracle.imdct_out_to_pcm16(out32, eadj, 0)
        got = o[0].cpu().numpy()
        for k in range(base_n):
            st[k], eo, ee = oracle.sbr_dec_lp(side[k], st[k], e16[k])
            assert ee == 0 and np.array_equal(got[k // 2, :, k % 2], eo), f"frame {f} unit {k}: tile 0 vs oracle"
    state.close()


def test_non_lockstep_ring_positions(ctx, oracle):
    """The reference keeps ring position and coefficient phase of both banks in lock step; the kernel's time-invariant
    window forms rely on that and fall back to the literal ring emulation otherwise.  Drive every (position, phase)
    combination the reference functions accept and compare with the oracle."""
    g = np.load(GOLD)
    n = 400
    side, st, tin = oracle_util.synth_sbr_lp_units(n, 17, g)
    rng = np.random.default_rng(17)
    st[:, 320] = 32 * rng.integers(0, 10, n)
    st[:, 321] = 64 * rng.integers(0, 10, n)
    st[:, 322] = 128 * rng.integers(0, 10, n)
    st[:, 323] = 64 * rng.integers(0, 10, n)
    outs, st2 = run_gpu(ctx, side, st, tin)
    out, err = outs[0]
    for u in range(n):
        es, eo, ee = oracle.sbr_dec_lp(side[u], st[u], tin[u])
        check(u, st2[u], out[u], err[u], es, eo, ee, "non-lockstep")


def test_sbr_dec_lp_w32_matches_handover_kernel(ctx):
    """xaac_b200_sbr_dec_lp_w32_dev (the WORD32 -> WORD16 hand-over of ixheaacd_allocate_sbr_scr inside the stage's load) against
    the separate hand-over kernel followed by xaac_b200_sbr_dec_lp_dev: PCM, err and state identical — lock-step units (vector load
    path) and units in every other (position, phase) state (literal ring emulation), incl. the hand-over's saturation corners."""
    import torch
    import libxaac_b200 as xb
    g = np.load(GOLD)
    n = 600
    side, st, _ = oracle_util.synth_sbr_lp_units(n, 23, g)
    rng = np.random.default_rng(23)
    half = n // 2
    st[half:, 320] = 32 * rng.integers(0, 10, n - half)
    st[half:, 321] = 64 * rng.integers(0, 10, n - half)
    res = []
    for w32_path in (False, True):
        state = xb.SbrState(ctx, n, low_power=True)
        state.upload(st, None)
        frames = []
        r2 = np.random.default_rng(5)
        for f in range(3):
            s = r2.integers(8, 31, (n, 1))
            w32 = ((r2.random((n, 1024)) * 2 - 1) * 2.0 ** s).astype(np.int64).clip(-2 ** 31, 2 ** 31 - 1).astype(np.int32)
            w32[0, :4] = (2 ** 31 - 1, -2 ** 31, 0x7FFF8000, -1)
            adj = r2.integers(1, 3, n).astype(np.int8)
            d_side = torch.from_numpy(np.ascontiguousarray(side)).cuda()
            d_w32, d_adj = torch.from_numpy(w32).cuda(), torch.from_numpy(adj).cuda()
            if w32_path:
                out, err = xb.sbr_dec_lp_w32(ctx, state, d_side, d_w32, d_adj, out_ch=2)
            else:
                out, err = xb.sbr_dec_lp(ctx, state, d_side, xb.imdct_out_to_pcm16(ctx, d_w32, d_adj, 0), out_ch=2)
            frames.append((out.cpu().numpy(), err.cpu().numpy()))
        res.append((frames, state.download()[0]))
        state.close()
    for f in range(3):
        assert np.array_equal(res[0][0][f][1], res[1][0][f][1]), f"frame {f}: err"
        assert np.array_equal(res[0][0][f][0], res[1][0][f][0]), f"frame {f}: PCM"
    assert np.array_equal(res[0][1], res[1][1])
