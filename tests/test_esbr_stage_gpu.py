"""GPU parity: the whole float eSBR stage (xaac_b200_esbr_dec_dev — four launches with the history shifts, the regrouping and
the hand-overs fused into the banks) against records tapped around ixheaacd_sbr_dec in a real USAC decode, and against the
composed oracle on a larger batch.  Everything compared bit for bit."""
import numpy as np
import pytest
import torch

from tests import oracle_util
from tests.test_oracle_esbr import esbr_stage_golden_frames, load_esbr_golden

pytestmark = pytest.mark.gpu


def _state_to_gpu(xb, st, n):
    s = xb.EsbrDecBatch(n)
    for k in oracle_util.ESD_KEYS:
        getattr(s, k).copy_(torch.from_numpy(np.ascontiguousarray(st[k])))
    return s


def test_stage_golden_stream(ctx):
    import libxaac_b200 as xb
    g = load_esbr_golden("esbr_stage_tapped.npz")
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
    s = _state_to_gpu(xb, {k: g["in0_" + k] for k in oracle_util.ESD_KEYS}, 2)
    for f, r, rg in esbr_stage_golden_frames(g):
        ipar = t(g["ec_ipar_in"][r])
        pcm = torch.zeros((1, 2048, 2), dtype=torch.int16, device="cuda")
        out, err = xb.esbr_dec(ctx, s, t(g["time_in"][r]), t(g["hf_par"][r]), ipar, t(g["ec_fpar"][r]), t(rg), pcm16=pcm, ch_fac=2)
        torch.cuda.synchronize()
        assert int(err.abs().max()) == 0, f"frame {f}: {err.cpu().numpy()}"
        want = g["time_out"][r]
        assert np.array_equal(out.cpu().numpy().view(np.int32), want.view(np.int32)), f"frame {f}: time output"
        assert np.array_equal(pcm.cpu().numpy()[0], np.trunc(np.clip(want, -32768, 32767)).astype(np.int16).T), f"frame {f}: PCM16"
        assert np.array_equal(ipar.cpu().numpy(), g["ec_ipar_out"][r]), f"frame {f}: in/out parameter words"
        for k in ("anal_states", "anal_pos", "synth_states", "synth_pos", "bw_prev", "patch", "ec_state"):
            assert np.array_equal(getattr(s, k).cpu().numpy().view(np.int32), g["out_" + k][r].view(np.int32)), f"frame {f}: {k}"
    for k in ("qmf_re", "qmf_im", "out_re", "out_im"):
        assert np.array_equal(getattr(s, k).cpu().numpy().view(np.int32), g["out_" + k].view(np.int32)), k


def test_stage_batch_vs_oracle(ctx, oracle):
    """512 channels: the 16 tapped parameter sets tiled over units with seeded core input (int32 USAC-core hand-over) and
    perturbed history, three frames"""
    import libxaac_b200 as xb
    g = load_esbr_golden("esbr_stage_tapped.npz")
    rp = oracle_util.esbr_random_phase()
    rng = np.random.default_rng(5)
    n = 512
    idx = np.arange(n) % 16
    st = {k: np.ascontiguousarray(g["in0_" + k][idx % 2]).copy() for k in oracle_util.ESD_KEYS}
    for k in ("qmf_re", "qmf_im", "out_re", "out_im", "ec_state"):
        st[k] = (st[k] * rng.uniform(0.5, 2.0, (n,) + (1,) * (st[k].ndim - 1))).astype(np.float32)
    s = _state_to_gpu(xb, st, n)
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
    head = g["head"]
    ipar = g["ec_ipar_in"][idx].copy()
    for f in range(3):
        j = (idx + 2 * f) % 16
        core = (rng.standard_normal((n, 1024)) * 2.0 ** rng.uniform(8, 14, (n, 1))).astype(np.int32)
        xf = np.zeros((n, 1024), np.float32)
        oracle.lib.xo_esbr_core_to_float(oracle_util.P(core), oracle_util.P(xf), n * 1024)
        rg = np.stack([head[j, 7], head[j, 8], 2 * head[j, 9], 0 * head[j, 9]], 1).astype(np.int32)
        # parameters of record j, state words carried from the previous frame of this unit
        ip = g["ec_ipar_in"][j].copy()
        E = oracle_util.EEC
        for w in (E["SHORT_PREV"], E["HARM_INDEX"], E["PHASE_INDEX"], E["START_UP"]):
            ip[:, w] = ipar[:, w]
        ip[:, E["HARM_PREV"]:E["HARM_PREV"] + 16] = ipar[:, E["HARM_PREV"]:E["HARM_PREV"] + 16]
        out_o, st, ipar, err_o = oracle_util.oracle_esbr_stage(oracle, rp, st, xf, g["hf_par"][j], ip, g["ec_fpar"][j], rg)
        assert not err_o.any()
        ip_g = t(ip)
        out, err = xb.esbr_dec(ctx, s, t(core), t(g["hf_par"][j]), ip_g, t(g["ec_fpar"][j]), t(rg))
        torch.cuda.synchronize()
        assert int(err.abs().max()) == 0
        assert np.array_equal(out.cpu().numpy().view(np.int32), out_o.view(np.int32)), f"frame {f}: time output"
        assert np.array_equal(ip_g.cpu().numpy(), ipar), f"frame {f}: parameter words"
        for k in oracle_util.ESD_KEYS:
            assert np.array_equal(getattr(s, k).cpu().numpy().view(np.int32), st[k].view(np.int32)), f"frame {f}: {k}"


def test_stage_full_batch_tiling(ctx):
    """BASELINE batch size (131 072 channel units = 65 536 stereo frames): the two tapped channels tiled over the whole batch,
    two frames with the state carried — every copy must reproduce the tapped records bit for bit (persistent grid-stride
    tiling, state addressing and the interleaved PCM16 store at full size)"""
    import libxaac_b200 as xb
    g = load_esbr_golden("esbr_stage_tapped.npz")
    n = 131072
    rep = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda().repeat((n // 2,) + (1,) * (a.ndim - 1)).contiguous()
    s = xb.EsbrDecBatch(n)
    for k in oracle_util.ESD_KEYS:
        getattr(s, k).copy_(rep(g["in0_" + k]))
    pcm = torch.zeros((n // 2, 2048, 2), dtype=torch.int16, device="cuda")
    for f, r, rg in esbr_stage_golden_frames(g):
        if f >= 2:
            break
        ipar = rep(g["ec_ipar_in"][r])
        out, err = xb.esbr_dec(ctx, s, rep(g["time_in"][r]), rep(g["hf_par"][r]), ipar, rep(g["ec_fpar"][r]), rep(rg), pcm16=pcm, ch_fac=2)
        torch.cuda.synchronize()
        assert int(err.abs().max()) == 0
        want = torch.from_numpy(np.ascontiguousarray(g["time_out"][r])).cuda()
        assert torch.equal(out.view(torch.int32).view(n // 2, 2, 2048), want.view(torch.int32).unsqueeze(0).expand(n // 2, 2, 2048)), f"frame {f}"
        assert torch.equal(pcm, pcm[:1].expand_as(pcm))
        assert torch.equal(ipar, torch.from_numpy(np.ascontiguousarray(g["ec_ipar_out"][r])).cuda().repeat(n // 2, 1))
        for k in ("anal_states", "synth_states", "ec_state", "bw_prev"):
            w = torch.from_numpy(np.ascontiguousarray(g["out_" + k][r])).cuda()
            t = getattr(s, k)
            assert torch.equal(t.view(torch.int32).view(n // 2, 2, -1), w.view(torch.int32).unsqueeze(0).expand(n // 2, 2, w.shape[-1])), f"frame {f}: {k}"
