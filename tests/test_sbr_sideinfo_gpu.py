"""GPU parity tests for the SBR side-info dequantisation (xaac_b200_dec_sbrdata_dev = ixheaacd_dec_sbrdata on the fixed-point
path) against the compiled reference function on the same seeded records (oracle/ref_shim_sd.c), and against
the committed golden records where the compiled reference is absent.  Bit-exact on the whole record."""
import os

import numpy as np
import pytest

from tests import oracle_util

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "sbr_sideinfo.npz")


def run_gpu(ctx, rec):
    import torch
    import libxaac_b200 as xb
    d = torch.from_numpy(rec.copy()).cuda()
    xb.dec_sbrdata(ctx, d)
    torch.cuda.synchronize()
    return d.cpu().numpy()


def explain(got, exp, rec):
    bad = np.argwhere(got != exp)
    u, w = bad[0]
    names = {v: k for k, v in oracle_util.SDC.items()}
    cw = (w - oracle_util.SD_CH) % oracle_util.SD_CH_WORDS if w >= oracle_util.SD_CH else -1
    field = max((o for o in names if o <= cw), default=None) if cw >= 0 else None
    return (f"{len(np.unique(bad[:, 0]))} records differ; first: record {u} word {w} (channel word {cw}, field "
            f"{names.get(field)}+{cw - field if field is not None else 0}) gpu={got[u, w]} ref={exp[u, w]} header={rec[u, :3]}")


@pytest.mark.parametrize("seed,n", [(11, 64), (12, 3000), (13, 20000)])
def test_seeded_records_vs_compiled_reference(ctx, ref, seed, n):
    rec = oracle_util.synth_sbrdata_records(n, seed)
    exp = ref.dec_sbrdata_batch(rec)
    got = run_gpu(ctx, rec)
    assert np.array_equal(got, exp), explain(got, exp, rec)
    assert (exp[:, 2] == 0).sum() > n // 2 and (exp[:, oracle_util.SD_CH + oracle_util.SDC["ERR_FLAG"]] != 0).sum() > 0


def test_golden_records(ctx):
    """records generated and run through the compiled reference in the build container (tools/make_golden_sideinfo.py)"""
    g = np.load(GOLDEN)
    got = run_gpu(ctx, g["records_in"])
    assert np.array_equal(got, g["records_out"]), explain(got, g["records_out"], g["records_in"])


def test_frames_carry_state(ctx, ref):
    """8 frames of 500 elements: sfb_nrg_prev / prev_noise_level / error flags are carried from each frame's output into the next
    frame's (freshly seeded) side info"""
    n = 500
    rec = oracle_util.synth_sbrdata_records(n, 21)
    cur_g, cur_r = rec.copy(), rec.copy()
    carry = []
    for c in range(2):
        o = oracle_util.SD_CH + c * oracle_util.SD_CH_WORDS
        for name, cnt in (("PREV_NRG", 56), ("PREV_NOISE", 5), ("ERR_FLAG", 2)):
            carry.append(slice(o + oracle_util.SDC[name], o + oracle_util.SDC[name] + cnt))
    for f in range(8):
        out_g, out_r = run_gpu(ctx, cur_g), ref.dec_sbrdata_batch(cur_r)
        assert np.array_equal(out_g, out_r), f"frame {f}: " + explain(out_g, out_r, cur_r)
        nxt = oracle_util.synth_sbrdata_records(n, 22 + f)
        for s in carry:
            nxt[:, s] = out_r[:, s]
        cur_g, cur_r = nxt.copy(), nxt.copy()


# ---- PS side info: ixheaacd_decode_ps_data ----------------------------------------------------------------------------
def run_gpu_ps(ctx, rec):
    import torch
    import libxaac_b200 as xb
    d = torch.from_numpy(rec.copy()).cuda()
    xb.decode_ps_data(ctx, d)
    torch.cuda.synchronize()
    return d.cpu().numpy()


def explain_ps(got, exp):
    bad = np.argwhere(got != exp)
    u, w = bad[0]
    names = {v: k for k, v in oracle_util.PSD.items()}
    field = max(o for o in names if o <= w)
    return (f"{len(np.unique(bad[:, 0]))} records differ; first: record {u} word {w} ({names[field]}+{w - field}) gpu={got[u, w]} "
            f"ref={exp[u, w]}")


@pytest.mark.parametrize("seed,n", [(31, 100), (32, 20000)])
def test_ps_records_vs_compiled_reference(ctx, ref, seed, n):
    rec = oracle_util.synth_psdata_records(n, seed)
    exp = ref.decode_ps_data_batch(rec)
    got = run_gpu_ps(ctx, rec)
    assert np.array_equal(got, exp), explain_ps(got, exp)


def test_ps_golden_records(ctx):
    g = np.load(GOLDEN)
    got = run_gpu_ps(ctx, g["ps_records_in"])
    assert np.array_equal(got, g["ps_records_out"]), explain_ps(got, g["ps_records_out"])


def test_ps_frames_carry_state(ctx, ref):
    """iid_par_prev / icc_par_prev carried over 6 frames"""
    n = 400
    cur = oracle_util.synth_psdata_records(n, 41)
    P = oracle_util.PSD
    for f in range(6):
        out_g, out_r = run_gpu_ps(ctx, cur), ref.decode_ps_data_batch(cur)
        assert np.array_equal(out_g, out_r), f"frame {f}: " + explain_ps(out_g, out_r)
        nxt = oracle_util.synth_psdata_records(n, 42 + f)
        nxt[:, P["IID_PREV"]: P["IID_PREV"] + 68] = out_r[:, P["IID_PREV"]: P["IID_PREV"] + 68]
        cur = nxt
