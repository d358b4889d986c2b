"""GPU parity: esbr_hfgen_kernel (xaac_b200_esbr_generate_hf_dev) against the oracle / the compiled ixheaacd_generate_hf on
the same seeded units — float results compared bit for bit (no tolerance; the pre-processing test states its own), incl. cells the
stage must leave untouched, the patch table, the chirp-factor state and the error returns."""
import numpy as np
import pytest
import torch

from tests import oracle_util

pytestmark = pytest.mark.gpu


def _run(ctx, d, with_pv=True):
    import libxaac_b200 as xb
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
    dr, di, bw = t(d["dst_re"]), t(d["dst_im"]), t(d["bw_prev"])
    patch, err = xb.esbr_generate_hf(ctx, t(d["src_re"]), t(d["src_im"]), dr, di, t(d["par"]), bw,
                                     pv_re=t(d["pv_re"]) if with_pv else None, pv_im=t(d["pv_im"]) if with_pv else None,
                                     patch_out=t(d["patch_in"]) if "patch_in" in d else None)
    torch.cuda.synchronize()
    return dr.cpu().numpy(), di.cpu().numpy(), bw.cpu().numpy(), patch.cpu().numpy(), err.cpu().numpy()


def _same(a, b, what):
    dr1, di1, bw1, pt1, e1 = a
    dr2, di2, bw2, pt2, e2 = b
    assert np.array_equal(e1, e2), f"{what}: err differs at {np.argwhere(e1 != e2).ravel()[:8]}"
    ok = np.flatnonzero(e2 == 0)
    assert np.array_equal(pt1[ok], pt2[ok]), f"{what}: patch tables"
    assert np.array_equal(bw1[ok].view(np.int32), bw2[ok].view(np.int32)), f"{what}: bw_array_prev"
    for x, y, nm in ((dr1, dr2, "re"), (di1, di2, "im")):
        bad = np.argwhere(x[ok].view(np.int32) != y[ok].view(np.int32))
        assert len(bad) == 0, f"{what}: {nm} differs in {len(bad)} cells, first (unit,row,band) {ok[bad[0][0]]},{bad[0][1:]}"
    return len(ok)


@pytest.mark.parametrize("with_pv", [True, False])
def test_generate_hf_vs_oracle(ctx, oracle, with_pv):
    d = oracle_util.synth_esbr_hfgen_units(1500, 21 + with_pv, hbe=with_pv)
    good = _same(_run(ctx, d, with_pv), oracle_util.oracle_esbr_hfgen_batch(oracle, d, with_pv), f"pv={with_pv}")
    assert good > 1000


def test_generate_hf_vs_reference(ctx, ref):
    d = oracle_util.synth_esbr_hfgen_units(600, 5)
    _same(_run(ctx, d), oracle_util.ref_esbr_hfgen_batch(ref, d), "compiled reference")


def test_generate_hf_stream_state(ctx, oracle):
    """six frames with bw_array_prev carried on the device"""
    import libxaac_b200 as xb
    n = 64
    d = oracle_util.synth_esbr_hfgen_units(n, 40)
    d["par"][15::16, oracle_util.EHF["INVF_TBL"]:oracle_util.EHF["INVF_TBL"] + 5] = 64   # no failing units in the stream
    bw_g = torch.from_numpy(d["bw_prev"]).cuda()
    bw_o = d["bw_prev"].copy()
    for f in range(6):
        e = oracle_util.synth_esbr_hfgen_units(n, 50 + f)
        for k in ("src_re", "src_im", "pv_re", "pv_im", "dst_re", "dst_im"):
            d[k] = e[k]
        d["par"][:, oracle_util.EHF["INVF_PREV"]:oracle_util.EHF["INVF_PREV"] + 5] = d["par"][:, oracle_util.EHF["INVF"]:oracle_util.EHF["INVF"] + 5]
        d["par"][:, oracle_util.EHF["INVF"]:oracle_util.EHF["INVF"] + 5] = e["par"][:, oracle_util.EHF["INVF"]:oracle_util.EHF["INVF"] + 5]
        d["bw_prev"] = bw_o
        o = oracle_util.oracle_esbr_hfgen_batch(oracle, d)
        t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
        dr, di = t(d["dst_re"]), t(d["dst_im"])
        patch, err = xb.esbr_generate_hf(ctx, t(d["src_re"]), t(d["src_im"]), dr, di, t(d["par"]), bw_g, pv_re=t(d["pv_re"]),
                                         pv_im=t(d["pv_im"]))
        torch.cuda.synchronize()
        _same((dr.cpu().numpy(), di.cpu().numpy(), bw_g.cpu().numpy(), patch.cpu().numpy(), err.cpu().numpy()), o, f"frame {f}")
        bw_o = o[2]


def test_generate_hf_refuses_unsupported(ctx):
    import libxaac_b200 as xb
    d = oracle_util.synth_esbr_hfgen_units(8, 3)
    d["par"][1, oracle_util.EHF["USF4"]] = 1
    d["par"][2, oracle_util.EHF["FS"]] = 0
    out = _run(ctx, d)
    assert list(out[4][1:3]) == [-2, -2]
    assert np.array_equal(out[0][1:3].view(np.int32), d["dst_re"][1:3].view(np.int32))


@pytest.mark.parametrize("with_pv", [True, False])
def test_generate_hf_pre_processing(ctx, oracle, with_pv):
    """pre_proc_flag = 1 (ixheaacd_pre_processing, decoder/ixheaacd_sbrdec_lpfuncs.c:928-979).  The oracle is pinned bit for bit on
    the compiled function (tests/test_oracle_esbr.py).  The two libm calls (log10, pow: in double, rounded to float) are CUDA's on
    the device: a gain may differ from glibc's by one float ulp when the double result sits on a float rounding boundary, so the
    tolerance written here is 2 float ulps per output cell, on at most 0.1 % of the units; everything else bit for bit."""
    d = oracle_util.synth_esbr_hfgen_units(1500, 61 + with_pv, hbe=with_pv)
    d["par"][:, oracle_util.EHF["PRE_PROC"]] = 1
    if with_pv:
        d["par"][::2, oracle_util.EHF["PATCHING_MODE"]] = 1
    got = _run(ctx, d, with_pv)
    want = oracle_util.oracle_esbr_hfgen_batch(oracle, d, with_pv)
    assert np.array_equal(got[4], want[4])
    ok = np.flatnonzero(want[4] == 0)
    assert len(ok) > 1000
    assert np.array_equal(got[3][ok], want[3][ok]) and np.array_equal(got[2][ok].view(np.int32), want[2][ok].view(np.int32))
    off_units = 0
    for x, y in ((got[0], want[0]), (got[1], want[1])):
        xi, yi = x[ok].view(np.int32).astype(np.int64), y[ok].view(np.int32).astype(np.int64)
        diff = np.abs(xi - yi)
        assert diff.max() <= 2, f"max ulp distance {diff.max()}"
        off_units = max(off_units, int((diff.reshape(len(ok), -1).max(axis=1) > 0).sum()))
    assert off_units <= max(1, len(ok) // 1000), f"{off_units} of {len(ok)} units are not bit-identical"
    d0 = dict(d)
    d0["par"] = d["par"].copy()
    d0["par"][:, oracle_util.EHF["PRE_PROC"]] = 0
    assert not np.array_equal(_run(ctx, d0, with_pv)[0].view(np.int32), got[0].view(np.int32)), "the gains changed nothing"


def test_generate_hf_golden(ctx):
    """records tapped from real USAC decodes of the unmodified reference (default patching + harmonic transposer)"""
    from tests.test_oracle_esbr import check_hfgen_golden, golden_hfgen_units, load_esbr_golden
    g = load_esbr_golden("esbr_hfgen_tapped.npz")
    check_hfgen_golden(_run(ctx, golden_hfgen_units(g)), g, "kernel vs tapped decode")
