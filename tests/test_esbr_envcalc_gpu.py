"""GPU parity: esbr_envcalc_kernel (xaac_b200_esbr_env_calc_dev) against the oracle / the compiled ixheaacd_sbr_env_calc on
the same seeded units — adjusted QMF cells, smoothing history, harmonic flags and indices compared bit for bit."""
import numpy as np
import pytest
import torch

from tests import oracle_util
from tests.test_oracle_esbr import same_envcalc

pytestmark = pytest.mark.gpu


def _run(ctx, d):
    import libxaac_b200 as xb
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
    re, im, ipar, state = t(d["re"]), t(d["im"]), t(d["ipar"]), t(d["state"])
    err = xb.esbr_env_calc(ctx, re, im, ipar, t(d["fpar"]), state)
    torch.cuda.synchronize()
    return re.cpu().numpy(), im.cpu().numpy(), ipar.cpu().numpy(), state.cpu().numpy(), err.cpu().numpy()


def test_env_calc_vs_oracle(ctx, oracle):
    rp = oracle_util.esbr_random_phase()
    d = oracle_util.synth_esbr_envcalc_units(2000, 31)
    good = same_envcalc(_run(ctx, d), oracle_util.oracle_esbr_envcalc_batch(oracle, d, rp), "oracle")
    assert good > 1800


def test_env_calc_vs_reference(ctx, ref):
    d = oracle_util.synth_esbr_envcalc_units(600, 8)
    same_envcalc(_run(ctx, d), oracle_util.ref_esbr_envcalc_batch(ref, d), "compiled reference")


def test_env_calc_stream_state(ctx, oracle):
    """five frames with the smoothing history, harmonic flags, phase / harmonic indices carried on the device"""
    import libxaac_b200 as xb
    rp = oracle_util.esbr_random_phase()
    n = 96
    d = oracle_util.synth_esbr_envcalc_units(n, 60)
    d["ipar"][:, oracle_util.EEC["NUM_NOISE_ENV"]] = np.where(d["ipar"][:, oracle_util.EEC["NUM_ENV"]] == 1, 1, 2)
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
    ipar_g, state_g = t(d["ipar"]), t(d["state"])
    carry = [oracle_util.EEC[k] for k in ("SHORT_PREV", "HARM_INDEX", "PHASE_INDEX", "START_UP")]
    hp = slice(oracle_util.EEC["HARM_PREV"], oracle_util.EEC["HARM_PREV"] + 16)
    for f in range(5):
        e = oracle_util.synth_esbr_envcalc_units(n, 70 + f)
        d["re"], d["im"], d["fpar"] = e["re"], e["im"], e["fpar"]
        o = oracle_util.oracle_esbr_envcalc_batch(oracle, d, rp)
        re, im = t(d["re"]), t(d["im"])
        err = xb.esbr_env_calc(ctx, re, im, ipar_g, t(d["fpar"]), state_g)
        torch.cuda.synchronize()
        same_envcalc((re.cpu().numpy(), im.cpu().numpy(), ipar_g.cpu().numpy(), state_g.cpu().numpy(), err.cpu().numpy()), o,
                     f"frame {f}")
        d["ipar"], d["state"] = o[2], o[3]


def test_env_calc_refuses_unsupported(ctx):
    d = oracle_util.synth_esbr_envcalc_units(8, 3)
    d["ipar"][0, oracle_util.EEC["RESET"]] = 1
    d["ipar"][1, oracle_util.EEC["SBR_MODE"]] = 2
    d["ipar"][2, oracle_util.EEC["INTER_TES"]] = 1
    d["ipar"][3, oracle_util.EEC["USF4"]] = 1
    out = _run(ctx, d)
    assert list(out[4][:4]) == [-2, -2, -2, -2]
    assert np.array_equal(out[0][:4].view(np.int32), d["re"][:4].view(np.int32))
    assert np.array_equal(out[3][:4].view(np.int32), d["state"][:4].view(np.int32))


def test_env_calc_golden(ctx):
    """records tapped from real USAC decodes of the unmodified reference, incl. 8 consecutive calls"""
    from tests.test_oracle_esbr import check_envcalc_golden, golden_envcalc_units, load_esbr_golden
    g = load_esbr_golden("esbr_envcalc_tapped.npz")
    check_envcalc_golden(_run(ctx, golden_envcalc_units(g)), g, "kernel vs tapped decode")


def test_env_calc_inter_tes_vs_reference(ctx, ref):
    """envelopes with inter-TES (gamma 1 / 2 / 4 mixed with 0): ixheaacd_apply_inter_tes between the gain pass and the sinusoids,
    against the compiled ixheaacd_sbr_env_calc fed with the same low band; all cells of both arrays compared (the low bands of
    the TES envelopes' slots are copied next to the high bands)"""
    import libxaac_b200 as xb
    d = oracle_util.synth_esbr_envcalc_tes_units(800, 17)
    want = oracle_util.ref_esbr_envcalc_tes_batch(ref, d)
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
    re, im, ipar, state = t(d["re"]), t(d["im"]), t(d["ipar"]), t(d["state"])
    err = xb.esbr_env_calc(ctx, re, im, ipar, t(d["fpar"]), state, low_re=t(d["low_re"]), low_im=t(d["low_im"]))
    torch.cuda.synchronize()
    got = (re.cpu().numpy(), im.cpu().numpy(), ipar.cpu().numpy(), state.cpu().numpy(), err.cpu().numpy())
    good = same_envcalc(got, want, "compiled reference, inter-TES")
    assert good > 600
    used = (d["ipar"][:, 44:52] != 0).any(1) & (want[4] == 0)
    assert used.sum() > 300
