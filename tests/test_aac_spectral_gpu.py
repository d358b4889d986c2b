"""GPU parity tests for the AAC pre-IMDCT spectral stage (xaac_b200_aac_spectral_dev = batched
ixheaacd_channel_pair_process for AAC-LC: M/S stereo, intensity stereo, TNS) against the COMPILED reference function driven
through oracle/ref_shim_sps.c on the same records.  Integer work: every spectral line must be identical."""
import numpy as np
import pytest

from tests import oracle_util as ou

pytestmark = pytest.mark.gpu


def run_gpu(ctx, spec, rec):
    import torch
    import libxaac_b200 as xb
    s = torch.from_numpy(np.ascontiguousarray(spec, np.int32)).cuda()
    err = xb.aac_channel_pair_process(ctx, s, torch.from_numpy(np.ascontiguousarray(rec, np.uint8)).cuda())
    torch.cuda.synchronize()
    return s.cpu().numpy(), err.cpu().numpy()


def check(got, want, rec, what):
    out, err = got
    wout, werr = want
    assert (werr == 0).all()
    ok = err == 0
    assert ok.mean() > 0.95, (what, np.unique(err, return_counts=True))
    bad = np.flatnonzero((out[ok] != wout[ok]).any((1, 2)))
    if bad.size:
        u = np.flatnonzero(ok)[bad[0]]
        w = np.argwhere(out[u] != wout[u])
        hdr = rec[u][:32].view(np.int32)[:2].tolist()
        ch = [rec[u][ou.SPS_CH + c * ou.SPS_CH_BYTES:][:32].view(np.int32)[:5].tolist() for c in range(2)]
        raise AssertionError(f"{what}: {bad.size} elements differ; element {u} hdr {hdr} channels {ch}: first cells {w[:6].tolist()} "
                             f"got {out[u][tuple(w[0])]} want {wout[u][tuple(w[0])]}")
    return int(ok.sum())


def test_stereo_tools_only(ctx, ref):
    spec, rec = ou.synth_sps_units(1500, 3, tns=False)
    check(run_gpu(ctx, spec, rec), ou.ref_channel_pair_process(ref, spec, rec), rec, "M/S + intensity")


def test_tns_only(ctx, ref):
    spec, rec = ou.synth_sps_units(1500, 4, stereo_tools=False)
    want = ou.ref_channel_pair_process(ref, spec, rec)
    check(run_gpu(ctx, spec, rec), want, rec, "TNS")
    assert (want[0] != spec).any((1, 2)).sum() > 500


def test_whole_stage(ctx, ref):
    spec, rec = ou.synth_sps_units(4000, 5)
    check(run_gpu(ctx, spec, rec), ou.ref_channel_pair_process(ref, spec, rec), rec, "M/S + intensity + TNS")


def test_pns_with_generator_state(ctx, ref):
    """perceptual noise substitution: the generator state (current_seed) enters and leaves per element, correlated bands reuse the
    first channel's seeds, noise bands common to both channels drop out of the M/S mask; three consecutive frames carry the seed"""
    import torch
    import libxaac_b200 as xb
    n = 2500
    seed = np.random.default_rng(1).integers(-2**31, 2**31, n).astype(np.int32)
    d_seed = torch.from_numpy(seed.copy()).cuda()
    for f in range(3):
        spec, rec = ou.synth_sps_units(n, 20 + f, pns=True)
        wout, werr, wseed = ou.ref_channel_pair_process(ref, spec, rec, seed)
        s = torch.from_numpy(spec.copy()).cuda()
        err = xb.aac_channel_pair_process(ctx, s, torch.from_numpy(rec).cuda(), pns_seed=d_seed)
        torch.cuda.synchronize()
        check((s.cpu().numpy(), err.cpu().numpy()), (wout, werr), rec, f"PNS frame {f}")
        assert np.array_equal(d_seed.cpu().numpy(), wseed), f"frame {f}: generator state"
        assert (wseed != seed).mean() > 0.4
        seed = wseed


def test_pns_elements_are_refused_without_generator_state(ctx):
    spec, rec = ou.synth_sps_units(64, 6, pns=True)
    out, err = run_gpu(ctx, spec, rec)
    pns = np.array([any(rec[u][ou.SPS_CH + c * ou.SPS_CH_BYTES:][:32].view(np.int32)[3] for c in range(int(rec[u][:4].view(np.int32)[0])))
                    for u in range(64)])
    assert pns.sum() > 10 and (err[pns] == -2).all() and (err[~pns] == 0).all()
    assert np.array_equal(out[pns], spec[pns])


def test_golden_records(ctx):
    """records of the compiled reference function committed under tests/golden (no oracle/_ref needed)"""
    import os
    import torch
    import libxaac_b200 as xb
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "aac_spectral_ref.npz"))
    s = torch.from_numpy(g["spec_in"].copy()).cuda()
    seed = torch.from_numpy(g["seed_in"].copy()).cuda()
    err = xb.aac_channel_pair_process(ctx, s, torch.from_numpy(g["rec"]).cuda(), pns_seed=seed)
    torch.cuda.synchronize()
    assert int(err.abs().max()) == 0
    assert np.array_equal(s.cpu().numpy(), g["spec_out"]) and np.array_equal(seed.cpu().numpy(), g["seed_out"])
