"""GPU parity tests for the HQ envelope-adjuster kernel (C-ABI) against the golden tapped records and the CPU oracle."""
import os

import numpy as np
import pytest

from tests import oracle_util

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden", "envcalc_tapped.npz")


def run_gpu(ctx, prm, sf, st, matrix):
    import torch
    import libxaac_b200 as xb
    d_m = torch.from_numpy(matrix.copy()).cuda()
    d_sf = torch.from_numpy(sf.copy()).cuda()
    d_st = torch.from_numpy(st.copy()).cuda()
    err = xb.calc_sbrenvelope(ctx, torch.from_numpy(prm).cuda(), d_sf, d_st, d_m)
    torch.cuda.synchronize()
    return d_m.cpu().numpy(), d_sf.cpu().numpy(), d_st.cpu().numpy(), err.cpu().numpy()


def compare(got, exp, what):
    for g, e, name in zip(got, exp, ("matrix", "sf", "state", "err")):
        if not np.array_equal(g, e):
            bad = np.argwhere(g != e)
            raise AssertionError(f"{what} {name}: {len(bad)} mismatches, first {bad[0]}: gpu={g[tuple(bad[0])]} "
                                 f"expected={e[tuple(bad[0])]}")


def test_golden_tapped_records(ctx):
    g = np.load(GOLD)
    got = run_gpu(ctx, g["prm"], g["sf_in"], g["st_in"], g["m_in"])
    compare(got, (g["m_out"], g["sf_out"], g["st_out"], g["err"].ravel()), "golden")


@pytest.mark.parametrize("seed,n", [(1, 1), (2, 70), (3, 3000)])
def test_random_units(ctx, oracle, seed, n):
    g = np.load(GOLD)
    prm, sf, st, matrix = oracle_util.synth_env_units(n, seed, g)
    got = run_gpu(ctx, prm, sf, st, matrix)
    exp = oracle.envcalc_batch(prm, sf, st, matrix)
    compare(got, exp, f"seed {seed}")


def test_error_units(ctx, oracle):
    """envelope borders beyond the matrix: the reference returns IA_FATAL_ERROR after partially updating the state"""
    g = np.load(GOLD)
    prm, sf, st, matrix = oracle_util.synth_env_units(64, 9, g)
    prm[::2, 16 + 1] = 20  # border_vec[1] = 20 -> end_pos 40 > 38
    got = run_gpu(ctx, prm, sf, st, matrix)
    exp = oracle.envcalc_batch(prm, sf, st, matrix)
    assert (exp[3] != 0).any() and (exp[3] == 0).any()
    compare(got, exp, "error units")
