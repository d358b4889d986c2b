"""GPU parity tests for the HQ HF-generator kernel (C-ABI) against the golden tapped records and the CPU oracle."""
import os

import numpy as np
import pytest

from tests import oracle_util

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden", "hfgen_tapped.npz")


def run_gpu(ctx, lpc, matrix, prm, bw_prev):
    import torch
    import libxaac_b200 as xb
    d_m = torch.from_numpy(matrix.copy()).cuda()
    d_bw = torch.from_numpy(bw_prev.copy()).cuda()
    hb = xb.hf_generator(ctx, torch.from_numpy(lpc).cuda(), d_m, torch.from_numpy(prm).cuda(), d_bw)
    torch.cuda.synchronize()
    return d_m.cpu().numpy(), d_bw.cpu().numpy(), hb.cpu().numpy()


def test_golden_tapped_records(ctx):
    g = np.load(GOLD)
    m, bw, hb = run_gpu(ctx, g["lpc"], g["m_in"], g["prm"], g["bw_in"])
    assert np.array_equal(m, g["m_out"])
    assert np.array_equal(bw, g["bw_out"])
    assert np.array_equal(hb, g["hb"].astype(np.int16))


@pytest.mark.parametrize("seed,n", [(1, 1), (2, 50), (3, 2000)])
def test_random_units(ctx, oracle, seed, n):
    g = np.load(GOLD)
    lpc, matrix, prm, bw_prev = oracle_util.synth_hfgen_units(n, seed, g["prm"])
    gm, gbw, ghb = run_gpu(ctx, lpc, matrix, prm, bw_prev)
    em, ebw, ehb = oracle.hfgen_batch(lpc, matrix, prm, bw_prev)
    if not np.array_equal(gm, em):
        bad = np.argwhere(gm != em)
        raise AssertionError(f"matrix: {len(bad)} mismatches, first {bad[0]}: gpu={gm[tuple(bad[0])]} oracle={em[tuple(bad[0])]}")
    assert np.array_equal(gbw, ebw) and np.array_equal(ghb, ehb)
