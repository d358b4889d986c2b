"""CPU tests: HQ envelope-adjuster oracle vs records tapped from real HE-AAC decodes (golden) and vs the compiled
reference on randomised units."""
import os

import numpy as np

from tests import oracle_util

GOLD = os.path.join(os.path.dirname(__file__), "golden", "envcalc_tapped.npz")


def test_oracle_matches_golden(oracle):
    g = np.load(GOLD)
    m, sf, st, err = oracle.envcalc_batch(g["prm"], g["sf_in"], g["st_in"], g["m_in"])
    assert len(g["prm"]) >= 30
    assert np.array_equal(m, g["m_out"])
    assert np.array_equal(sf, g["sf_out"])
    assert np.array_equal(st, g["st_out"])
    assert np.array_equal(err, g["err"].ravel())
    assert (g["m_out"] != g["m_in"]).any()


def test_env_rom_matches_reference(ref):
    assert np.array_equal(ref.rom_blob("ref_rom_env_tables", 2404), oracle_util.rom("env_rom.bin"))
    assert np.array_equal(ref.rom_blob("ref_rom_misc_tables", 2470), oracle_util.rom("misc_rom.bin"))


def test_oracle_matches_reference_random(oracle, ref):
    g = np.load(GOLD)
    n = 400
    prm, sf, st, matrix = oracle_util.synth_env_units(n, 41, g)
    m, s, t, err = oracle.envcalc_batch(prm, sf, st, matrix)
    for u in range(n):
        rm, rs, rt, rerr = ref.envcalc(prm[u], sf[u], st[u], matrix[u])
        assert np.array_equal(m[u], rm), f"unit {u} matrix"
        assert np.array_equal(s[u], rs) and np.array_equal(t[u], rt) and err[u] == rerr, f"unit {u} state"
