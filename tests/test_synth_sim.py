"""CPU check of the restructured QMF synthesis data flow (lane = slot modulation in registers, linear-time window, ring <->
row mapping): the kernel's per-lane source (libxaac_b200/csrc/qmf_synth_core.cuh, __host__ __device__) is run lane after
lane on the host by tests/sim/synth_sim.cu and compared with the oracle, bit for bit.  Test infrastructure: the simulator
is compiled into its own shared object under tests/sim/_build and is never part of libxaac_b200.so."""
import ctypes
import os
import shutil
import subprocess

import numpy as np
import pytest

from tests import oracle_util
from tests.oracle_util import P

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "tests", "sim", "synth_sim.cu")
CORE = os.path.join(ROOT, "libxaac_b200", "csrc", "qmf_synth_core.cuh")
OUT = os.path.join(ROOT, "tests", "sim", "_build", "libsynth_sim.so")


@pytest.fixture(scope="module")
def sim():
    if shutil.which("nvcc") is None:
        pytest.skip("nvcc not available")
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    if not os.path.exists(OUT) or os.path.getmtime(OUT) < max(os.path.getmtime(SRC), os.path.getmtime(CORE)):
        subprocess.check_call(["nvcc", "-O2", "-std=c++17", "-w", "-Wno-deprecated-gpu-targets", "-shared", "-Xcompiler", "-fPIC",
                               "-o", OUT, SRC])
    return ctypes.CDLL(OUT)


def run_sim(sim, qrom, matrix, fs, pos, params, fast_bits, ch_fac=1):
    n = matrix.shape[0]
    fs = fs.copy()
    pos = pos.copy()
    pcm = np.zeros((n, 2048 * ch_fac), np.int16)
    exact = 0
    for u in range(n):
        m = np.ascontiguousarray(matrix[u])
        f, p, pr = fs[u], pos[u], np.ascontiguousarray(params[u])
        exact += sim.synth_sim_unit(P(qrom), P(m), P(f), P(p), P(pr), P(pcm[u]), int(fast_bits), int(ch_fac))
    return pcm, fs, pos, exact


def check(a, b, what):
    for x, y, nm in zip(a, b, ("pcm", "filter_states", "pos")):
        if not np.array_equal(x, y):
            bad = np.argwhere(x != y)
            raise AssertionError(f"{what} {nm}: {len(bad)} mismatches, first at {bad[0]}: sim={x[tuple(bad[0])]} "
                                 f"oracle={y[tuple(bad[0])]}")


@pytest.mark.parametrize("fast_bits", [24, 0])
def test_lane_simulator_equals_oracle(sim, oracle, fast_bits):
    """fast_bits = 24 (below the bound the kernel derives, 25): small units take the wrapping path, the saturating test units
    the exact one; fast_bits = 0 forces every unit through the exact path"""
    matrix, fs, pos, params = oracle_util.synth_qmf_units(160, 4242)
    pcm, fs2, pos2, exact = run_sim(sim, oracle.qrom, matrix, fs, pos, params, fast_bits)
    check((pcm, fs2, pos2), oracle.synth_batch(matrix, fs, pos, params), f"fast_bits {fast_bits}")
    assert 0 < exact < 160 if fast_bits else exact >= 159


def test_every_ring_and_filter_phase(sim, oracle):
    matrix, fs, pos, params = oracle_util.synth_qmf_units(100, 9)
    pos[:, 0] = (np.arange(100) % 10) * 128
    pos[:, 1] = (np.arange(100) // 10) * 64
    pcm, fs2, pos2, _ = run_sim(sim, oracle.qrom, matrix, fs, pos, params, 24)
    check((pcm, fs2, pos2), oracle.synth_batch(matrix, fs, pos, params), "phases")


def test_scale_factor_sweep_and_state_carry(sim, oracle):
    vals = np.arange(-45, 26)
    n = len(vals)
    matrix, fs, pos, params = oracle_util.synth_qmf_units(n, 17)
    params[:, 0] = vals
    params[:, 1] = vals[::-1]
    params[:, 2] = np.roll(vals, 7)
    params[:, 4] = 20
    params[:, 5] = 48
    for f in range(3):  # state carried over three frames
        pcm, fs2, pos2, _ = run_sim(sim, oracle.qrom, matrix, fs, pos, params, 24)
        e = oracle.synth_batch(matrix, fs, pos, params)
        check((pcm, fs2, pos2), e, f"shifts frame {f}")
        fs, pos = e[1], e[2]
        matrix = np.roll(matrix, 1, axis=0)
