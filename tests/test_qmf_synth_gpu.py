"""GPU parity tests for the fixed-point HQ 64-band QMF synthesis kernel (through the C-ABI) against the CPU oracle
on the same seeded inputs, plus size-independent properties at the BASELINE.json batch size. Bit-exact."""
import numpy as np
import pytest

from tests import oracle_util

pytestmark = pytest.mark.gpu


def run_gpu(ctx, matrix, fs, pos, params, ch_fac=1):
    import torch
    import libxaac_b200 as xb
    n = matrix.shape[0]
    st = xb.QmfSynthBatch(n)
    st.filter_states.copy_(torch.from_numpy(fs))
    st.pos.copy_(torch.from_numpy(pos))
    d_m = torch.from_numpy(matrix).cuda()
    out = xb.cplx_synt_qmffilt(ctx, st, d_m, torch.from_numpy(params).cuda(), ch_fac=ch_fac)
    torch.cuda.synchronize()
    assert torch.equal(d_m.cpu(), torch.from_numpy(matrix)), "matrix must not be modified"
    return out.cpu().numpy(), st.filter_states.cpu().numpy(), st.pos.cpu().numpy()


def assert_same(g, o, what=""):
    for a, b, nm in zip(g, o, ("pcm", "filter_states", "pos")):
        if not np.array_equal(a, b):
            bad = np.argwhere(a != b)
            raise AssertionError(f"{what} {nm}: {len(bad)} mismatches, first at {bad[0]}: gpu={a[tuple(bad[0])]} "
                                 f"oracle={b[tuple(bad[0])]}")


@pytest.mark.parametrize("seed,n", [(1, 1), (2, 13), (3, 200), (4, 3000)])
def test_random_units(ctx, oracle, seed, n):
    matrix, fs, pos, params = oracle_util.synth_qmf_units(n, seed)
    assert_same(run_gpu(ctx, matrix, fs, pos, params), oracle.synth_batch(matrix, fs, pos, params), f"seed {seed}")


def test_every_ring_and_filter_phase(ctx, oracle):
    matrix, fs, pos, params = oracle_util.synth_qmf_units(100, 9)
    pos[:, 0] = (np.arange(100) % 10) * 128
    pos[:, 1] = (np.arange(100) // 10) * 64
    assert_same(run_gpu(ctx, matrix, fs, pos, params), oracle.synth_batch(matrix, fs, pos, params), "phases")


def test_scale_factor_sweep(ctx, oracle):
    """all block shifts from -31 to +31 in low and high band"""
    vals = np.arange(-45, 26)
    n = len(vals)
    matrix, fs, pos, params = oracle_util.synth_qmf_units(n, 17)
    params[:, 0] = vals
    params[:, 1] = vals[::-1]
    params[:, 2] = np.roll(vals, 7)
    params[:, 4] = 20
    params[:, 5] = 48
    assert_same(run_gpu(ctx, matrix, fs, pos, params), oracle.synth_batch(matrix, fs, pos, params), "shifts")


def test_interleaved_stereo_output(ctx, oracle):
    matrix, fs, pos, params = oracle_util.synth_qmf_units(32, 5)
    pcm, fs2, pos2 = run_gpu(ctx, matrix, fs, pos, params, ch_fac=2)
    e_pcm, e_fs, e_pos = oracle.synth_batch(matrix, fs, pos, params)
    assert pcm.shape == (16, 2048, 2)
    assert np.array_equal(pcm[:, :, 0], e_pcm[0::2]) and np.array_equal(pcm[:, :, 1], e_pcm[1::2])
    assert np.array_equal(fs2, e_fs) and np.array_equal(pos2, e_pos)


def test_stream_state_carry(ctx, oracle):
    """48 channels x 8 frames, state stays on the device between calls"""
    import torch
    import libxaac_b200 as xb
    n = 48
    st = xb.QmfSynthBatch(n)
    fs = np.zeros((n, 1280), np.int16)
    pos = np.zeros((n, 2), np.int16)
    for f in range(8):
        matrix, _, _, params = oracle_util.synth_qmf_units(n, 300 + f)
        out = xb.cplx_synt_qmffilt(ctx, st, torch.from_numpy(matrix).cuda(), torch.from_numpy(params).cuda())
        e_pcm, fs, pos = oracle.synth_batch(matrix, fs, pos, params)
        assert np.array_equal(out.cpu().numpy(), e_pcm), f"frame {f}"
    assert np.array_equal(st.filter_states.cpu().numpy(), fs) and np.array_equal(st.pos.cpu().numpy(), pos)


def test_host_entry_point(ctx, oracle):
    import torch
    import libxaac_b200 as xb
    n = 9000  # > 2 chunks of 4096
    matrix, fs, pos, params = oracle_util.synth_qmf_units(n, 55)
    st = xb.QmfSynthHostState(ctx, n)
    st.upload(torch.from_numpy(fs), torch.from_numpy(pos))
    h_pcm = torch.empty((n, 2048), dtype=torch.int16).pin_memory()
    xb.cplx_synt_qmffilt_host(ctx, st, torch.from_numpy(matrix).pin_memory(), torch.from_numpy(params), h_pcm)
    e = oracle.synth_batch(matrix, fs, pos, params)
    d_fs, d_pos = st.download()
    assert_same((h_pcm.numpy(), d_fs.numpy(), d_pos.numpy()), e, "host api")
    st.close()


def test_full_batch_properties(ctx, oracle):
    """BASELINE.json batch (131072 units): (1) tiled units give position-independent results that equal the oracle
    on the base set; (2) zero in + zero state -> zero out, zero state; (3) ring/coefficient offsets advance by
    32 slots (mod 10)."""
    import torch
    import libxaac_b200 as xb
    base_n, reps = 1024, 128
    n = base_n * reps
    matrix, fs, pos, params = oracle_util.synth_qmf_units(base_n, 777)
    st = xb.QmfSynthBatch(n)
    st.filter_states.copy_(torch.from_numpy(fs).cuda().repeat(reps, 1))
    st.pos.copy_(torch.from_numpy(pos).cuda().repeat(reps, 1))
    d_m = torch.from_numpy(matrix).cuda().repeat(reps, 1, 1)
    d_p = torch.from_numpy(params).cuda().repeat(reps, 1)
    out = xb.cplx_synt_qmffilt(ctx, st, d_m, d_p)
    torch.cuda.synchronize()
    e_pcm, e_fs, e_pos = oracle.synth_batch(matrix, fs, pos, params)
    o = out.view(reps, base_n, 2048)
    assert torch.equal(o[0].cpu(), torch.from_numpy(e_pcm))
    assert bool((o == o[0:1]).all())
    assert bool((st.filter_states.view(reps, base_n, 1280) == torch.from_numpy(e_fs).cuda()[None]).all())
    assert bool((st.pos.view(reps, base_n, 2) == torch.from_numpy(e_pos).cuda()[None]).all())
    exp_off = (pos[:, 0].astype(np.int64) - 32 * 128) % 1280
    exp_fp = (pos[:, 1].astype(np.int64) + 32 * 64) % 640
    assert np.array_equal(e_pos[:, 0], exp_off) and np.array_equal(e_pos[:, 1], exp_fp)
    st0 = xb.QmfSynthBatch(256)
    z = xb.cplx_synt_qmffilt(ctx, st0, torch.zeros((256, 32, 128), dtype=torch.int32, device="cuda"),
                             xb.synth_params(-8, -8, -8, 32, 64).expand(256, 8).contiguous().cuda())
    assert int(z.abs().max()) == 0 and int(st0.filter_states.abs().max()) == 0


def test_tma_staged_variant_is_bit_identical():
    """The opt-in bulk-copy-staged lane = slot kernel (XAAC_B200_SYNTH_TMA=1, read once per process) must pass this same
    module; its arithmetic is also checked on the CPU by tests/test_synth_sim.py."""
    import os
    import subprocess
    import sys
    if os.environ.get("XAAC_B200_SYNTH_TMA"):
        pytest.skip("already running the variant")
    env = dict(os.environ, XAAC_B200_SYNTH_TMA="1")
    r = subprocess.run([sys.executable, "-m", "pytest", os.path.abspath(__file__), "-x", "-q", "-m", "gpu", "-k",
                        "not tma_staged_variant"], env=env, capture_output=True, text=True, timeout=900,
                       cwd=os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
