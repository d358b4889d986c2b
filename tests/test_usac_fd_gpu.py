"""GPU parity tests for the USAC frequency-domain core transform (xaac_b200_usac_fd_frm_dec_dev = ixheaacd_fd_frm_dec for
pure FD streams) against the CPU oracle (itself pinned to the compiled reference by tests/test_oracle_usac.py), and —
where oracle/_ref is present — against the compiled reference directly."""
import numpy as np
import pytest

from tests import oracle_util

pytestmark = pytest.mark.gpu


def run_gpu(ctx, coef, ov, seq, shape, shape_prev):
    import torch
    import libxaac_b200 as xb
    n = coef.shape[0]
    st = xb.UsacFdBatch(n)
    st.overlap.copy_(torch.from_numpy(ov))
    st.wstate.copy_(torch.from_numpy(shape_prev.astype(np.uint8)))
    ics = torch.from_numpy(np.stack([seq, shape], 1).astype(np.uint8)).cuda()
    out = xb.usac_fd_frm_dec(ctx, st, torch.from_numpy(coef).cuda(), ics)
    torch.cuda.synchronize()
    return out.cpu().numpy(), st.overlap.cpu().numpy(), st.wstate.cpu().numpy()


@pytest.mark.parametrize("seed,n", [(1, 7), (2, 500), (3, 3000)])
def test_all_sequences_vs_oracle(ctx, oracle, seed, n):
    coef, ov = oracle_util.synth_usac_units(n, seed)
    rng = np.random.default_rng(seed)
    seq = rng.integers(0, 5, n).astype(np.int32)
    shape = rng.integers(0, 2, n).astype(np.int32)
    shape_prev = rng.integers(0, 2, n).astype(np.int32)
    out, ov2, ws = run_gpu(ctx, coef, ov, seq, shape, shape_prev)
    eo, ev, _ = oracle.usac_fd_batch(coef, ov, seq, shape, shape_prev)
    for u in range(n):
        if not np.array_equal(out[u], eo[u]):
            raise AssertionError(f"unit {u} (seq {seq[u]}): output differs at {np.argwhere(out[u] != eo[u]).ravel()[:8]}")
        if not np.array_equal(ov2[u], ev[u]):
            raise AssertionError(f"unit {u} (seq {seq[u]}): overlap differs at {np.argwhere(ov2[u] != ev[u]).ravel()[:8]}")
    assert np.array_equal(ws, shape.astype(np.uint8))


def test_vs_compiled_reference(ctx, ref):
    n = 300
    coef, ov = oracle_util.synth_usac_units(n, 11)
    rng = np.random.default_rng(11)
    seq = rng.integers(0, 5, n).astype(np.int32)
    shape = rng.integers(0, 2, n).astype(np.int32)
    shape_prev = rng.integers(0, 2, n).astype(np.int32)
    out, ov2, _ = run_gpu(ctx, coef, ov, seq, shape, shape_prev)
    eo, ev, _ = ref.usac_fd_batch(coef, ov, seq, shape, shape_prev)
    assert np.array_equal(out, eo) and np.array_equal(ov2, ev)


def test_streams_state_resident(ctx, oracle):
    """10 consecutive frames per stream with a legal window-sequence walk, overlap and previous shape staying on the device"""
    import torch
    import libxaac_b200 as xb
    n, frames = 256, 10
    seq, shape = oracle_util.usac_seq_walk(n, frames, 5)
    st = xb.UsacFdBatch(n)
    ov = np.zeros((n, 1024), np.int32)
    prev = np.zeros(n, np.int32)
    for f in range(frames):
        coef, _ = oracle_util.synth_usac_units(n, 200 + f)
        ics = torch.from_numpy(np.stack([seq[f], shape[f]], 1).astype(np.uint8)).cuda()
        out = xb.usac_fd_frm_dec(ctx, st, torch.from_numpy(coef).cuda(), ics)
        torch.cuda.synchronize()
        eo, ov, _ = oracle.usac_fd_batch(coef, ov, seq[f], shape[f], prev)
        prev = shape[f]
        assert np.array_equal(out.cpu().numpy(), eo), f"frame {f}"
    assert np.array_equal(st.overlap.cpu().numpy(), ov)


def test_golden_tapped_real_decode(ctx):
    """records tapped from the reference decoding a real xHE-AAC stream (tests/golden/usac_fd_tapped.npz)"""
    import os
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "usac_fd_tapped.npz"))
    h = g["hdr"]
    out, ov, ws = run_gpu(ctx, g["coef"], g["ov_in"], h[:, 1], h[:, 2], h[:, 3])
    assert np.array_equal(out, g["out"]) and np.array_equal(ov, g["ov_out"])
    assert np.array_equal(ws, h[:, 2].astype(np.uint8))


def test_full_batch_tiling_property(ctx):
    """BASELINE batch size (262 144 channel units = 131 072 stereo frames): the tapped records tiled over the whole batch —
    every copy must reproduce its record bit for bit (grid-stride tiling, all window sequences side by side)"""
    import os
    import torch
    import libxaac_b200 as xb
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "usac_fd_tapped.npz"))
    h = g["hdr"]
    m = len(h)
    n = 262144
    reps = (n + m - 1) // m
    tile = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda().repeat((reps,) + (1,) * (a.ndim - 1))[:n].contiguous()
    st = xb.UsacFdBatch(n)
    st.overlap.copy_(tile(g["ov_in"]))
    st.wstate.copy_(tile(h[:, 3].astype(np.uint8)))
    out = xb.usac_fd_frm_dec(ctx, st, tile(g["coef"]), tile(np.stack([h[:, 1], h[:, 2]], 1).astype(np.uint8)))
    torch.cuda.synchronize()
    assert torch.equal(out, tile(g["out"])) and torch.equal(st.overlap, tile(g["ov_out"]))
