"""CPU test of the multi-GPU host logic (world_size 2, gloo): stream partition, scatter of pre-parsed frames from
rank 0, per-rank processing with the CPU oracle standing in for the device stage, gather of PCM in stream order."""
import os
import socket
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_stream_range_covers_everything():
    from libxaac_b200.shard import stream_range
    for n in (0, 1, 7, 8, 131072, 131075):
        for world in (1, 2, 3, 8):
            blocks = [stream_range(n, r, world) for r in range(world)]
            assert blocks[0][0] == 0 and blocks[-1][1] == n
            assert all(blocks[i][1] == blocks[i + 1][0] for i in range(world - 1))
            sizes = [b - a for a, b in blocks]
            assert max(sizes) - min(sizes) <= 1


def _worker(rank, world, port, n_streams, ret):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from libxaac_b200.shard import gather_streams, scatter_streams, stream_range
    from tests import oracle_util
    orc = oracle_util.Oracle()
    spec, ovl, wstate, ics = oracle_util.synth_units(n_streams, 5)
    full = torch.from_numpy(spec) if rank == 0 else torch.empty((n_streams, 1024), dtype=torch.int32)
    mine = scatter_streams(full, n_streams, src=0)
    a, b = stream_range(n_streams, rank, world)
    assert mine.shape[0] == b - a and np.array_equal(mine.numpy(), spec[a:b])
    out, _, _, _ = orc.imdct_batch(mine.numpy(), ovl[a:b], wstate[a:b], ics[a:b])   # stand-in for the device stage
    got = gather_streams(torch.from_numpy(out), n_streams, dst=0)
    if rank == 0:
        exp, _, _, _ = orc.imdct_batch(spec, ovl, wstate, ics)
        ret.put(bool(np.array_equal(got.numpy(), exp)))
    dist.barrier()
    dist.destroy_process_group()


def test_scatter_process_gather_world2():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    ret = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, 37, ret)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    assert ret.get(timeout=5) is True
