"""Multi-GPU sharding of a batch of streams.

Streams are independent (frames of one stream are not: overlap / QMF / SBR / PS state carries from frame to frame), so
the batch is partitioned by stream: rank r of G owns the contiguous block [r*N/G, (r+1)*N/G) together with all its
state; ROM tables are replicated.  The DSP needs no collective.  Only when the pre-parsed buffers originate on one rank
(or the PCM must end up on one) is there an I/O exchange, done here with torch.distributed point-to-point / gather
calls (NCCL over NVLink on GPUs, gloo in the CPU tests)."""
import torch
import torch.distributed as dist


def stream_range(n_streams, rank, world):
    """contiguous block partition; the first n_streams % world ranks own one extra stream"""
    base, rem = divmod(int(n_streams), int(world))
    start = rank * base + min(rank, rem)
    return start, start + base + (1 if rank < rem else 0)


def scatter_streams(full, n_streams, src=0, group=None):
    """rank `src` holds `full` [n_streams, ...]; every rank receives its own block (frames in, per-frame side info).
    NCCL has no native scatter with ragged blocks: grouped send / recv."""
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    a, b = stream_range(n_streams, rank, world)
    if rank == src:
        reqs = []
        for r in range(world):
            ra, rb = stream_range(n_streams, r, world)
            if r != src and rb > ra:
                reqs.append(dist.isend(full[ra:rb].contiguous(), dst=r, group=group))
        mine = full[a:b].clone()
        for q in reqs:
            q.wait()
        return mine
    shape = list(full.shape)
    shape[0] = b - a
    mine = torch.empty(shape, dtype=full.dtype, device=full.device)
    if b > a:
        dist.recv(mine, src=src, group=group)
    return mine


def gather_streams(mine, n_streams, dst=0, group=None):
    """inverse of scatter_streams for the PCM: rank `dst` returns [n_streams, ...] in stream order, others None"""
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    if rank == dst:
        shape = [n_streams] + list(mine.shape[1:])
        out = torch.empty(shape, dtype=mine.dtype, device=mine.device)
        for r in range(world):
            ra, rb = stream_range(n_streams, r, world)
            if rb <= ra:
                continue
            if r == dst:
                out[ra:rb] = mine
            else:
                dist.recv(out[ra:rb], src=r, group=group)
        return out
    if mine.shape[0] > 0:
        dist.send(mine.contiguous(), dst=dst, group=group)
    return None
