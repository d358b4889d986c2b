"""2-rank probe of the peer-buffer path of the C-ABI: rank 0 exports a buffer (xaac_b200_ipc_export), rank 1 imports it on its own
device and runs an IMDCT launch whose INPUT lives in rank 0's HBM and whose output goes back there by a peer copy."""
import datetime
import os
import sys
import traceback

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import libxaac_b200 as xb  # noqa: E402

rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dev = torch.device("cuda", lr)
dist.init_process_group("nccl", device_id=dev, timeout=datetime.timedelta(seconds=60))
ctx = xb.Context(lr)
ok = 1
n = 256
try:
    objs = [None]
    if rank == 0:
        big = ctx.dev_tensor(n * 4096, torch.int32, (n, 1024))
        big.copy_(torch.randint(-1 << 20, 1 << 20, (n, 1024), dtype=torch.int32, device=dev))
        out = ctx.dev_tensor(n * 4096, torch.int32, (n, 1024))
        out.zero_()
        torch.cuda.synchronize(dev)
        objs = [[ctx.ipc_export(big), ctx.ipc_export(out)]]
    dist.broadcast_object_list(objs, src=0)
    if rank != 0:
        big = ctx.ipc_import(objs[0][0], n * 4096, torch.int32, (n, 1024))
        out = ctx.ipc_import(objs[0][1], n * 4096, torch.int32, (n, 1024))
        print("rank", rank, "mapped", big.device, tuple(big.shape), flush=True)
        st = xb.ImdctBatch(n, device=dev)
        ics = torch.zeros((n, 2), dtype=torch.uint8, device=dev)
        o, adj = xb.imdct_process(ctx, st, big, ics)
        torch.cuda.synchronize(dev)
        st2 = xb.ImdctBatch(n, device=dev)
        o2, _ = xb.imdct_process(ctx, st2, big.to(dev).contiguous(), ics)
        torch.cuda.synchronize(dev)
        print("kernel with its input in the peer's HBM: identical to a local run:", bool(torch.equal(o, o2)), flush=True)
        out.copy_(o)
        torch.cuda.synchronize(dev)
except Exception:
    ok = 0
    print("rank", rank, "FAILED:", traceback.format_exc(), file=sys.stderr, flush=True)
t = torch.tensor([ok], device=dev)
dist.all_reduce(t, op=dist.ReduceOp.MIN)
dist.barrier()
if rank == 0:
    torch.cuda.synchronize()
    print("all ok:", int(t.item()), "rank 0 sees the peer's result:", bool((out != 0).any()), flush=True)
dist.barrier()
del big, out
dist.destroy_process_group()
