"""torchrun --nproc-per-node N tools/check_gather.py : P2PGather (copy-engine pushes into symmetric memory) against the NCCL
all-gather -- bit equality over several rounds with changing data, and the time of one gather of the bench's result shape."""
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from lns_b200.dist import OverlappedGather, P2PGather, init_from_env  # noqa: E402

rank, local, world = init_from_env("nccl")
dev = torch.device("cuda", local)
shape = (1184, 20, 1, 64, 64)
p2p = P2PGather(shape, torch.float32, dev)
ref = OverlappedGather(shape, torch.float32, dev)
ok = True
for it in range(4):
    g = torch.Generator(device=dev).manual_seed(100 * it + rank)
    x = torch.randn(shape, device=dev, generator=g)
    p2p.submit(x)
    ref.submit(x)
    x.zero_()  # the producer may overwrite its buffer at once
    a, b = p2p.wait(), ref.wait()
    torch.cuda.synchronize()
    ok = ok and torch.equal(a, b)
    dist.barrier()
res = {}
for name, og in (("p2p", p2p), ("nccl", ref)):
    x = torch.randn(shape, device=dev)
    for _ in range(2):
        og.submit(x); og.wait()
    torch.cuda.synchronize(); dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        og.submit(x); og.wait()
    e1.record(); torch.cuda.synchronize()
    res[name] = e0.elapsed_time(e1) / 5
t = torch.tensor([float(ok)], device=dev)
dist.all_reduce(t, op=dist.ReduceOp.MIN)
if rank == 0:
    print(f"world {world}: p2p == nccl bitwise: {bool(t.item())}; one gather of {shape} fp32 per rank: p2p {res['p2p']:.2f} ms, nccl {res['nccl']:.2f} ms")
dist.barrier()
dist.destroy_process_group()
sys.exit(0 if t.item() else 1)
