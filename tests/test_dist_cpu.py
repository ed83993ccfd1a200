"""world_size-2 gloo tests (CPU) of the trajectory-sharding host logic in lns_b200/dist.py."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from lns_b200.dist import OverlappedGather, gather_fields, make_gather, shard_bounds


def test_shard_bounds_cover_the_batch():
    for batch in (1, 2, 7, 8, 64, 1000):
        for world in (1, 2, 3, 4, 8):
            spans = [shard_bounds(batch, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == batch
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, batch, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    full = torch.arange(batch * 2 * 3 * 4, dtype=torch.float32).view(batch, 2, 3, 4)  # [B, K, ...] "fields"
    lo, hi = shard_bounds(batch, rank, world)
    got = gather_fields(full[lo:hi].clone(), batch)
    q.put((rank, bool(torch.equal(got, full))))
    dist.destroy_process_group()


@pytest.mark.parametrize("batch", [8, 7])  # equal and ragged shards
def test_gather_fields_gloo_world2(batch):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, batch, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
    assert res == [(0, True), (1, True)]


def _worker_overlap(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    og = make_gather((3, 2, 4), torch.float32, torch.device("cpu"))  # gloo / CPU: the collective transport, never the P2P one
    ok = isinstance(og, OverlappedGather)
    buf = torch.empty(3, 2, 4)  # the producer's fixed output buffer, overwritten every step
    for step in range(4):
        buf.copy_(torch.full((3, 2, 4), float(10 * step + rank)))
        og.submit(buf)
        buf.fill_(-1.0)          # the next step overwrites the producer buffer while the gather is in flight
        want = torch.cat([torch.full((3, 2, 4), float(10 * step + r)) for r in range(world)], 0)
        ok = ok and bool(torch.equal(og.wait(), want))
    q.put((rank, ok))
    dist.destroy_process_group()


def test_overlapped_gather_gloo_world2():
    """the staging copy decouples the gather of step i from the producer buffer that step i+1 overwrites"""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker_overlap, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
    assert res == [(0, True), (1, True)]
