"""Trajectory-sharded multi-GPU rollout: one process per GPU (torch.distributed, NCCL over NVLink 5 / NVSwitch).

Trajectories never interact (every normalisation on the path is per sample, eval mode), so rank r
simply rolls out its own contiguous slice of the batch with the single-GPU engine -- no per-step communication.  The
only collective is the final all-gather of the predicted fields.  Host-side logic here is backend agnostic and is
covered by world_size-2 gloo tests on CPU (tests/test_dist_cpu.py)."""
import os

import torch
import torch.distributed as dist


def shard_bounds(batch, rank, world):
    """Contiguous, as-even-as-possible slice [lo, hi) of the trajectories owned by `rank` (ragged batches allowed)."""
    base, rem = divmod(batch, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def init_from_env(backend=None):
    """Initialise torch.distributed from torchrun's environment (RANK / WORLD_SIZE / LOCAL_RANK / MASTER_*)."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1 and not dist.is_initialized():
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29500")
        kw = {}
        if backend == "nccl":
            torch.cuda.set_device(local)
            kw["device_id"] = torch.device("cuda", local)
        dist.init_process_group(backend=backend, rank=rank, world_size=world, **kw)
    return rank, local, world


def gather_fields(local, batch, group=None):
    """All-gather per-rank results [b_r, ...] (b_r from shard_bounds) into the full [batch, ...] tensor on every rank.
    Equal shards use one all_gather_into_tensor; ragged shards pad to the largest shard."""
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return local
    world = dist.get_world_size(group)
    sizes = [shard_bounds(batch, r, world) for r in range(world)]
    counts = [hi - lo for lo, hi in sizes]
    tail = tuple(local.shape[1:])
    if len(set(counts)) == 1:
        out = torch.empty((batch,) + tail, dtype=local.dtype, device=local.device)
        dist.all_gather_into_tensor(out, local.contiguous(), group=group)
        return out
    mx = max(counts)
    padded = torch.zeros((mx,) + tail, dtype=local.dtype, device=local.device)
    padded[:local.shape[0]] = local
    buf = torch.empty((world * mx,) + tail, dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(buf, padded, group=group)
    return torch.cat([buf[r * mx:r * mx + counts[r]] for r in range(world)], 0)


class OverlappedGather:
    """All-gather of equal per-rank results that overlaps the NEXT step's compute: ``submit(local)`` copies the rank's result
    into a staging buffer (so the producer may overwrite ``local`` at once -- the rollout graph writes into a fixed buffer)
    and starts the all-gather asynchronously on the backend's own stream; ``wait()`` (also called by the next ``submit``)
    returns the gathered [world * b, ...] tensor.  Per-rank shards must be equal (use gather_fields for ragged batches)."""

    def __init__(self, local_shape, dtype, device, group=None):
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.stage = torch.empty(tuple(local_shape), dtype=dtype, device=device)
        self.full = torch.empty((local_shape[0] * self.world,) + tuple(local_shape[1:]), dtype=dtype, device=device)
        self.pending = None

    def submit(self, local):
        self.wait()  # the previous gather still reads the staging buffer
        self.stage.copy_(local, non_blocking=True)
        if self.world == 1:
            self.full.copy_(self.stage, non_blocking=True)
        else:
            self.pending = dist.all_gather_into_tensor(self.full, self.stage, group=self.group, async_op=True)

    def wait(self):
        if self.pending is not None:
            self.pending.wait()
            self.pending = None
        return self.full


class P2PGather:
    """The same contract as OverlappedGather, with the transfer OFF the SMs: every rank's [world * b, ...] result buffer is
    CUDA symmetric memory (torch.distributed._symmetric_memory: one allocation mapped into every peer of the node), and
    ``submit`` PUSHES the rank's shard into slot `rank` of every peer's buffer with plain device-to-device copies on a side
    stream -- the copy engines move the bytes over NVLink / NVSwitch while the next rollout owns all SMs and no NCCL kernel
    competes with the decode launches, which are sized to whole 148-SM waves (the NCCL all-gather cost 1.8 ms of a
    97 ms step at 8 GPUs).  Two tiny device-side barriers on the side stream order the exchange: one before the pushes (every
    peer has finished reading the previous result) and one after (every shard has landed everywhere)."""

    def __init__(self, local_shape, dtype, device, group=None):
        import torch.distributed._symmetric_memory as symm
        self.group = group if group is not None else dist.group.WORLD
        self.world = dist.get_world_size(self.group)
        self.rank = dist.get_rank(self.group)
        self.b = local_shape[0]
        full_shape = (self.b * self.world,) + tuple(local_shape[1:])
        self.stage = torch.empty(tuple(local_shape), dtype=dtype, device=device)
        self.full = symm.empty(full_shape, dtype=dtype, device=device)
        self.hdl = symm.rendezvous(self.full, self.group)
        # slot `rank` of every peer's result buffer, as local tensors
        self.slots = [self.hdl.get_buffer(r, full_shape, dtype)[self.rank * self.b:(self.rank + 1) * self.b] for r in range(self.world)]
        self.stream = torch.cuda.Stream(device)
        self.staged = torch.cuda.Event()
        self.done = torch.cuda.Event()
        self.pending = False

    def submit(self, local):
        cur = torch.cuda.current_stream(self.stage.device)
        self.wait()
        self.stage.copy_(local, non_blocking=True)
        self.staged.record(cur)
        with torch.cuda.stream(self.stream):
            self.stream.wait_event(self.staged)   # (also orders this rank's reads of the previous result before the barrier)
            self.hdl.barrier(channel=0)
            for r in range(self.world):
                self.slots[(self.rank + r) % self.world].copy_(self.stage, non_blocking=True)
            self.hdl.barrier(channel=1)
            self.done.record(self.stream)
        self.pending = True

    def wait(self):
        if self.pending:
            torch.cuda.current_stream(self.stage.device).wait_event(self.done)
            self.pending = False
        return self.full


def make_gather(local_shape, dtype, device, group=None):
    """The overlapped gather of the node: copy-engine pushes into symmetric memory (P2PGather) on CUDA + NCCL unless
    LNS_GATHER=nccl, the NCCL all-gather (OverlappedGather) otherwise (gloo / single rank / LNS_GATHER=nccl)."""
    want = os.environ.get("LNS_GATHER", "p2p")
    if (want == "p2p" and dist.is_initialized() and dist.get_world_size(group) > 1 and torch.device(device).type == "cuda"
            and dist.get_backend(group) == "nccl"):
        # symmetric memory needs peer access between every pair of ranks of the group (one NVLink / NVSwitch node); where the
        # rendezvous is refused (no P2P, a group spanning nodes) EVERY rank takes the NCCL transport -- the vote keeps them in step
        try:
            og = P2PGather(local_shape, dtype, device, group)
            ok = 1
        except Exception as ex:  # noqa: BLE001
            og, ok = None, 0
            err = repr(ex)[:200]
        vote = torch.tensor([ok], device=device)
        dist.all_reduce(vote, op=dist.ReduceOp.MIN, group=group)
        if int(vote.item()) == 1:
            return og
        if dist.get_rank(group) == 0:
            print(f"lns_b200.dist: symmetric-memory gather unavailable ({err if not ok else 'on another rank'}); using NCCL", flush=True)
    return OverlappedGather(local_shape, dtype, device, group)


class ShardedRollout:
    """Rolls out this rank's slice of a global batch and (optionally) all-gathers the fields."""

    def __init__(self, model, global_batch, steps, to_x=True, precision=None, use_graph=True, group=None, **kw):
        from .rollout import Rollout
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.global_batch = global_batch
        self.lo, self.hi = shard_bounds(global_batch, self.rank, self.world)
        self.local = Rollout(model, batch=self.hi - self.lo, steps=steps, to_x=to_x, precision=precision,
                             use_graph=use_graph, **kw)

    def __call__(self, x_global_or_local, param=None, gather=True):
        x = x_global_or_local
        if x.shape[0] == self.global_batch and self.world > 1:
            x = x[self.lo:self.hi]
            if param is not None:
                param = param[self.lo:self.hi]
        out = self.local(x, param)
        return gather_fields(out, self.global_batch, self.group) if gather else out
