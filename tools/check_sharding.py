"""torchrun --nproc-per-node N tools/check_sharding.py : the trajectory-sharded N-GPU rollout (local rollouts + NCCL
all-gather) must equal the single-GPU rollout of the full batch bit for bit."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

import lns_oracle as O  # noqa: E402
from lns_b200.configs import get_config  # noqa: E402
from lns_b200.dist import ShardedRollout, init_from_env  # noqa: E402
from lns_b200.latent_dynamics import LatentDynamics  # noqa: E402
from lns_b200.rollout import Rollout  # noqa: E402

rank, local, world = init_from_env("nccl")
dev = torch.device("cuda", local)
torch.cuda.set_device(dev)
ok = True
for name, B, K in (("ns2d", 16, 4), ("twophase_cond", 6, 3)):
    cfg = get_config(name)
    torch.manual_seed(1234)
    model = LatentDynamics(cfg).eval()
    model.load_state_dict(O.randomize_zero_init(model.state_dict()))
    model = model.to(dev)
    x, p = O.make_inputs(cfg, B, seed=5)
    x = x.to(dev)
    p = p.to(dev) if p is not None else None
    with torch.no_grad():
        sharded = ShardedRollout(model, global_batch=B, steps=K, precision="bf16", use_graph=True)
        got = sharded(x, p).clone()
        full = Rollout(model, batch=B, steps=K, precision="bf16", use_graph=True)(x, p).clone()
    same = bool(torch.equal(got, full))
    flags = torch.tensor([1 if same else 0], device=dev)
    if world > 1:
        dist.all_reduce(flags, op=dist.ReduceOp.MIN)
    if rank == 0:
        print(f"{name}: world={world} B={B} K={K} sharded+all-gather == single-GPU rollout: {bool(flags.item())}")
    ok = ok and bool(flags.item())
if world > 1:
    dist.barrier()
    dist.destroy_process_group()
sys.exit(0 if ok else 1)
