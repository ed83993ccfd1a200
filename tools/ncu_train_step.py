"""One NS2d training step (batch 1024, t_out 2, fp16s) for kernel-level profiling:  ncu ... python tools/ncu_train_step.py"""
import os, sys, torch, torch.nn.functional as F
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import lns_oracle as O
from lns_b200 import ops
from lns_b200.configs import get_config
from lns_b200.latent_dynamics import LatentDynamics
cfg = get_config("ns2d"); torch.manual_seed(1234)
model = LatentDynamics(cfg); model.load_state_dict(O.randomize_zero_init(model.state_dict())); model = model.cuda()
for p in model.autoencoder.parameters(): p.requires_grad_(False)
z_in, z_out = O.train_inputs(cfg, 1024, 2, seed=0); z_in, z_out = z_in.cuda(), z_out.cuda()
for _ in range(2):
    model.zero_grad(set_to_none=True)
    with ops.precision("fp16s"):
        model(z_in, z_out, F.smooth_l1_loss).backward()
torch.cuda.synchronize()
