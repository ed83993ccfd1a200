"""20-step free-running drift of the NS2d rollout (bench.py's parity leg) under the current environment switches."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import torch
import bench
import lns_oracle as O
from lns_b200.configs import get_config
from lns_b200.latent_dynamics import LatentDynamics
cfg = get_config("ns2d"); torch.manual_seed(1234)
model = LatentDynamics(cfg).eval(); model.load_state_dict(O.randomize_zero_init(model.state_dict())); model = model.to("cuda:0")
d = bench.parity_check(model, cfg, sys.argv[1] if len(sys.argv) > 1 else "fp16s", torch.device("cuda:0"), batch=8, drift_steps=20)
print(json.dumps({k: v for k, v in d.items() if k != "vs"}))
