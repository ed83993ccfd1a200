"""Drop-in check: execute an UNMODIFIED reference stage-2 script (from /root/reference or the staged
oracle/_ref) with THIS repository first on sys.path, so that its `from modules.autoencoder2d import SimpleAutoencoder`,
`from modules.basics import ResidualBlock, GroupNorm` and `from utils import dict2namespace` resolve to the drop-in package, build
the script's own ``LatentDynamics`` (its SimpleCNN / DilatedResidualBlock classes, its YAML), and roll it out through
``lns_b200.rollout.Rollout``.  The result must equal the repository's own LatentDynamics built under the same seed, bit for bit.

    python tools/run_unmodified_script.py [ns2d|sw|twophase|twophase_cond]"""
import importlib.machinery
import importlib.util
import os
import sys
import types

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = next((c for c in ("/root/reference", os.path.join(ROOT, "oracle", "_ref")) if os.path.isdir(os.path.join(c, "modules"))), None)
SCRIPTS = {"ns2d": ("configs/ns2d_stage2_prop.yml", "train_stage2_ns2d.py"),
           "sw": ("configs/SW_stage2_prop.yml", "train_stage2_SW.py"),
           "twophase": ("configs/twophase_stage2_prop.yml", "train_stage2_twophase.py"),
           "twophase_cond": ("configs/twophase_stage2_cond_prop.yml", "train_stage2_twophase_conditional.py")}


def main():
    name = sys.argv[1] if len(sys.argv) > 1 else "ns2d"
    if REF is None:
        print("SKIP: no reference tree (run oracle/vendor_ref.py in the build container)")
        return 0
    # this repository FIRST (modules/, utils.py), the reference second (dataset/, training_utils.py, the script itself)
    sys.path[:0] = [ROOT, os.path.join(ROOT, "oracle"), REF]
    for pkg, subs in (("matplotlib", ("pyplot",)), ("mpl_toolkits", ("axes_grid1",)), ("xarray", ()), ("zarr", ()), ("wandb", ())):
        if importlib.util.find_spec(pkg) is None:   # plotting / IO packages the scripts import but the rollout never touches
            m = types.ModuleType(pkg)
            m.__spec__ = importlib.machinery.ModuleSpec(pkg, None)
            sys.modules[pkg] = m
            for sname in subs:
                sm = types.ModuleType(f"{pkg}.{sname}")
                sm.__spec__ = importlib.machinery.ModuleSpec(f"{pkg}.{sname}", None)
                sys.modules[f"{pkg}.{sname}"] = sm
                setattr(m, sname, sm)
            if pkg == "mpl_toolkits":
                m.axes_grid1.ImageGrid = object
    import torch
    import yaml
    import modules
    assert os.path.dirname(os.path.abspath(modules.__file__)) == os.path.join(ROOT, "modules"), "the drop-in package must win"
    from utils import dict2namespace
    yml, script = SCRIPTS[name]
    cfg = dict2namespace(yaml.safe_load(open(os.path.join(REF, yml))))
    if not hasattr(cfg, "disable_coarse_attn"):
        cfg.disable_coarse_attn = None   # read by the decoder but absent from two YAMLs
    spec = importlib.util.spec_from_file_location("unmodified_" + script[:-3], os.path.join(REF, script))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)   # the __main__ guard keeps the script inert
    torch.manual_seed(1234)
    theirs = mod.LatentDynamics(cfg).eval().to("cuda:0")          # the script's own class, SimpleCNN and all
    import lns_oracle as O
    from lns_b200.configs import get_config
    from lns_b200.latent_dynamics import LatentDynamics
    from lns_b200.rollout import Rollout
    mycfg = get_config(name)
    torch.manual_seed(1234)
    ours = LatentDynamics(mycfg).eval().to("cuda:0")
    sd_t, sd_o = theirs.state_dict(), ours.state_dict()
    assert list(sd_t) == list(sd_o), "state_dict keys differ"
    assert all(torch.equal(sd_t[k], sd_o[k]) for k in sd_t), "same seed must give the same parameters"
    ours.load_state_dict(sd_t, strict=True)
    theirs.load_state_dict(sd_o, strict=True)
    x, param = O.make_inputs(mycfg, 4, seed=3)
    x = x.to("cuda:0")
    param = param.to("cuda:0") if param is not None else None
    with torch.no_grad():
        a = Rollout(theirs, batch=4, steps=3, to_x=True).build()(x, param).clone()
        b = Rollout(ours, batch=4, steps=3, to_x=True).build()(x, param).clone()
        # and the script's own predict(): its Python loop over OUR autoencoder's forward + its own torch propagator
        args = (x, 2) if param is None else (x, 2, param)
        y = theirs.predict(*args, to_x=True)
    assert tuple(y.shape) == (4, 2) + tuple(a.shape[2:])
    diff = (a - b).abs().max().item()
    print(f"unmodified {script}: Rollout output {tuple(a.shape)}, max |theirs - ours| = {diff:.3e}; script.predict OK {tuple(y.shape)}")
    assert diff == 0.0
    print("OK")
    return 0


if __name__ == "__main__":
    sys.exit(main())
