"""TEST INFRASTRUCTURE ONLY -- never imported by the product path.

Loads the *unmodified* reference (BaratiLab/LNS-Latent-Neural-PDE-Solver) from
``/root/reference`` (build container) or from the git-ignored copy ``oracle/_ref``
that ``oracle/vendor_ref.py`` stages (GPU box), so that golden vectors can be
generated and the reference arm of ``bench.py`` (``--impl reference``) can time the
reference's own ``LatentDynamics.predict``.  Parity tests never depend on it: they
use the committed fixtures in ``tests/golden/`` and ``oracle/lns_oracle.py``.

The reference does not import as shipped; five shims are
installed in ``sys.modules`` -- no reference file is modified or copied:

1. ``modules.siren_module`` is imported (modules/basics.py:6,
   modules/factorized_attention.py:8) but missing -> stub with the two names.
2. ``utils.dict2namespace`` is imported by every train script
   (train_stage2_ns2d.py:16) but missing -> recursive dict->Namespace.
3. ``padding_mode`` is an undefined name inside the NS2d ``Encoder.__init__``
   (modules/autoencoder2d.py:32) -> resolved through the module global.
4. matplotlib / mpl_toolkits / xarray / zarr are not installed -> stubs that carry
   a ``__spec__`` (torch._dynamo's import scan requires one).
5. ``disable_coarse_attn`` is read (modules/autoencoder2d_nonsquared.py:170) but is
   absent from two YAMLs -> defaulted to None.
"""
import argparse
import importlib.machinery
import importlib.util
import os
import sys
import types

import yaml

_HERE = os.path.dirname(os.path.abspath(__file__))


def _default_root():
    """/root/reference in the build container; on the GPU box the copy that oracle/vendor_ref.py staged under oracle/_ref
    (git-ignored, travels with the gpurun snapshot)."""
    for cand in (os.environ.get("LNS_REFERENCE_ROOT"), "/root/reference", os.path.join(_HERE, "_ref")):
        if cand and os.path.isdir(os.path.join(cand, "modules")):
            return cand
    return "/root/reference"


REF = _default_root()

CONFIGS = {
    # name -> (yaml, stage-2 script, AE attribute on LatentDynamics)
    "ns2d": ("configs/ns2d_stage2_prop.yml", "train_stage2_ns2d.py", "vq_ae"),
    "sw": ("configs/SW_stage2_prop.yml", "train_stage2_SW.py", "vq_ae"),
    "twophase": ("configs/twophase_stage2_prop.yml", "train_stage2_twophase.py", "vq_ae"),
    "twophase_cond": ("configs/twophase_stage2_cond_prop.yml",
                      "train_stage2_twophase_conditional.py", "ae"),
}


def available():
    return os.path.isdir(os.path.join(REF, "modules"))


def dict2namespace(d):
    ns = argparse.Namespace()
    for k, v in d.items():
        setattr(ns, k, dict2namespace(v) if isinstance(v, dict) else v)
    return ns


def _stub(name, **attrs):
    m = types.ModuleType(name)
    m.__spec__ = importlib.machinery.ModuleSpec(name, None)
    m.__dict__.update(attrs)
    sys.modules[name] = m
    return m


_installed = False


def install():
    """Install the shims and put the reference root first on sys.path."""
    global _installed
    if _installed:
        return
    if not available():
        raise RuntimeError(f"reference tree not found at {REF}")
    # our own drop-in package is also called ``modules``; the reference must win here
    for k in [k for k in sys.modules if k == "modules" or k.startswith("modules.")]:
        del sys.modules[k]
    sys.path.insert(0, REF)
    _stub("modules.siren_module", SirenNet=type("SirenNet", (), {}),
          SirenWrapper=type("SirenWrapper", (), {}))
    _stub("utils", dict2namespace=dict2namespace)
    for pkg, subs in (("matplotlib", ("pyplot",)), ("mpl_toolkits", ("axes_grid1",)),
                      ("xarray", ()), ("zarr", ()), ("wandb", ())):
        if importlib.util.find_spec(pkg) is None:
            p = _stub(pkg)
            for s in subs:
                setattr(p, s, _stub(f"{pkg}.{s}"))
            if pkg == "mpl_toolkits":
                p.axes_grid1.ImageGrid = object
    _installed = True


def load_config(name):
    yml = CONFIGS[name][0]
    with open(os.path.join(REF, yml)) as f:
        cfg = dict2namespace(yaml.safe_load(f))
    if not hasattr(cfg, "disable_coarse_attn"):
        cfg.disable_coarse_attn = None
    return cfg


def load_script(name):
    install()
    script = CONFIGS[name][1]
    modname = "ref_" + script[:-3]
    if modname in sys.modules:
        return sys.modules[modname]
    if name == "ns2d":
        import modules.autoencoder2d as ae2d  # the reference's
        ae2d.padding_mode = "circular"  # cfg.is_periodic is True in the NS2d YAML
    spec = importlib.util.spec_from_file_location(modname, os.path.join(REF, script))
    mod = importlib.util.module_from_spec(spec)
    sys.modules[modname] = mod
    spec.loader.exec_module(mod)  # the __main__ guard keeps the script inert
    return mod


def build_reference(name, seed=1234):
    """-> (LatentDynamics in eval mode built under torch.manual_seed(seed), cfg)."""
    import torch
    cfg = load_config(name)
    mod = load_script(name)
    torch.manual_seed(seed)
    model = mod.LatentDynamics(cfg).eval()
    return model, cfg
