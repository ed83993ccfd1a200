"""Generate tests/golden/*.pt by running the UNMODIFIED reference (imported from /root/reference through
oracle/ref_shims.py) on seeded synthetic inputs.  Run in the build container only:

    python oracle/make_golden.py

Weights are NOT stored for the four LatentDynamics models: the drop-in constructors consume the torch RNG exactly like
the reference's, so ``torch.manual_seed(1234)`` reproduces them bit-for-bit (a per-tensor SHA-1 is stored and checked).
Fields are stored sub-sampled (every 2nd pixel) to keep the fixtures small; latents are stored in full.
"""
import copy
import hashlib
import os
import sys

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import ref_shims  # noqa: E402
import lns_oracle as O  # noqa: E402

OUT = os.path.join(os.path.dirname(HERE), "tests", "golden")
CASES = {"ns2d": (2, 3), "sw": (2, 2), "twophase": (2, 2), "twophase_cond": (2, 2)}  # (batch, steps)


def sha(t):
    return hashlib.sha1(t.detach().cpu().contiguous().numpy().tobytes()).hexdigest()


def main():
    os.makedirs(OUT, exist_ok=True)
    torch.set_num_threads(os.cpu_count())
    for name, (B, K) in CASES.items():
        model, cfg = ref_shims.build_reference(name, seed=1234)
        sd = O.randomize_zero_init(model.state_dict())
        model.load_state_dict(sd, strict=True)
        from lns_b200_cfg import get_config  # see bottom: loaded by path to avoid the `modules` name clash
        mycfg = get_config(name)
        x, param = O.make_inputs(mycfg, B, seed=0)
        args = (x, K) if param is None else (x, K, param)
        with torch.no_grad():
            y32 = model.predict(*args, to_x=True)
            z32 = model.predict(*args, to_x=False)
            m64 = copy.deepcopy(model).double()
            a64 = (x.double(), K) if param is None else (x.double(), K, param.double())
            if param is not None:
                # the reference's fourier_embedding always returns fp32 (modules/cond_utils.py:34 `.float()`), which
                # makes its fp64 copy fail in the first Linear; for the fp64 ground truth only, cast the embedding up
                script = ref_shims.load_script(name)
                orig = script.fourier_embedding
                script.fourier_embedding = lambda t, dim: orig(t, dim).double()
            y64 = m64.predict(*a64, to_x=True)
            z64 = m64.predict(*a64, to_x=False)
            if param is not None:
                script.fourier_embedding = orig
        fix = {
            "config": name, "batch": B, "steps": K, "weight_seed": 1234, "input_seed": 0,
            "state_sha1": {k: sha(v) for k, v in sd.items()},
            "field_fp32_sub2": y32[..., ::2, ::2].contiguous(), "latent_fp32": z32.contiguous(),
            "field_fp64_sub2": y64[..., ::2, ::2].contiguous(), "latent_fp64": z64.contiguous(),
            "field_norm_fp64": (y64 ** 2).sum(dim=(-1, -2, -3)).sqrt(),
            "ref_fp32_vs_fp64_rel_l2": float(O.rel_l2(y32.flatten(0, 1), y64.flatten(0, 1)).max()),
        }
        torch.save(fix, os.path.join(OUT, f"{name}_predict.pt"))
        print(name, "fields", tuple(y32.shape), "latents", tuple(z32.shape), "ref fp32 vs fp64 rel-L2 %.2e"
              % fix["ref_fp32_vs_fp64_rel_l2"])

    # stand-alone blocks that no shipped config instantiates (small instances, weights stored)
    ref_shims.install()
    import modules.basics as rb
    import modules.fourier_cond as rfc
    torch.manual_seed(7)
    blk = rb.FourierBasicBlock(8, 8, modes=[4, 5]).eval()
    cblk = rfc.CondFourierBasicBlock(8, 8, modes=[4, 5]).eval()
    g = torch.Generator().manual_seed(3)
    xs = torch.randn(2, 8, 13, 20, generator=g)
    emb = torch.randn(2, 8, generator=g)
    with torch.no_grad():
        fix = {
            "x": xs, "emb": emb,
            "fourier_sd": blk.state_dict(), "fourier_out": blk(xs),
            "fourier_out_fp64": copy.deepcopy(blk).double()(xs.double()),
            "cond_sd": cblk.state_dict(), "cond_out": cblk(xs, emb),
            "cond_out_fp64": copy.deepcopy(cblk).double()(xs.double(), emb.double()),
        }
    torch.save(fix, os.path.join(OUT, "fourier_blocks.pt"))
    print("fourier blocks", tuple(fix["fourier_out"].shape))


if __name__ == "__main__":
    # load lns_b200/configs.py by path (importing the lns_b200 package would also be fine, but keep this script free of
    # the product package)
    import importlib.util
    p = os.path.join(os.path.dirname(HERE), "lns-latent-neural-pde-solver_b200", "configs.py")
    spec = importlib.util.spec_from_file_location("lns_b200_cfg", p)
    mod = importlib.util.module_from_spec(spec)
    sys.modules["lns_b200_cfg"] = mod
    spec.loader.exec_module(mod)
    main()
