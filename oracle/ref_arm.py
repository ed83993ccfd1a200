"""BASELINE INFRASTRUCTURE (executed only by bench.py's reference / cpu_baseline / eager_gpu_baseline legs, always in its own
process): times the reference's OWN implementation of the hot path -- ``LatentDynamics.predict(x, R, to_x=True)`` of the
unmodified reference tree (train_stage2_ns2d.py:143-158 and the SW / two-phase / conditional variants), loaded through
oracle/ref_shims.py from /root/reference or from the git-ignored copy oracle/_ref that oracle/vendor_ref.py stages.

    python oracle/ref_arm.py --workload ns2d --device cpu  --batch 8   --rollout-steps 20 --steps 5 --warmup 1
    python oracle/ref_arm.py --workload ns2d --device cuda --batch 256 --rollout-steps 20 --steps 3 --warmup 1 [--autocast]

Prints ONE JSON object: {"value": trajectory-steps/s, "ms_per_step", "kind": "reference" | "port", "cores", "sample", ...}.
If neither copy of the reference exists the CPU restatement oracle/lns_oracle.py is timed instead (kind "port").
It runs in a separate process because the reference's package is also called ``modules`` (the drop-in boundary), and because
torchrun exports OMP_NUM_THREADS=1 to its workers: bench.py starts this script with a clean thread environment."""
import argparse
import json
import os
import sys
import time

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)

R_DEFAULT = {"ns2d": 20, "sw": 20, "twophase": 20, "twophase_cond": 50}
CPU_SAMPLE = {"ns2d": 8, "sw": 1, "twophase": 2, "twophase_cond": 1}


def load_cfg(name):
    import importlib.util
    p = os.path.join(os.path.dirname(HERE), "lns-latent-neural-pde-solver_b200", "configs.py")
    spec = importlib.util.spec_from_file_location("lns_b200_cfg", p)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod.get_config(name)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="ns2d")
    ap.add_argument("--device", default="cpu")
    ap.add_argument("--batch", type=int, default=None)
    ap.add_argument("--rollout-steps", type=int, default=None)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=1)
    ap.add_argument("--autocast", action="store_true", help="torch.autocast(bfloat16) around predict (the same-box bf16 bar)")
    ap.add_argument("--budget-s", type=float, default=150.0, help="stop timing further steps once this much time has been spent")
    args = ap.parse_args()

    import torch
    import lns_oracle as O
    import ref_shims as RS
    name = args.workload
    R = args.rollout_steps or R_DEFAULT[name]
    B = args.batch or CPU_SAMPLE[name]
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    dev = torch.device(args.device)
    mycfg = load_cfg(name)
    x, param = O.make_inputs(mycfg, B, seed=0)
    x = x.to(dev)
    param = param.to(dev) if param is not None else None
    if RS.available():
        kind = "reference"
        model, _ = RS.build_reference(name, seed=1234)
        model.load_state_dict(O.randomize_zero_init(model.state_dict()), strict=True)
        model = model.to(dev).eval()
        call_args = (x, R) if param is None else (x, R, param)

        def step():
            return model.predict(*call_args, to_x=True)
        what = f"unmodified reference LatentDynamics.predict (from {os.path.relpath(RS.REF, os.path.dirname(HERE))})"
    else:
        kind = "port"
        sys.path.insert(0, os.path.dirname(HERE))
        from lns_b200.latent_dynamics import LatentDynamics  # parameter container only
        torch.manual_seed(1234)
        sd = {k: v.to(dev) for k, v in O.randomize_zero_init(LatentDynamics(mycfg).state_dict()).items()}

        def step():
            return O.predict(sd, mycfg, x, R, param=param, to_x=True)
        what = "oracle port of LatentDynamics.predict (reference tree not staged)"

    def sync():
        if dev.type == "cuda":
            torch.cuda.synchronize(dev)

    ctx = torch.autocast(device_type=dev.type, dtype=torch.bfloat16) if args.autocast else torch.no_grad()
    times = []
    t_begin = time.perf_counter()
    with torch.no_grad(), ctx:
        for _ in range(max(0, args.warmup)):
            step()
            sync()
            if time.perf_counter() - t_begin > args.budget_s:
                break
        for _ in range(max(1, args.steps)):
            sync()
            t0 = time.perf_counter()
            y = step()
            sync()
            times.append(time.perf_counter() - t0)
            if time.perf_counter() - t_begin > args.budget_s:
                break
    total = sum(times)
    n = len(times)
    out = {
        "value": B * R * n / total, "unit": "trajectory-steps/s", "ms_per_step": 1e3 * total / n, "best_ms": 1e3 * min(times),
        "steps_run": n, "kind": kind, "cores": cores if dev.type == "cpu" else 0, "device": str(dev),
        "dtype": "bf16 autocast" if args.autocast else "f32", "sample_batch": B, "rollout_steps": R,
        "output_shape": list(y.shape),
        "sample": f"{what}, {'autocast(bfloat16)' if args.autocast else 'fp32'}, {B} trajectories x {R} steps per call, "
                  f"{n} timed calls after {args.warmup} warm-up ({1e3 * total / n:.1f} ms each), torch {torch.__version__} "
                  + (f"CPU, {cores} threads" if dev.type == "cpu" else f"eager on {torch.cuda.get_device_name(dev)}"),
    }
    print(json.dumps(out), flush=True)


if __name__ == "__main__":
    main()
