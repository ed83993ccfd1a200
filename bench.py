#!/usr/bin/env python
"""bench.py -- latent-rollout throughput (trajectory-steps/s) of the LNS hot path on N B200s of one node.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload ns2d|sw|twophase|twophase_cond]
    torchrun --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...      (one rank per GPU, NCCL)

One "step" = one full rollout of the per-GPU batch: autoencoder encode -> R autoregressive propagator steps -> decode of
all R states (LatentDynamics.predict(x, R, to_x=True) of the reference), i.e. B*R trajectory-steps per GPU per step.
Default workload = BASELINE.json's NS2d 64x64 large-batch rollout (config 5: R=20, B=1184 = 8 x 148 SMs trajectories per GPU --
inside the config's 256..8192 sweep, sized so the per-sample kernels run whole waves --, trajectory-sharded, weak scaling).
Default precision = 'fp16s', the 16-bit tensor-core mode that meets the north_star's <= 2e-3 per-stage bound (IEEE-half
operands, split hi+lo operands on the layers that carry the rounding error; `parity` reports the measured error of THIS run).

Output: ONE JSON line with `value` (inputs resident in HBM, CUDA-graph replay, CUDA-event timing, max over ranks), `e2e`
(same metric through the public API with pinned host buffers, H2D and D2H copies inside the timed region, every timed
step), `roofline` (the kernel with the largest share of THIS run's step time, from a CUDA-event-bracketed eager pass of the
same step), `roofline_by_time`, `parity`, `cpu_baseline` (the unmodified reference on the host cores), `eager_gpu_baseline`
(the unmodified reference in PyTorch eager on the same B200, fp32 and autocast(bfloat16)), `workloads` (the other three
configurations), `sweep` (NS2d batch sweep), `clocks`, `gpu_launches`.  The extras run on rank 0 at N=1 only.
`--impl reference` times the reference's own LatentDynamics.predict (oracle/_ref, staged by oracle/vendor_ref.py) on the host
cores on a bounded sample of the same workload.
"""
import argparse
import importlib.util
import collections
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "latent rollout trajectory-steps/sec"
WORKLOADS = {
    # name: (config, rollout steps R, trajectories per GPU, GFLOP per trajectory-step, label)
    "ns2d": ("ns2d", 20, 1184, 1.4611, "NS2d 64x64 latent rollout, R=20, 1184 trajectories/GPU = 8 per SM (BASELINE config 5)"),
    "sw": ("sw", 20, 64, 8.5138, "shallow water 96x192 rollout, R=20, 64 trajectories/GPU (BASELINE config 2)"),
    "twophase": ("twophase", 20, 128, 2.5532, "two-phase 61x121 rollout, R=20, 128 trajectories/GPU (BASELINE config 3)"),
    "twophase_cond": ("twophase_cond", 50, 128, 2.4806,
                      "two-phase conditional rollout, R=50, 128 trajectories/GPU (BASELINE config 4)"),
}
DTYPE_OF = {"bf16": "bf16", "fp16": "f16", "fp16s": "f16", "tf32": "tf32", "fp32": "f32"}
CPU_SAMPLE = {"ns2d": 8, "sw": 1, "twophase": 2, "twophase_cond": 1}       # trajectories per CPU call of the reference
GPU_EAGER_SAMPLE = {"ns2d": 256, "sw": 16, "twophase": 32, "twophase_cond": 32}


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="ns2d", choices=sorted(WORKLOADS))
    ap.add_argument("--batch", type=int, default=None, help="trajectories per GPU (default: workload's)")
    ap.add_argument("--rollout-steps", type=int, default=None)
    ap.add_argument("--precision", default="fp16s", choices=["fp16s", "bf16", "fp16", "tf32", "fp32"])
    ap.add_argument("--quick", action="store_true", help="skip the extras (baselines, other workloads, sweep, per-kernel pass)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-gather", action="store_true", help="skip the final all-gather of fields (step i overlaps the rollout of step i+1) for N > 1")
    ap.add_argument("--decode-chunk", type=int, default=None, help="samples per decode launch group (default: engine's)")
    return ap.parse_args()


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return {"hbm_gbs": d["hbm_gbs"], "bf16_tflops": d["bf16_tflops"],
                "bf16_tflops_sustained": d.get("bf16_tflops_sustained", d["bf16_tflops"]), "source": "MEASURED_PEAKS.json"}
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0,
            "source": "fallback of B200_PROFILING.md (MEASURED_PEAKS.json absent)"}


_GATHER_KIND = [None]  # class name of the gather object bench.py actually built (P2PGather | OverlappedGather)


def _gather_label():
    return ("copy-engine pushes into CUDA symmetric memory (lns_b200.dist.P2PGather)" if _GATHER_KIND[0] == "P2PGather"
            else "NCCL all_gather_into_tensor")


def config_block(label, R, B, world, gather, precision, sample_batch=None):
    """the `config` object -- identical keys on both arms (the reference arm adds its bounded sample)"""
    c = {"workload": label, "rollout_steps": R, "trajectories_per_gpu": B, "global_batch": B * world,
         "parallelism": f"trajectory-sharded x{world}" + (f", final all-gather of fields by {_gather_label()} (step i overlaps the rollout of step i+1)" if gather else ""),
         "l2": "per-step working set (activations) is far larger than the 126 MB L2; no explicit flush",
         "cuda_graph": True, "random_init_weights_seed": 1234, "precision_mode": precision}
    if sample_batch is not None:
        c["sample_batch"] = sample_batch
    return c


# ---- the reference itself, in its own process -------------------------------------------------------------------------------------
def run_ref_arm(workload, device, batch, R, steps, warmup, autocast=False, budget_s=150.0, timeout=900):
    """oracle/ref_arm.py in a fresh interpreter with a clean thread environment (torchrun exports OMP_NUM_THREADS=1, which
    throttled round 1's CPU leg 100x) -> its JSON dict, or {"error": ...}."""
    env = {k: v for k, v in os.environ.items()
           if k not in ("OMP_NUM_THREADS", "MKL_NUM_THREADS", "RANK", "LOCAL_RANK", "WORLD_SIZE", "MASTER_ADDR", "MASTER_PORT",
                        "TORCHELASTIC_RUN_ID", "GROUP_RANK", "ROLE_RANK", "LOCAL_WORLD_SIZE", "ROLE_WORLD_SIZE")}
    cmd = [sys.executable, os.path.join(ROOT, "oracle", "ref_arm.py"), "--workload", workload, "--device", device, "--batch", str(batch),
           "--rollout-steps", str(R), "--steps", str(steps), "--warmup", str(warmup), "--budget-s", str(budget_s)]
    if autocast:
        cmd.append("--autocast")
    try:
        r = subprocess.run(cmd, capture_output=True, text=True, timeout=timeout, env=env)
        for ln in reversed(r.stdout.strip().splitlines()):
            if ln.startswith("{"):
                return json.loads(ln)
        return {"error": (r.stderr or r.stdout)[-400:]}
    except Exception as ex:  # noqa: BLE001
        return {"error": repr(ex)[:400]}


def cpu_baseline(workload, R, steps=3, warmup=1, budget_s=25.0):
    d = run_ref_arm(workload, "cpu", CPU_SAMPLE[workload], R, steps, warmup, budget_s=budget_s)
    if "error" in d:
        return {"value": None, "unit": "trajectory-steps/s", "cores": os.cpu_count(), "kind": "reference", "sample": "failed: " + d["error"]}
    return {"value": round(d["value"], 3), "unit": "trajectory-steps/s", "cores": d["cores"], "kind": d["kind"], "sample": d["sample"],
            "ms_per_call": round(d["ms_per_step"], 2), "sample_batch": d["sample_batch"]}


def eager_gpu_baseline(workload, R):
    """the same-box GPU bar: the unmodified reference modules in PyTorch eager on this B200 (fp32 and autocast(bfloat16))"""
    out = {}
    for key, ac in (("fp32", False), ("autocast_bf16", True)):
        d = run_ref_arm(workload, "cuda", GPU_EAGER_SAMPLE[workload], R, 3, 1, autocast=ac, budget_s=60.0, timeout=400)
        out[key] = ({"error": d["error"]} if "error" in d else
                    {"value": round(d["value"], 1), "unit": "trajectory-steps/s", "ms_per_call": round(d["ms_per_step"], 2),
                     "sample_batch": d["sample_batch"], "kind": d["kind"], "sample": d["sample"]})
    return out


# ---- parity (the oracle is only ever the checker) ----------------------------------------------------------------------------------
def parity_check(model, cfg, precision, device, batch=8, drift_steps=0):
    """Teacher-forced per-stage relative L2 of the GPU path (at `precision`, through the same engines as the timed run: batch >= 8
    puts the latent-grid conv engine on the checked path) against the fp64 oracle; optionally the free-running drift."""
    import torch
    import lns_oracle as O
    from lns_b200 import ops
    from lns_b200.rollout import Rollout
    sd64 = O.to_dtype({k: v.detach().cpu() for k, v in model.state_dict().items()}, torch.float64)
    x, param = O.make_inputs(cfg, batch, seed=77)
    ae = O.ae_name(cfg)
    z_ref = O.encode(sd64, cfg, x.double(), ae)
    cond = O.cond_embedding(sd64, cfg, param.double(), torch.float64) if param is not None else None
    z1_ref = O.propagator_step(sd64, cfg, z_ref, cond)
    y_ref = O.decode(sd64, cfg, z1_ref, ae)
    with torch.no_grad(), ops.precision(precision):
        z = model.autoencoder.encode(x.to(device))
        z1 = model.propagator(z_ref.float().to(device)) if param is None else \
            model.propagator(z_ref.float().to(device), param.to(device))
        y = model.autoencoder.decode(z1_ref.float().to(device))
    out = {"vs": f"fp64 oracle of the reference, teacher-forced per stage, {batch} trajectories, max rel-L2 (bound 2e-3; fp32 mode 1e-5)",
           "encode": float(O.rel_l2(z.cpu(), z_ref).max()), "propagator_step": float(O.rel_l2(z1.cpu(), z1_ref).max()),
           "decode": float(O.rel_l2(y.cpu(), y_ref).max())}
    if drift_steps:
        nb = min(batch, 4)
        yk_ref = O.predict(sd64, cfg, x[:nb].double(), drift_steps, param=None if param is None else param[:nb].double(), to_x=True)
        ro = Rollout(model, batch=nb, steps=drift_steps, to_x=True, precision=precision, use_graph=False)
        with torch.no_grad():
            yk = ro(x[:nb].to(device), None if param is None else param[:nb].to(device)).cpu()
        out["free_running_field_rel_l2_per_step"] = [round(float(O.rel_l2(yk[:, t], yk_ref[:, t]).max()), 6) for t in range(drift_steps)]
    return out


# ---- clocks ----------------------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.samples, self.stop_flag, self.thread = index, [], False, None

    def _loop(self):
        while not self.stop_flag:
            try:
                out = subprocess.run(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-i",
                                      str(self.index)], capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.samples.append([c.strip() for c in out.split(",")])
            except Exception:
                pass
            time.sleep(0.1)

    def start(self):
        self.thread = threading.Thread(target=self._loop, daemon=True)
        self.thread.start()

    def stop(self):
        self.stop_flag = True
        if self.thread:
            self.thread.join(timeout=6)
        sm = sorted(int(s[0]) for s in self.samples if s and s[0].isdigit())
        mx = max([int(s[1]) for s in self.samples if len(s) > 1 and s[1].isdigit()] or [0])
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for s in self.samples for i in range(4) if len(s) > 2 + i and s[2 + i] == "Active"})
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx or None, "reasons": reasons,
                "samples": len(sm)}


# ---- per-kernel roofline of THIS run ------------------------------------------------------------------------------------------
KERNEL_OF = [  # timeline label prefix -> kernel (csrc file)
    ("conv e0", "conv_simt_kernel / lift1x1 (conv_simt.cu)"), ("conv e1", "conv_umma_kernel (conv_umma.cu, tcgen05 gather engine)"),
    ("conv e2", "conv_halo_kernel (conv_halo.cu, tcgen05 halo engine)"), ("conv e3", "conv_latent_kernel (conv_latent.cu, tcgen05)"),
    ("conv e4", "conv_coarse_kernel (conv_coarse.cu, tcgen05 block-halo engine)"), ("fablock_full", "fablock_full2_kernel (fablock_full.cu; fablock_full_kernel for shapes outside 16/32)"),
    ("fablock_tc", "fablock_tc_kernel (fablock_tc.cu, tcgen05)"), ("fa_axis", "fa_axis_kernel (fa_axis.cu)"),
    ("sablock_fused", "sablock_fused_kernel (sablock_fused.cu)"), ("ffn_fused", "ffn_fused_kernel (ffn_fused.cu, tcgen05)"),
    ("proj", "pointwise_proj64_kernel (pointwise.cu)"), ("gn_stats", "gn_affine_small / chan_stats kernels (pointwise.cu, norm.cu)"),
    ("gn_act_fused", "gn_act_small_kernel (pointwise.cu)"), ("affine_act", "affine_act2_kernel (norm.cu)"),
    ("fablock_prepass", "fablock_prepass2_kernel (fablock.cu)"),
]


def kernel_name(label):
    for pre, k in KERNEL_OF:
        if label.startswith(pre):
            return k
    return label.split(" ")[0]


def per_kernel_pass(torch, ops, model, B, R, precision, x_dev, p_dev, peaks, decode_chunk):
    """One eager, serial-order rollout of the SAME step with a CUDA-event pair around every library call (on the launching
    stream) -> per call-site time, algorithmic flops and bytes.  `roofline` = the call site with the largest time share."""
    from lns_b200.rollout import Rollout
    ro = Rollout(model, batch=B, steps=R, to_x=True, precision=precision, use_graph=False, pipeline=False, decode_chunk=decode_chunk)
    with torch.no_grad():
        ro.build()
        ro(x_dev, p_dev)
        torch.cuda.synchronize()
        ops._state.timeline = []
        ro(x_dev, p_dev)
        torch.cuda.synchronize()
    tl, ops._state.timeline = ops._state.timeline, None
    agg = collections.OrderedDict()
    for label, e0, e1, fl, by in tl:
        a = agg.setdefault(label, [0, 0.0, 0.0, 0.0])
        a[0] += 1
        a[1] += e0.elapsed_time(e1)
        a[2] += fl
        a[3] += by
    total = sum(a[1] for a in agg.values())
    rows = []
    for label, (n, ms, fl, by) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        t = ms * 1e-3
        tf, gbs = fl / t / 1e12, by / t / 1e9
        # the bound that a perfect kernel of this layer would hit first at the measured peaks
        t_tensor, t_hbm = fl / (peaks["bf16_tflops_sustained"] * 1e12), by / (peaks["hbm_gbs"] * 1e9)
        bound = "tensor" if t_tensor >= t_hbm else "hbm"
        row = {"call_site": label, "kernel": kernel_name(label), "share_of_step": round(ms / total, 4), "launches": n,
               "avg_launch_ms": round(ms / n, 4), "bound": bound,
               "achieved": round(tf if bound == "tensor" else gbs, 2), "unit": "TFLOP/s" if bound == "tensor" else "GB/s",
               "peak": peaks["bf16_tflops_sustained"] if bound == "tensor" else peaks["hbm_gbs"],
               "algorithmic_flops_per_launch": fl / n, "algorithmic_bytes_per_launch": by / n}
        row["frac"] = round(row["achieved"] / row["peak"], 4)
        rows.append(row)
    del ro
    torch.cuda.empty_cache()
    return rows, total


def traffic_of(call_site):
    """DRAM read + write bytes per launch of this call site from the committed `ncu --set full` summary, if one exists"""
    tp = os.path.join(ROOT, "profiles", "r02_traffic.json")
    if os.path.exists(tp):
        try:
            return json.load(open(tp)).get(call_site)
        except Exception:  # noqa: BLE001
            return None
    return None


def time_rollout(torch, ro, x_dev, p_dev, steps, warmup=3):
    for _ in range(warmup):
        ro(x_dev, p_dev)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        ro(x_dev, p_dev)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps


def side_workload(torch, O, name, precision, device, batch=None, steps=3, parity=True):
    """value (+ live parity) of another configuration on this GPU: CUDA-graph replay, inputs resident, CUDA events"""
    from lns_b200.configs import get_config
    from lns_b200.latent_dynamics import LatentDynamics
    from lns_b200.rollout import Rollout
    cfg_name, R, B, gflop, label = WORKLOADS[name]
    B = batch or B
    cfg = get_config(cfg_name)
    torch.manual_seed(1234)
    model = LatentDynamics(cfg).eval()
    model.load_state_dict(O.randomize_zero_init(model.state_dict()))
    model = model.to(device)
    x, p = O.make_inputs(cfg, B, seed=0)
    x, p = x.to(device), (p.to(device) if p is not None else None)
    ro = Rollout(model, batch=B, steps=R, to_x=True, precision=precision, use_graph=True)
    with torch.no_grad():
        ro.build()
        ms = time_rollout(torch, ro, x, p, steps)
    out = {"workload": label, "value": round(B * R / (ms * 1e-3), 1), "unit": "trajectory-steps/s", "ms_per_step": round(ms, 3),
           "trajectories_per_gpu": B, "rollout_steps": R, "launches_per_step": int(ro.launches_per_call),
           "whole_path_tflops": round(B * R / (ms * 1e-3) * gflop / 1e3, 2)}
    if parity:
        out["parity"] = parity_check(model, cfg, precision, device, batch=4 if name == "sw" else 8)
    del ro, model
    torch.cuda.empty_cache()
    return out


# ---- main ---------------------------------------------------------------------------------------------------------------------
def main():
    args = parse()
    cfg_name, R, B, gflop_per_ts, label = WORKLOADS[args.workload]
    R = args.rollout_steps or R
    B = args.batch or B
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))

    if args.impl == "reference":
        # the reference's own CPU implementation of the path on the host cores (rank 0 only), each step = one predict() on a
        # bounded sample of the workload
        if rank != 0:
            return 0
        sb = CPU_SAMPLE[cfg_name]
        d = run_ref_arm(cfg_name, "cpu", sb, R, max(1, args.steps), max(0, args.warmup), budget_s=170.0)
        if "error" in d:
            print(json.dumps({"impl": "reference", "unavailable": d["error"][:200]}), flush=True)
            return 0
        base = {"value": round(d["value"], 3), "unit": "trajectory-steps/s", "cores": d["cores"], "kind": d["kind"], "sample": d["sample"]}
        line = {"metric": METRIC, "impl": "reference", "value": round(d["value"], 3), "unit": "trajectory-steps/s",
                "n_gpus": args.gpus, "steps": d["steps_run"], "warmup": args.warmup, "ms_per_step": round(d["ms_per_step"], 3),
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": config_block(label, R, B, args.gpus, False, "fp32 (reference, CPU)", sample_batch=sb),
                "cpu_baseline": base,
                "e2e": {"value": round(d["value"], 3), "unit": "trajectory-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                "gpu_launches": 0}
        print(json.dumps(line), flush=True)
        return 0

    import torch
    import torch.distributed as dist
    from lns_b200 import ops
    from lns_b200.configs import get_config
    from lns_b200.dist import OverlappedGather, init_from_env, make_gather
    from lns_b200.latent_dynamics import LatentDynamics
    from lns_b200.rollout import Rollout
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import lns_oracle as O  # only for seeded synthetic inputs / zero-init redraw and the parity leg

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (there is no CPU path; use --impl reference for the CPU baseline)")
    rank, local, world = init_from_env("nccl")
    device = torch.device("cuda", local)
    torch.cuda.set_device(device)
    peaks = load_peaks()

    cfg = get_config(cfg_name)
    torch.manual_seed(1234)
    model = LatentDynamics(cfg).eval()
    model.load_state_dict(O.randomize_zero_init(model.state_dict()))
    model = model.to(device)
    x_cpu, p_cpu = O.make_inputs(cfg, B, seed=rank)
    x_pin = x_cpu.pin_memory()
    p_pin = p_cpu.pin_memory() if p_cpu is not None else None
    x_dev = x_cpu.to(device)
    p_dev = p_cpu.to(device) if p_cpu is not None else None

    ro = Rollout(model, batch=B, steps=R, to_x=True, precision=args.precision, use_graph=True,
                 decode_chunk=args.decode_chunk)
    with torch.no_grad():
        ro.build()
        ro(x_dev, p_dev)
    torch.cuda.synchronize()
    gather = world > 1 and not args.no_gather
    # N > 1: the final all-gather of the predicted fields (NCCL over NVLink) of step i runs on NCCL's stream out of a staging
    # copy while the rollout of step i+1 computes; the timed region ends when the last gather has completed.
    og = make_gather((B, R, ro.C, ro.Ly, ro.Lx), torch.float32, device) if gather else None
    _GATHER_KIND[0] = type(og).__name__ if og is not None else None

    def one_step():
        out = ro(x_dev, p_dev)
        if og is not None:
            og.submit(out)
        return out

    def drain():
        if og is not None:
            og.wait()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    warmup = max(args.warmup, 3)
    for _ in range(warmup):
        one_step()
    drain()
    sampler = ClockSampler(local)
    barrier()
    if rank == 0:
        sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for _ in range(args.steps):
        one_step()
    drain()
    e1.record()
    barrier()
    elapsed_ms = e0.elapsed_time(e1)
    clocks = sampler.stop() if rank == 0 else None
    if world > 1:
        t = torch.tensor([elapsed_ms], device=device)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        elapsed_ms = float(t.item())
    value = world * B * R * args.steps / (elapsed_ms * 1e-3)

    # end to end through the public API: pinned host input -> H2D -> rollout -> D2H of the predicted fields, EVERY timed step.
    # The D2H copy of step i runs on a copy stream out of a device staging buffer while the rollout of step i+1 computes
    # (double-buffered, event-ordered); the timed region ends when the last copy has landed.
    out_host = [torch.empty((B, R, ro.C, ro.Ly, ro.Lx), dtype=torch.float32).pin_memory() for _ in range(2)]
    stage = [torch.empty((B, R, ro.C, ro.Ly, ro.Lx), dtype=torch.float32, device=device) for _ in range(2)]
    copy_stream = torch.cuda.Stream(device)
    compute = torch.cuda.current_stream(device)
    staged = [torch.cuda.Event() for _ in range(2)]
    landed = [torch.cuda.Event() for _ in range(2)]
    for ev in landed:
        ev.record(compute)
    e2e_steps = max(2, args.steps)

    def e2e_step(i):
        s = i & 1
        xd = x_pin.to(device, non_blocking=True)
        pd = p_pin.to(device, non_blocking=True) if p_pin is not None else None
        out = ro(xd, pd)
        compute.wait_event(landed[s])          # the copy that last read this staging slot has finished
        stage[s].copy_(out, non_blocking=True)
        staged[s].record(compute)
        with torch.cuda.stream(copy_stream):
            copy_stream.wait_event(staged[s])
            out_host[s].copy_(stage[s], non_blocking=True)
            landed[s].record(copy_stream)

    e2e_step(0)
    compute.wait_stream(copy_stream)
    barrier()
    e0.record()
    for i in range(e2e_steps):
        e2e_step(i)
    compute.wait_stream(copy_stream)
    e1.record()
    barrier()
    e2e_ms = e0.elapsed_time(e1)
    if world > 1:
        t = torch.tensor([e2e_ms], device=device)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_ms = float(t.item())
    e2e_value = world * B * R * e2e_steps / (e2e_ms * 1e-3)
    h2d = x_pin.numel() * 4 + (p_pin.numel() * 4 if p_pin is not None else 0)
    d2h = out_host[0].numel() * 4
    del out_host, stage
    torch.cuda.empty_cache()

    if rank == 0:
        tflops = value * gflop_per_ts / 1e3
        line = {
            "metric": METRIC, "value": round(value, 1), "unit": "trajectory-steps/s",
            "n_gpus": world, "steps": args.steps, "warmup": warmup,
            "ms_per_step": round(elapsed_ms / args.steps, 4), "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": DTYPE_OF[args.precision], "data": "synthetic",
            "config": config_block(label, R, B, world, gather, args.precision),
            "whole_path_tflops_per_gpu": round(tflops / world, 2),
            "whole_path_frac_of_bf16_sustained": round(tflops / world / peaks["bf16_tflops_sustained"], 4),
            "e2e": {"value": round(e2e_value, 1), "unit": "trajectory-steps/s", "h2d_bytes_per_step": h2d,
                    "d2h_bytes_per_step": d2h, "steps": e2e_steps},
            "gpu_launches": int(ro.launches_per_call * args.steps),
            "launches_per_step": int(ro.launches_per_call),
            "clocks": clocks, "roofline": None,
        }
        extras = world == 1 and not args.quick
        if extras:
            # the kernel with the largest share of THIS run's step (eager, serial order, one CUDA-event pair per library call)
            rows, total = per_kernel_pass(torch, ops, model, B, R, args.precision, x_dev, p_dev, peaks, args.decode_chunk)
            top = dict(rows[0])
            top["traffic"] = traffic_of(top["call_site"])
            top["peak_source"] = peaks["source"] + " (sustained bf16 / HBM copy: the kernel is timed inside a long step)"
            top["timed"] = (f"CUDA events around every launch of one eager rollout of this step on the launching stream "
                            f"({total:.1f} ms in serial order vs {elapsed_ms / args.steps:.1f} ms as a two-stream CUDA graph)")
            line["roofline"] = top
            line["roofline_by_time"] = rows[:8]
            fam = collections.defaultdict(float)
            for r_ in rows:
                fam[r_["kernel"]] += r_["share_of_step"]
            line["share_by_kernel"] = {k: round(v, 4) for k, v in sorted(fam.items(), key=lambda kv: -kv[1])[:10]}
            line["parity"] = parity_check(model, cfg, args.precision, device, batch=8, drift_steps=R if cfg_name == "ns2d" else 0)
            if not args.no_cpu_baseline:
                line["cpu_baseline"] = cpu_baseline(cfg_name, R)
                line["eager_gpu_baseline"] = eager_gpu_baseline(cfg_name, R)
            del ro
            torch.cuda.empty_cache()
            if args.workload == "ns2d":
                others = {}
                for w in ("sw", "twophase", "twophase_cond"):
                    try:
                        others[w] = side_workload(torch, O, w, args.precision, device)
                    except Exception as ex:  # noqa: BLE001
                        others[w] = {"error": repr(ex)[:300]}
                line["workloads"] = others
                sweep = {}
                for nb in (256, 1024, 4096, 8192):
                    try:
                        s_ = side_workload(torch, O, "ns2d", args.precision, device, batch=nb, steps=2, parity=False)
                        sweep[str(nb)] = {"value": s_["value"], "ms_per_step": s_["ms_per_step"]}
                    except Exception as ex:  # noqa: BLE001
                        sweep[str(nb)] = {"error": repr(ex)[:200]}
                line["sweep"] = {"workload": "NS2d 64x64, R=20, trajectories per GPU (BASELINE config 5)", "unit": "trajectory-steps/s", **sweep}
            # stand-alone legs outside the rollout: the spectral block (BASELINE config 4's wording) and the training rollout
            # (SURVEY 8(f) row 3); tools/bench_spectral.py, tools/bench_train.py
            for key, fname, kw in (("spectral", "bench_spectral.py", {"peak": peaks["hbm_gbs"]}), ("training", "bench_train.py", {})):
                try:
                    spec = importlib.util.spec_from_file_location(key + "_bench", os.path.join(ROOT, "tools", fname))
                    mod = importlib.util.module_from_spec(spec)
                    spec.loader.exec_module(mod)
                    line[key] = mod.run(**kw)
                except Exception as ex:  # noqa: BLE001
                    line[key] = {"error": repr(ex)[:300]}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
