#!/usr/bin/env python
"""bench.py -- latent-rollout throughput (trajectory-steps/s) of the LNS hot path on N B200s of one node.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload ns2d|sw|twophase|twophase_cond]
    torchrun --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...      (one rank per GPU, NCCL)

One "step" = one full rollout of the per-GPU batch: autoencoder encode -> R autoregressive propagator steps -> decode of
all R states (LatentDynamics.predict(x, R, to_x=True) of the reference), i.e. B*R trajectory-steps per GPU per step.
Default workload = BASELINE.json's NS2d 64x64 large-batch rollout (config 5: R=20, B=1184 = 8 x 148 SMs trajectories per
GPU -- inside the config's 256..8192 sweep, sized so the per-sample kernels run whole waves --, trajectory-sharded, weak scaling), the configuration the metric's target is quoted on; other configs via --workload.

Output: ONE JSON line (see the contract in the task description) with `value` (inputs resident in HBM, CUDA-graph
replay, CUDA-event timing, max over ranks), `e2e` (same metric through the public API with pinned host buffers, H2D and
D2H copies inside the timed region), `roofline` (dominant kernel = the tcgen05 3x3 implicit-GEMM conv, timed alone with
CUDA events), `cpu_baseline` (the oracle port of the reference timed on the host cores), `clocks`, `gpu_launches`.
`--impl reference` times the reference's CPU implementation (oracle port; the reference is pure Python/PyTorch) on the
host cores on a bounded sample of the same workload.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (config, rollout steps R, trajectories per GPU, GFLOP per trajectory-step, label)
    "ns2d": ("ns2d", 20, 1184, 1.4611, "NS2d 64x64 latent rollout, R=20, 1184 trajectories/GPU = 8 per SM (BASELINE config 5)"),
    "sw": ("sw", 20, 64, 8.5138, "shallow water 96x192 rollout, R=20, 64 trajectories/GPU (BASELINE config 2)"),
    "twophase": ("twophase", 20, 128, 2.5532, "two-phase 61x121 rollout, R=20, 128 trajectories/GPU (BASELINE config 3)"),
    "twophase_cond": ("twophase_cond", 50, 128, 2.4806,
                      "two-phase conditional rollout, R=50, 128 trajectories/GPU (BASELINE config 4)"),
}


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="ns2d", choices=sorted(WORKLOADS))
    ap.add_argument("--batch", type=int, default=None, help="trajectories per GPU (default: workload's)")
    ap.add_argument("--rollout-steps", type=int, default=None)
    ap.add_argument("--precision", default="bf16", choices=["bf16", "fp16", "tf32", "fp32"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-gather", action="store_true", help="skip the final all-gather of fields (step i overlaps the rollout of step i+1) for N > 1")
    ap.add_argument("--decode-chunk", type=int, default=None, help="samples per decode launch group (default: engine's)")
    return ap.parse_args()


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return {"hbm_gbs": d["hbm_gbs"], "bf16_tflops": d["bf16_tflops"],
                "bf16_tflops_sustained": d.get("bf16_tflops_sustained", d["bf16_tflops"]), "source": "measured"}
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "source": "fallback"}


# ---- CPU reference (oracle port) ---------------------------------------------------------------------------------------
def cpu_reference_throughput(workload, budget_s=20.0, min_runs=2, max_runs=5):
    """trajectory-steps/s of the reference's CPU path (oracle port: same ATen CPU kernels in the same order as
    LatentDynamics.predict) on all host cores, on a bounded sample (small batch, full R)."""
    import torch
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import lns_oracle as O
    from lns_b200.configs import get_config
    from lns_b200.latent_dynamics import LatentDynamics
    cfg_name, R, _, _, _ = WORKLOADS[workload]
    cfg = get_config(cfg_name)
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    torch.manual_seed(1234)
    sd = O.randomize_zero_init(LatentDynamics(cfg).state_dict())
    B = {"ns2d": 8, "sw": 1, "twophase": 2, "twophase_cond": 1}[cfg_name]
    x, param = O.make_inputs(cfg, B, seed=0)
    t0 = time.perf_counter()
    O.predict(sd, cfg, x, R, param=param, to_x=True)  # warm-up (also sizes the sample)
    first = time.perf_counter() - t0
    runs = max(min_runs, min(max_runs, int(budget_s / max(first, 1e-3))))
    best = float("inf")
    for _ in range(runs):
        t0 = time.perf_counter()
        O.predict(sd, cfg, x, R, param=param, to_x=True)
        best = min(best, time.perf_counter() - t0)
    return {"value": B * R / best, "unit": "trajectory-steps/s", "cores": cores, "kind": "port",
            "sample": f"oracle port of LatentDynamics.predict, fp32, B={B}, R={R}, best of {runs} runs "
                      f"({best:.2f} s each), torch {torch.__version__} CPU, {cores} threads"}


def parity_check(model, cfg, precision, device):
    """Teacher-forced per-stage relative L2 of the GPU path (at `precision`) against the fp64 oracle on 2 trajectories: part of
    the cpu_baseline leg (the oracle is only ever the checker).  Returns {"encode", "propagator_step", "decode"}."""
    import torch
    import lns_oracle as O
    from lns_b200 import ops
    sd64 = O.to_dtype({k: v.detach().cpu() for k, v in model.state_dict().items()}, torch.float64)
    x, param = O.make_inputs(cfg, 2, seed=77)
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
    return {"vs": "fp64 oracle of the reference, teacher-forced per stage, 2 trajectories, max rel-L2",
            "encode": float(O.rel_l2(z.cpu(), z_ref).max()), "propagator_step": float(O.rel_l2(z1.cpu(), z1_ref).max()),
            "decode": float(O.rel_l2(y.cpu(), y_ref).max())}


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


# ---- dominant-kernel roofline ---------------------------------------------------------------------------------------------
def conv_roofline(torch, ops, device, peaks, workload, precision="bf16"):
    """Time the dominant kernel alone (tcgen05 implicit-GEMM 3x3 conv at the workload's largest layer) with CUDA events
    on the launching stream; operands are larger than L2 so every launch streams from HBM."""
    import math
    shapes = {"ns2d": (64, 64, 64, 64, (1, 1)), "sw": (96, 192, 64, 64, (0, 1)),
              "twophase": (61, 121, 64, 64, (0, 0)), "twophase_cond": (61, 121, 64, 64, (0, 0))}
    H, W, Cin, Cout, modes = shapes[workload]
    nb = max(8, (768 << 20) // (H * W * Cin * 2))  # ~768 MB of bf16 input: >> 126 MB L2
    dt16 = torch.float16 if precision == "fp16" else torch.bfloat16
    x = ops.Act(torch.randn(nb * H * W * Cin, device=device).to(dt16), nb, H, W, Cin)
    wt = torch.nn.Parameter(torch.randn(Cout, Cin, 3, 3, device=device) / math.sqrt(9 * Cin))
    bs = torch.nn.Parameter(torch.zeros(Cout, device=device))
    filt = ops.PackedFilter.of(wt, bs)
    out = ops.Act.empty(nb, H, W, Cout, dt16, device)
    with ops.precision(precision):
        for _ in range(3):
            ops.conv2d(x, filt, pad=(1, 1, 1, 1), pad_mode=modes, out=out)
        torch.cuda.synchronize()
        reps = 10
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(reps + 1)]
        ev[0].record()
        for i in range(reps):
            ops.conv2d(x, filt, pad=(1, 1, 1, 1), pad_mode=modes, out=out)
            ev[i + 1].record()
        torch.cuda.synchronize()
    ms = sorted(ev[i].elapsed_time(ev[i + 1]) for i in range(reps))
    t = sum(ms) / len(ms) * 1e-3
    flops = 2.0 * nb * H * W * Cout * 9 * Cin
    achieved = flops / t / 1e12
    traffic = None  # dram read+write bytes per launch from the committed ncu --set full capture of this exact shape
    tp = os.path.join(ROOT, "profiles", "r01_halo_traffic.json")
    if workload == "ns2d" and os.path.exists(tp):
        tj = json.load(open(tp))
        if tj.get("algorithmic_bytes") == nb * H * W * (Cin + Cout) * 2:
            traffic = tj["dram_bytes_read"] + tj["dram_bytes_write"]
    peak = peaks["bf16_tflops"]  # burst figure: this kernel is timed alone
    return {"bound": "tensor", "kernel": f"conv_halo_kernel<{Cout}> (tcgen05, smem halo + resident filter) 3x3 {Cin}->{Cout} "
                                        f"@ {H}x{W}, batch {nb}",
            "achieved": round(achieved, 2), "peak": peak, "unit": "TFLOP/s", "frac": round(achieved / peak, 4),
            "traffic": traffic, "algorithmic_bytes": nb * H * W * (Cin + Cout) * 2, "algorithmic_flops": flops,
            "avg_launch_ms": round(t * 1e3, 4), "peak_source": peaks["source"],
            "hbm_gbs_at_algorithmic_bytes": round(nb * H * W * (Cin + Cout) * 2 / t / 1e9, 1)}


def time_share(pattern):
    """share of the step's kernel time of kernels whose name contains `pattern`, from the committed ncu launch list"""
    tp = os.path.join(ROOT, "profiles", "r01_launches_bench.txt")
    if not os.path.exists(tp):
        return None
    tot = 0.0
    for ln in open(tp):
        parts = ln.split()
        if len(parts) > 3 and parts[1] == "ms" and parts[2].endswith("%") and pattern in ln:
            tot += float(parts[2].rstrip("%"))
    return round(tot / 100.0, 4) if tot else None


def kernels_by_time(torch, ops, device, peaks, roof):
    """The three kernels with the largest share of the NS2d step's time, each timed alone (CUDA events, operands > L2 or
    one CTA wave per SM) against the measured bf16 tensor peak.  Shares come from profiles/r01_launches_bench.txt."""
    import math
    out = []

    def timed(fn, reps=5):
        for _ in range(2):
            fn()
        torch.cuda.synchronize()
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(reps + 1)]
        ev[0].record()
        for i in range(reps):
            fn()
            ev[i + 1].record()
        torch.cuda.synchronize()
        return sum(ev[i].elapsed_time(ev[i + 1]) for i in range(reps)) / reps * 1e-3

    with torch.no_grad(), ops.precision("bf16"):
        # FABlock2D whole-block kernel, 32x32, 4736 samples (one decode chunk = 32 per SM)
        nb, H, W = 4736, 32, 32
        u = ops.Act(torch.randn(nb * H * W * 64, device=device).bfloat16(), nb, H, W, 64)
        sc, sh = torch.rand(nb * 64, device=device) + 0.5, torch.randn(nb * 64, device=device) * 0.1
        w = torch.nn.Parameter(torch.randn(512, 64, device=device) / 8)
        w1 = torch.nn.Parameter(torch.randn(64, 512, 1, 1, device=device) / 22)
        w2 = torch.nn.Parameter(torch.randn(64, 64, 1, 1, device=device) / 8)
        kx = torch.randn(nb, 8, H, H, device=device) / H ** 0.5
        ky = torch.randn(nb, 8, W, W, device=device) / W ** 0.5
        t = timed(lambda: ops.fablock_full(u, sc, sh, w, kx, ky, 8, 1e-5, w1, w2))
        fl = nb * (2.0 * H * W * 64 * 512 * 2 + 2.0 * 8 * (H * H * W + H * W * W) * 64 + 2.0 * H * W * 64 * 64)
        by = nb * (2 * H * W * 64 * 2 + 8 * (H * H + W * W) * 4)
        out.append({"kernel": "fablock_full_kernel<512> (FABlock2D per sample, mma.sync phases + tcgen05 to_out, TMEM-resident "
                              "accumulator) 32x32, batch 4736", "share_of_step": time_share("fablock_full_kernel"),
                    "bound": "tensor", "achieved": round(fl / t / 1e12, 2), "peak": peaks["bf16_tflops"], "unit": "TFLOP/s",
                    "frac": round(fl / t / 1e12 / peaks["bf16_tflops"], 4), "avg_launch_ms": round(t * 1e3, 4),
                    "algorithmic_bytes": by, "hbm_gbs_at_algorithmic_bytes": round(by / t / 1e9, 1)})
        del u, kx, ky
        # propagator conv: 3x3 128->128 circular @ 8x8, one propagator step of the bench batch (latent-grid engine)
        nb, H, W, C = 1184, 8, 8, 128
        x = ops.Act(torch.randn(nb * H * W * C, device=device).bfloat16(), nb, H, W, C)
        wt = torch.nn.Parameter(torch.randn(C, C, 3, 3, device=device) / math.sqrt(9 * C))
        bs = torch.nn.Parameter(torch.zeros(C, device=device))
        filt = ops.PackedFilter.of(wt, bs)
        y = ops.Act.empty(nb, H, W, C, torch.bfloat16, device)
        t = timed(lambda: ops.conv2d(x, filt, pad=(1, 1, 1, 1), pad_mode=(1, 1), act=ops.ACT_GELU, out=y), reps=20)
        fl = 2.0 * nb * H * W * C * 9 * C
        out.append({"kernel": "conv_latent_kernel (tcgen05 latent-grid engine: resident halos of 4 samples, streamed filter) 3x3 "
                              "128->128 circular @ 8x8, batch 1184 (L2-resident: 19 MB in + 19 MB out)",
                    "share_of_step": time_share("conv_latent_kernel"),
                    "bound": "tensor", "achieved": round(fl / t / 1e12, 2), "peak": peaks["bf16_tflops"], "unit": "TFLOP/s",
                    "frac": round(fl / t / 1e12 / peaks["bf16_tflops"], 4), "avg_launch_ms": round(t * 1e3, 4)})
    out.append({"kernel": roof["kernel"], "share_of_step": time_share("conv_halo_kernel"), "bound": "tensor",
                "achieved": roof["achieved"], "peak": roof["peak"], "unit": "TFLOP/s", "frac": roof["frac"],
                "avg_launch_ms": roof["avg_launch_ms"]})
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
        if rank != 0:
            return 0
        # bounded sample per step; steps/warmup only scale the repeat count (the CPU path is batch-flat)
        base = cpu_reference_throughput(args.workload, budget_s=min(60.0, 6.0 * max(1, args.steps)))
        line = {"metric": "latent rollout trajectory-steps/sec", "impl": "reference", "value": round(base["value"], 3),
                "unit": "trajectory-steps/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": None, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
                "data": "synthetic", "config": {"workload": label, "rollout_steps": R},
                "cpu_baseline": base,
                "e2e": {"value": round(base["value"], 3), "unit": "trajectory-steps/s", "h2d_bytes_per_step": 0,
                        "d2h_bytes_per_step": 0}, "gpu_launches": 0}
        print(json.dumps(line), flush=True)
        return 0

    import torch
    import torch.distributed as dist
    from lns_b200 import ops
    from lns_b200.configs import get_config
    from lns_b200.dist import OverlappedGather, init_from_env
    from lns_b200.latent_dynamics import LatentDynamics
    from lns_b200.rollout import Rollout
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import lns_oracle as O  # only for seeded synthetic inputs / zero-init redraw and the cpu_baseline leg

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
    og = OverlappedGather((B, R, ro.C, ro.Ly, ro.Lx), torch.float32, device) if gather else None

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

    for _ in range(max(args.warmup, 3)):
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

    # end to end through the public API: pinned host input -> H2D -> rollout -> D2H of the predicted fields, EVERY step.
    # The D2H copy of step i (335 MB, ~6 ms of PCIe time) runs on a copy stream out of a device staging buffer while the
    # rollout of step i+1 computes (double-buffered, event-ordered); the timed region ends when the last copy has landed.
    out_host = [torch.empty((B, R, ro.C, ro.Ly, ro.Lx), dtype=torch.float32).pin_memory() for _ in range(2)]
    stage = [torch.empty((B, R, ro.C, ro.Ly, ro.Lx), dtype=torch.float32, device=device) for _ in range(2)]
    copy_stream = torch.cuda.Stream(device)
    compute = torch.cuda.current_stream(device)
    staged = [torch.cuda.Event() for _ in range(2)]
    landed = [torch.cuda.Event() for _ in range(2)]
    for ev in landed:
        ev.record(compute)
    e2e_steps = max(2, min(args.steps, 5))

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

    if rank == 0:
        roof = conv_roofline(torch, ops, device, peaks, args.workload, args.precision) if args.precision in ("bf16", "fp16") else None
        tflops = value * gflop_per_ts / 1e3
        line = {
            "metric": "latent rollout trajectory-steps/sec", "value": round(value, 1), "unit": "trajectory-steps/s",
            "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": round(elapsed_ms / args.steps, 4), "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": {"bf16": "bf16", "fp16": "f16", "tf32": "tf32", "fp32": "f32"}[args.precision], "data": "synthetic",
            "config": {"workload": label, "rollout_steps": R, "trajectories_per_gpu": B, "global_batch": B * world,
                       "parallelism": f"trajectory-sharded x{world}" + (", final all-gather of fields (step i overlaps the rollout of step i+1)" if gather else ""),
                       "l2": "per-step working set (activations) is far larger than the 126 MB L2; no explicit flush",
                       "cuda_graph": True, "random_init_weights_seed": 1234},
            "whole_path_tflops_per_gpu": round(tflops / world, 2),
            "whole_path_frac_of_bf16_sustained": round(tflops / world / peaks["bf16_tflops_sustained"], 4),
            "e2e": {"value": round(e2e_value, 1), "unit": "trajectory-steps/s", "h2d_bytes_per_step": h2d,
                    "d2h_bytes_per_step": d2h, "steps": e2e_steps},
            "gpu_launches": int(ro.launches_per_call * args.steps),
            "launches_per_step": int(ro.launches_per_call),
            "clocks": clocks, "roofline": roof,
        }
        if roof is not None and args.workload == "ns2d" and args.precision == "bf16":
            line["roofline_by_time"] = kernels_by_time(torch, ops, device, peaks, roof)
        if not args.no_cpu_baseline:
            line["cpu_baseline"] = cpu_reference_throughput(args.workload, budget_s=15.0)
            line["parity"] = parity_check(model, cfg, args.precision, device)
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
