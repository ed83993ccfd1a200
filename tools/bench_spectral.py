"""Stand-alone benchmark of the spectral path (VERDICT r1 item 7 / BASELINE config 4's wording): SpectralConv2d and the whole
CondFourierBasicBlock (modules/fourier_cond.py:32-118) at [128,64,61,121] modes [16,31] and [128,64,64,64] modes [16,16].
Reports ms per call and ALGORITHMIC GB/s (one read of x + one write of the result, as stored) against the measured HBM peak.
LNS_SPECTRAL_SCALAR=1 selects the CUDA-core reference path (run the script twice to compare)."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from lns_b200 import ops  # noqa: E402
from modules.fourier_cond import CondFourierBasicBlock  # noqa: E402


def timed(fn, iters=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def run(peak=None, precisions=("fp32", "fp16s")):
    if peak is None:
        peak = 6548.5
        try:
            peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
        except Exception:
            pass
    path = "scalar" if os.environ.get("LNS_SPECTRAL_SCALAR") else "tensor-core (TF32x3 mma.sync)"
    out = {"path": path, "hbm_peak_gbs": peak, "cases": []}
    for (B, C, H, W, modes) in ((128, 64, 61, 121, [16, 31]), (128, 64, 64, 64, [16, 16])):
        torch.manual_seed(0)
        blk = CondFourierBasicBlock(C, C, modes=modes).cuda().eval()
        x = torch.randn(B, C, H, W, device="cuda")
        emb = torch.randn(B, C, device="cuda")
        for prec in precisions:
            with torch.no_grad(), ops.precision(prec):
                xa = ops.nchw_to_act(x)
                er = ops.rows_act(emb)
                ems = blk.fourier.cond_emb._fwd(er)
                wm = blk.fourier._mode_weights()
                t_spec = timed(lambda: ops.spectral_conv2d(xa, wm, modes[0], modes[1], C, emb=ems))
                t_blk = timed(lambda: blk._fwd(xa, er))
            esz = xa.t.element_size()
            by_spec = B * H * W * C * (esz + 4)          # x read as stored + fp32 result written
            by_blk = B * H * W * C * (2 * esz + esz)     # block: x read by the spectral conv and by the 1x1 conv, result as stored
            flops = 2.0 * B * (H * W * 2 * modes[1] * C * 2 + modes[1] * (4 * modes[0] * 2 * H * C * 2 + 2 * modes[0] * 4 * C * C))
            out["cases"].append({
                "shape": [B, C, H, W], "modes": modes, "precision": prec,
                "spectral_conv_ms": round(t_spec, 4), "spectral_conv_gbs": round(by_spec / t_spec / 1e6, 1),
                "spectral_conv_frac_of_hbm_peak": round(by_spec / t_spec / 1e6 / peak, 4),
                "spectral_conv_tflops": round(flops / t_spec / 1e9, 2),
                "block_ms": round(t_blk, 4), "block_gbs": round(by_blk / t_blk / 1e6, 1),
                "block_frac_of_hbm_peak": round(by_blk / t_blk / 1e6 / peak, 4)})
    return out


if __name__ == "__main__":
    print(json.dumps(run()))
