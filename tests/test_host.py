"""CPU tests of the host side: the C-ABI library loads and exports every symbol include/lns_b200.h declares, argument
validation fails loudly, configs / state_dict layouts match the reference's, no CPU fallback exists."""
import ctypes
import os
import re

import pytest
import torch

from lns_b200 import REPO_ROOT, _C, ops
from lns_b200.configs import CONFIGS, get_config
from lns_b200.latent_dynamics import LatentDynamics


def declared_symbols():
    text = open(os.path.join(REPO_ROOT, "include", "lns_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(lns_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    lib = _C.lib()
    names = declared_symbols()
    assert len(names) >= 20
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/lns_b200.h but not exported"
    assert set(names) == set(_C.SIGNATURES), set(names) ^ set(_C.SIGNATURES)
    assert b"sm_100a" in lib.lns_version()


def test_convdesc_layout_matches_header():
    """ctypes mirror of LnsConvDesc: same field order as the C struct (guards against silent ABI drift)."""
    text = open(os.path.join(REPO_ROOT, "include", "lns_b200.h")).read()
    body = text[text.index("typedef struct LnsConvDesc {") + len("typedef struct LnsConvDesc {"):text.index("} LnsConvDesc;")]
    body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
    fields = []
    for decl in body.split(";"):
        decl = decl.strip()
        if not decl:
            continue
        names = re.sub(r"^(const\s+)?(void|float|int32_t|int64_t)\s*\*?", "", decl)
        fields += [n.strip().lstrip("*").strip() for n in names.split(",")]
    assert fields == [f[0] for f in _C.ConvDesc._fields_]


def test_pure_host_entry_points():
    lib = _C.lib()
    assert lib.lns_chan_stats_chunks(64, 64) == 4 and lib.lns_chan_stats_chunks(8, 8) == 1
    assert lib.lns_packed_weight_bytes(128, 128, 3, 3, ops.W_UMMA_BF16) == 128 * 128 * 9 * 2
    assert lib.lns_packed_weight_bytes(128, 16, 1, 1, ops.W_UMMA_BF16) == -1  # Cin % 64 != 0 -> not packable
    assert lib.lns_packed_weight_bytes(1, 64, 1, 1, ops.W_SIMT_F32) == 256
    # X1 | Z planar rows (B H 2 m2 (Ci + Co)) + X2 | Y mode-major (2 m1 m2 B 2 (Ci + Co)) fp32
    assert lib.lns_spectral_work_bytes(2, 61, 121, 64, 64, 16, 31) == (2 * 61 * 2 * 31 * 128 + 2 * 16 * 31 * 2 * 2 * 128) * 4
    assert lib.lns_conv_stats_chunks(8, 8) == 4 and lib.lns_conv_stats_chunks(16, 16) == 16 and lib.lns_conv_stats_chunks(7, 15) == 8


def test_invalid_arguments_are_errors_not_fallbacks():
    lib = _C.lib()
    d = _C.ConvDesc()  # all-null descriptor
    assert lib.lns_conv2d(ctypes.byref(d), None) == -1
    assert b"null" in lib.lns_last_error()
    assert lib.lns_chan_stats(None, 0, 1, 1, 1, 64, 64, None, None) == -1
    assert lib.lns_spectral_conv2d(ctypes.c_void_p(16), 0, 1, 8, 8, 4, 4, 5, 3, ctypes.c_void_p(16), None,
                                   ctypes.c_void_p(16), ctypes.c_void_p(16), None) == -1  # 2*m1 > H
    assert b"modes1" in lib.lns_last_error()


def test_cpu_tensors_raise():
    with pytest.raises(ops.LnsError):
        ops.nchw_to_act(torch.zeros(1, 1, 8, 8))
    with pytest.raises(ops.LnsError):
        ops.Act.from_nchw(torch.zeros(1, 1, 8, 8))
    from modules.basics import ResidualBlock
    blk = ResidualBlock(64, 64, num_dimensions=2, padding_mode="circular")
    with pytest.raises(ops.LnsError):
        blk(torch.zeros(1, 64, 8, 8))


def test_missing_library_fails_loudly(monkeypatch, tmp_path):
    monkeypatch.setattr(_C, "_lib", None)
    monkeypatch.setattr(_C, "LIB", str(tmp_path / "nope.so"))
    with pytest.raises(_C.LnsError, match="no CPU or PyTorch-eager fallback"):
        _C.lib()


@pytest.mark.parametrize("name", sorted(CONFIGS))
def test_state_dict_layout(name):
    """Key names / shapes the reference's checkpoints have and strict round trip."""
    model = LatentDynamics(get_config(name))
    sd = model.state_dict()
    ae = "ae" if name == "twophase_cond" else "vq_ae"
    n_expected = {"ns2d": 213, "sw": 228, "twophase": 187, "twophase_cond": 231}[name]
    assert len(sd) == n_expected
    assert f"{ae}.quant_conv.weight" in sd and f"{ae}.post_quant_conv.bias" in sd
    assert sd["propagator.in_proj.weight"].shape == (128, model.latent_dim, 1, 1)
    if name == "ns2d":
        assert sd["vq_ae.decoder.model.2.pe"].shape == (1, 64, 128)
        assert sd["vq_ae.decoder.model.2.to_q.weight"].shape == (512, 128)
        assert sd["vq_ae.decoder.model.8.in_proj.weight"].shape == (512, 64, 1, 1)
        assert sd["vq_ae.decoder.model.11.low_rank_kernel_x.to_qk.weight"].shape == (2048, 64)
        assert sd["vq_ae.decoder.model.8.low_rank_kernel_y.pos_emb.inv_freq"].shape == (64,)
        assert "propagator.net.2.ffn.3.weight" in sd and "propagator.net.0.ffn.1.bias" not in sd
    if name == "twophase":
        assert not any("low_rank_kernel" in k for k in sd)  # FABlock2D is never instantiated for two-phase
        assert sd["vq_ae.decoder.model.2.pe"].shape == (1, 119, 128)
    if name == "twophase_cond":
        assert sd["propagator.net.3.cond_conv1.2.weight"].abs().max() == 0  # zero_module
        assert sd["propagator.cond_emb_proj.2.weight"].shape == (64, 64)
    clone = LatentDynamics(get_config(name))
    assert not clone.load_state_dict(sd, strict=True).missing_keys


def test_reference_defects_are_mirrored():
    import argparse
    from modules.propagator import SimpleMLP, SimpleResNet, ConditionalResNet
    args = argparse.Namespace(latent_dim=4, latent_resolution=4, propagator_dim=32, is_periodic=True)
    assert len(SimpleMLP(args).state_dict()) == 6
    with pytest.raises(TypeError):
        SimpleResNet(args)  # ResidualBlock without num_dimensions, modules/propagator.py:22-24
    with pytest.raises(TypeError):
        ConditionalResNet(args)


def test_rollout_chunking_rules():
    """Decode chunk = about 16 M output pixels per channel, split evenly, rounded up to a multiple of the SM count; pipelined
    decode groups = whole rollout steps, about one chunk each, at least four groups per rollout (host logic, no GPU)."""
    import types

    import torch
    from lns_b200.rollout import Rollout
    stub = types.SimpleNamespace(Ly=64, Lx=64, device=torch.device("cpu"), decode_chunk=None, B=1184, K=20)
    stub._default_chunk = lambda n: Rollout._default_chunk(stub, n)
    assert Rollout._default_chunk(stub, 1184 * 20) == 4736          # 5 equal chunks of 32 x 148 samples
    assert Rollout._default_chunk(stub, 1024 * 20) % 148 == 0
    assert Rollout._default_chunk(stub, 8) == 8                     # never more than the samples there are
    assert Rollout._steps_per_group(stub, 1184 * 20) == 4           # 4 steps x 1184 trajectories = one chunk
    stub.B, stub.K = 4, 2
    assert Rollout._steps_per_group(stub, 8) == 1
    stub.B, stub.K, stub.Ly, stub.Lx = 64, 20, 96, 192              # shallow water, BASELINE config 2
    s = Rollout._steps_per_group(stub, 64 * 20)
    assert 1 <= s <= 5 and -(-20 // s) >= 4
    stub.decode_chunk = 640
    assert Rollout._steps_per_group(stub, 64 * 20) == 5


def test_training_rollout_host_logic():
    """lns_b200.train on the CPU: which propagators it recognises, that the conditional flag and `param` must agree, that CPU
    tensors fail loudly (no fallback), and that the null-argument checks of the backward entry points answer LNS_E_INVALID."""
    import torch
    import torch.nn.functional as F
    from lns_b200 import _C, ops, train
    from lns_b200.configs import get_config
    from lns_b200.latent_dynamics import LatentDynamics
    plain = LatentDynamics(get_config("ns2d")).propagator
    cond = LatentDynamics(get_config("twophase_cond")).propagator
    assert train._check_net(plain) is False and train._check_net(cond) is True
    with pytest.raises(NotImplementedError):
        train._check_net(torch.nn.Linear(4, 4))
    z = torch.zeros(2, 16, 8, 8)
    with pytest.raises(ops.LnsError):
        train.rollout_train(plain, z, 2, param=torch.zeros(2))      # param with the unconditional propagator
    with pytest.raises(ops.LnsError):
        train.rollout_train(cond, torch.zeros(2, 64, 7, 15), 2)     # the conditional one without it
    with pytest.raises(ops.LnsError):
        train.rollout_train(plain, z, 2)                            # CPU tensor: there is no CPU path
    model = LatentDynamics(get_config("ns2d"))
    with pytest.raises(ValueError):
        model(torch.zeros(2, 2, 16, 8, 8), torch.zeros(2, 2, 16, 8, 8), F.mse_loss)   # more than one input frame
    lib = _C.lib()
    assert lib.lns_conv2d_wgrad_work_bytes(4, 8, 8, 128, 128, 3, 3) == 64 * 9 * 128 * 128 * 4
    assert lib.lns_conv2d_wgrad_work_bytes(0, 8, 8, 128, 128, 3, 3) == -1
    assert lib.lns_chan_sum_slices(7) == 7 and lib.lns_chan_sum_slices(4096) == 128
    assert lib.lns_conv2d_wgrad(None, 0, None, None, 0, None, 0, 1, 8, 8, 4, 4, 3, 3, 1, 1, 1, 0, 0, 0, 1.0, None, None, None, None) == -1
    assert lib.lns_group_norm_bwd(None, 0, None, 0, None, 0, 1, 64, 128, 1, 1e-5, None, None, 0, None, None, None) == -1
    assert lib.lns_act_bwd(None, None, 4, 2, None, None) == -1
    assert lib.lns_pixel_dot(None, 0, None, 0, 1, 1, 128, 1, None, None) == -1
    assert lib.lns_scale_add(None, None, None, 1, 1, 4, None, None) == -1
    assert lib.lns_absmax(None, 4, None, None) == -1 and lib.lns_loss_scale(None, 64.0, None, None) == -1
    assert lib.lns_norm_finalize_centred(None, 1, 1, 4, 1, 1, 1e-5, None, None, None, None, None, None) == -1
    assert b"lns_" in lib.lns_last_error()


def test_fablock_staged_host_logic():
    """Host side of the pre-staged FABlock2D path (modules/factorized_attention.py:144-159): operand preparation, the shapes the
    staged entry points accept, and that nothing falls back on bad arguments."""
    torch.manual_seed(5)
    heads = 8
    w_in = torch.randn(heads * 64, 64, 1, 1)
    w_out1 = torch.randn(64, heads * 64, 1, 1)
    w_in16, w1h = ops.fablock_staged_operands(w_in, w_out1, heads, torch.float16)
    assert w_in16.shape == (heads, 64, 72) and w_in16.dtype == torch.float16 and w_in16.is_contiguous()
    assert torch.equal(w_in16[:, :, :64], w_in.reshape(heads, 64, 64).half()) and float(w_in16[:, :, 64:].abs().max()) == 0.0
    assert w1h.shape == (heads, 64, 64) and w1h.dtype == torch.float32 and w1h.is_contiguous()
    # w1h[h][o][c] = to_out[1].weight[o][h * 64 + c]
    assert torch.equal(w1h[3, 5], w_out1.reshape(64, heads * 64)[5, 3 * 64:4 * 64])
    lib = _C.lib()
    for H, W, ok in ((16, 16, 1), (32, 32, 1), (16, 32, 1), (32, 16, 1), (24, 48, 0), (8, 8, 0), (31, 16, 0), (64, 64, 0)):
        assert lib.lns_fablock_full_staged_supported(H, W, 64, 64, 64) == ok
    assert lib.lns_fablock_full_staged_supported(32, 32, 128, 64, 64) == 0 and lib.lns_fablock_full_staged_supported(32, 32, 64, 32, 64) == 0
    # null pointers / unsupported shapes answer with an error code, never a fallback
    assert lib.lns_fablock_full_staged(None, None, 2, 1, 32, 32, 8, None, None, None, 1e-5, None, None, None, None) != 0
    assert lib.lns_fablock_prepass_staged(None, 2, 1, 32, 32, 64, 32 * 32 * 64, 1e-5, None, None, None, None, None, None, None, None) != 0
    assert b"lns_fablock" in lib.lns_last_error()


@pytest.mark.parametrize("H,W", [(16, 16), (32, 32), (16, 32), (32, 16)])
def test_fablock_staged_layout_properties(H, W):
    """The staged FABlock2D image (csrc/fablock.cu fablock_prepass3_kernel -> csrc/fablock_full.cu): row s holds pixel
    s ^ ((s >> log2 W) & 7), 16-byte chunk ch sits at ch ^ (s & 7).  Properties the whole-block kernel relies on: the row map is a
    bijection that keeps the image row, a COLUMN walk (fixed x, 8 consecutive y) and a ROW walk (8 consecutive x) both touch all 8
    swizzle phases (conflict-free ldmatrix in both contraction directions), and a 128-row MMA tile is 128 consecutive rows."""
    lg = W.bit_length() - 1
    s = torch.arange(H * W)
    src = s ^ ((s >> lg) & 7)
    assert sorted(src.tolist()) == list(range(H * W))            # bijection
    assert torch.equal(src >> lg, s >> lg)                       # x is permuted inside its image row only
    row_of = torch.empty_like(s)
    row_of[src] = s                                              # pixel index -> staged row
    for x in range(0, W, 5):
        for y0 in range(0, H - 7):
            phases = {int(row_of[(y0 + k) * W + x]) & 7 for k in range(8)}
            assert len(phases) == 8
    for y in range(H):
        for x0 in range(0, W, 8):
            phases = {int(row_of[y * W + x0 + k]) & 7 for k in range(8)}
            assert len(phases) == 8
    # chunk swizzle: an involution per row, the 8 chunks of a row stay inside its 128 bytes
    for srow in (0, 5, H * W - 1):
        pos = [(ch ^ (srow & 7)) for ch in range(8)]
        assert sorted(pos) == list(range(8)) and [p ^ (srow & 7) for p in pos] == list(range(8))
