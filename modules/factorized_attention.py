"""Axial factorized attention -- drop-in for the reference's ``modules/factorized_attention.py``
(``LowRankKernel`` :11-69, ``PoolingReducer`` :72-94, ``FABlock2D`` :97-159)."""
import torch
import torch.nn as nn

from lns_b200 import ops

from ._base import LnsModule, LnsError, conv_layer, filt_of, norm_affine, cache_of
from .embedding import RotaryEmbedding


class LowRankKernel(LnsModule):
    """K = rot(q) rot(k)^T per head along one axis, no softmax, no 1/sqrt(d) (reference :43-69)."""

    def __init__(self, dim, dim_head, heads, use_rotary_emb=False, dropout=0, scaling=1, qk_norm=False):
        super().__init__()
        self.layers = nn.ModuleList([])
        self.dim_head, self.heads = dim_head, heads
        self.dropout = nn.Dropout(dropout) if dropout > 1e-6 else nn.Identity()
        self.to_qk = nn.Linear(dim, dim_head * heads * 2, bias=False)
        self.qk_norm = qk_norm
        if qk_norm:
            self.q_norm = nn.LayerNorm(dim_head, elementwise_affine=False)
            self.k_norm = nn.LayerNorm(dim_head, elementwise_affine=False)
        self.use_rotary_emb = use_rotary_emb
        if use_rotary_emb:
            self.pos_emb = RotaryEmbedding(dim_head)
        self.scaling = scaling

    def _tables(self, n, device):
        """cos / sin [n, dim_head/2] for pos = linspace(0,1,n) (reference :47), cached per n."""
        cache = cache_of(self).setdefault("rot", {})
        key = (n, str(device), self.pos_emb.inv_freq._version if self.use_rotary_emb else 0)
        if key not in cache:
            if self.use_rotary_emb:
                pos = torch.linspace(0, 1, n, device=device)  # fp32 like the reference
                ang = self.pos_emb.angles(pos)
                cache[key] = (ang.cos().float().contiguous(), ang.sin().float().contiguous())
            else:
                z = torch.zeros(n, self.dim_head // 2, device=device)
                cache[key] = (torch.ones_like(z), z)
        return cache[key]

    def _fwd(self, u):
        """u: Act [B, n, 1, dim] -> torch fp32 [B, heads, n, n]"""
        if self.qk_norm:
            raise LnsError("LowRankKernel: qk_norm=True is not on the rollout path")
        n = u.H * u.W
        # [B, n, 2*heads*d]: fp32 on the validation path, bf16 (half the bytes; it is read once) on the bf16 path
        qk = ops.conv2d(u, filt_of(self.to_qk), out_dtype=torch.float32 if u.t.dtype == torch.float32 else ops.act_dtype())
        cos_t, sin_t = self._tables(n, u.t.device)
        return ops.lowrank_kernel(qk, self.heads, self.dim_head, cos_t, sin_t, self.scaling)


class PoolingReducer(LnsModule):
    """Linear -> mean over the other axis -> LN -> Linear -> GELU -> Linear (reference :72-94).  The bias-free input
    linear commutes with the mean, so the mean is taken first (on the caller's side) and this block only sees
    [B, n, 1, C] rows."""

    def __init__(self, in_dim, hidden_dim, out_dim):
        super().__init__()
        self.to_in = nn.Linear(in_dim, hidden_dim, bias=False)
        self.out_ffn = nn.Sequential(
            nn.LayerNorm(hidden_dim),
            nn.Linear(hidden_dim, hidden_dim * 2, bias=False),
            nn.GELU(),
            nn.Linear(hidden_dim * 2, out_dim, bias=True))

    def _fwd(self, pooled):
        f32 = pooled.t.dtype  # fp32 on the validation path, bf16 on the fast path (decided by FABlock2D)
        h = ops.conv2d(pooled, filt_of(self.to_in), out_dtype=f32)
        ln = self.out_ffn[0]
        h = ops.layernorm(h, ln.weight, ln.bias, ln.eps)
        h = ops.conv2d(h, filt_of(self.out_ffn[1]), act=ops.ACT_GELU, out_dtype=f32)
        # in the bf16 mode the result is stored as bf16 so that the 64 -> 2048 to_qk GEMM (the only sizeable GEMM of the
        # pooled branch) runs on the tensor-core engine
        return ops.conv2d(h, filt_of(self.out_ffn[3]), out_dtype=ops.act_dtype())


class _SwapAxes(nn.Module):
    """Parameter-free placeholder for einops' Rearrange('b c nx ny -> b c ny nx') at index 0 of ``to_y`` -- keeps the
    reducer at ``to_y.1`` so reference state_dict keys match."""

    def forward(self, x):
        return x.transpose(-1, -2)


class FABlock2D(LnsModule):
    """Factorized attention block (reference :97-159):
        u = GN1(u); u_phi = in_proj(u) [heads*dim_head ch]; t = to_in(u)
        u_x = reducer_x(mean_W t), u_y = reducer_y(mean_H t); K_x = kernel(u_x) [H,H], K_y = kernel(u_y) [W,W]
        u_phi <- K_x (over H) then K_y (over W); InstanceNorm -> 1x1 -> GELU -> 1x1; + skip
    """

    def __init__(self, dim, dim_head, latent_dim, heads, dim_out, use_rope=True, kernel_multiplier=2, qk_norm=False):
        super().__init__()
        self.dim, self.latent_dim, self.heads = dim, latent_dim, heads
        self.dim_head = dim_head
        self.in_norm = nn.GroupNorm(1, dim)
        self.in_proj = nn.Conv2d(dim, heads * dim_head, 1, 1, 0, bias=False)
        self.to_in = nn.Sequential(nn.Conv2d(dim, dim, 1, 1, 0, bias=False))
        self.to_x = nn.Sequential(PoolingReducer(dim, dim, latent_dim))
        self.to_y = nn.Sequential(_SwapAxes(), PoolingReducer(dim, dim, latent_dim))
        self.low_rank_kernel_x = LowRankKernel(latent_dim, dim_head * kernel_multiplier, heads,
                                               use_rotary_emb=use_rope, qk_norm=qk_norm)
        self.low_rank_kernel_y = LowRankKernel(latent_dim, dim_head * kernel_multiplier, heads,
                                               use_rotary_emb=use_rope, qk_norm=qk_norm)
        self.to_out = nn.Sequential(
            nn.InstanceNorm2d(dim_head * heads),
            nn.Conv2d(dim_head * heads, dim_out, 1, 1, 0, bias=False),
            nn.GELU(),
            nn.Conv2d(dim_out, dim_out, 1, 1, 0, bias=False))

    def _fwd(self, u):
        u = ops.as_h16(u)
        skip = u
        f32 = torch.float32
        inorm = self.to_out[0]
        fused = (ops.fast16() and self.in_proj.out_channels == self.heads * 64
                 and ops.fablock_core_supported(u, self.dim_head) and inorm.weight is None)
        staged = None
        if fused:
            # one read of u: GroupNorm(1) affine + both pooled tensors (means commute with the per-channel affine); when the
            # producer-warp whole-block kernel takes the shape, also the normalised sample as that kernel's shared-memory image
            if (self.to_out[1].bias is None and self.to_out[3].bias is None and self.to_out[1].out_channels == 64
                    and ops.fablock_full_staged_supported(u, self.dim_head, self.to_out[3].out_channels)
                    and not ops.fablock_tc_supported(u, self.dim_head, self.to_out[3].out_channels)):
                s, t, mx, my, staged = ops.fablock_prepass(u, self.in_norm.eps, self.in_norm.weight, self.in_norm.bias, staged=True)
            else:
                s, t, mx, my = ops.fablock_prepass(u, self.in_norm.eps, self.in_norm.weight, self.in_norm.bias)
        else:
            s, t = norm_affine(u, self.in_norm)
            un = ops.affine_act(u, s, t, ops.ACT_NONE)
            mx, my = ops.axis_mean(un, axis=1), ops.axis_mean(un, axis=0)
        # mean over the other axis first (exact: to_in convs have no bias), then the two tiny linears.  On the bf16 path the
        # pooled branch is stored in bf16 from here on, so that its 64/128-wide linears run on the tensor-core engine
        # (on the CUDA-core engine they were 8 % of the rollout).
        if fused:
            kk = self._axis_kernels(mx, my)
            if kk is not None:
                return self._finish(u, skip, s, t, kk[0], kk[1], inorm, fused, staged=staged)
        pd = ops.act_dtype() if fused else f32
        px = ops.conv2d(mx, filt_of(self.to_in[0]), out_dtype=pd)   # rows indexed by H
        py = ops.conv2d(my, filt_of(self.to_in[0]), out_dtype=pd)   # rows indexed by W
        k_x = self.low_rank_kernel_x._fwd(self.to_x[0]._fwd(px))
        k_y = self.low_rank_kernel_y._fwd(self.to_y[1]._fwd(py))
        return self._finish(u, skip, s, t, k_x, k_y, inorm, fused, un=None if fused else un, staged=staged)

    def _axis_operands(self, reducer, lrk, dt16):
        """Host-prepared operands of lns_fa_axis_kernel for one axis, cached until a source parameter changes."""
        srcs = [self.to_in[0].weight, reducer.to_in.weight, reducer.out_ffn[0].weight, reducer.out_ffn[0].bias,
                reducer.out_ffn[1].weight, reducer.out_ffn[3].weight, reducer.out_ffn[3].bias, lrk.to_qk.weight]
        key = (dt16,) + tuple((t.data_ptr(), t._version, str(t.device)) for t in srcs)
        cache = cache_of(self).setdefault("axis", {})
        ent = cache.get(id(reducer))
        if ent is None or ent[0] != key:
            with torch.no_grad():
                w_in = self.to_in[0].weight.detach().double().reshape(self.dim, self.dim)
                w1 = reducer.to_in.weight.detach().double() @ w_in          # both bias-free: one 64x64 linear
                ops_ = dict(
                    w1t=w1.t().contiguous().float(),
                    ln_g=reducer.out_ffn[0].weight.detach().float().contiguous(),
                    ln_b=reducer.out_ffn[0].bias.detach().float().contiguous(),
                    wf1t=reducer.out_ffn[1].weight.detach().float().t().contiguous(),
                    wf2t=reducer.out_ffn[3].weight.detach().float().t().contiguous(),
                    bf2=reducer.out_ffn[3].bias.detach().float().contiguous(),
                    wqk16=lrk.to_qk.weight.detach().to(dt16).contiguous())
            ent = (key, ops_)
            cache[id(reducer)] = ent
        return ent[1]

    def _axis_kernels(self, mx, my):
        """(K_x, K_y) through the one-kernel-per-axis pooled branch, or None when the block's shape is not covered."""
        rx, ry = self.to_x[0], self.to_y[1]
        lx, ly = self.low_rank_kernel_x, self.low_rank_kernel_y
        ok = (not lx.qk_norm and not ly.qk_norm
              and rx.out_ffn[1].bias is None and rx.out_ffn[3].bias is not None
              and ry.out_ffn[1].bias is None and ry.out_ffn[3].bias is not None
              and rx.out_ffn[1].out_features == 128 and ry.out_ffn[1].out_features == 128
              and rx.to_in.out_features == 64 and ry.to_in.out_features == 64
              and ops.fa_axis_kernel_supported(max(mx.H, my.H), self.dim, 64, self.latent_dim, self.heads, lx.dim_head)
              and lx.dim_head == ly.dim_head and lx.heads == self.heads and ly.heads == self.heads)
        if not ok:
            return None
        dt16 = ops.act_dtype()
        out = []
        for pooled, red, lrk in ((mx, rx, lx), (my, ry, ly)):
            o = self._axis_operands(red, lrk, dt16)
            cos_t, sin_t = lrk._tables(pooled.H * pooled.W, pooled.t.device)
            out.append(ops.fa_axis_kernel(pooled, self.heads, o["w1t"], o["ln_g"], o["ln_b"], red.out_ffn[0].eps, o["wf1t"],
                                          o["wf2t"], o["bf2"], o["wqk16"], cos_t, sin_t, lrk.scaling))
        return out

    def _staged_operands(self, dt16):
        """(w_in16, w1h) of ops.fablock_full_staged, cached until in_proj / to_out[1] change."""
        srcs = [self.in_proj.weight, self.to_out[1].weight]
        key = (dt16,) + tuple((w.data_ptr(), w._version, str(w.device)) for w in srcs)
        cache = cache_of(self)
        ent = cache.get("staged")
        if ent is None or ent[0] != key:
            ent = (key, ops.fablock_staged_operands(self.in_proj.weight, self.to_out[1].weight, self.heads, dt16))
            cache["staged"] = ent
        return ent[1]

    def _finish(self, u, skip, s, t, k_x, k_y, inorm, fused, un=None, staged=None):
        if staged is not None and k_x.dtype == torch.float32 and k_y.dtype == torch.float32:
            w_in16, w1h = self._staged_operands(u.t.dtype)
            return ops.fablock_full_staged(staged, u, w_in16, k_x.contiguous(), k_y.contiguous(), self.heads, inorm.eps, w1h,
                                           self.to_out[3].weight)
        if (fused and self.to_out[1].bias is None and self.to_out[3].bias is None
                and ops.fablock_full_supported(u, self.dim_head, self.to_out[3].out_channels)
                and self.to_out[1].out_channels == 64):
            # everything up to the block output in one kernel per sample, accumulators in TMEM: all contractions on tcgen05
            # (16x16 / 32x32), else the mma.sync phases + tcgen05 to_out convs
            if ops.fablock_tc_supported(u, self.dim_head, self.to_out[3].out_channels):
                return ops.fablock_tc(u, s, t, self.in_proj.weight, k_x, k_y, self.heads, inorm.eps, self.to_out[1].weight,
                                      self.to_out[3].weight)
            return ops.fablock_full(u, s, t, self.in_proj.weight, k_x, k_y, self.heads, inorm.eps, self.to_out[1].weight,
                                    self.to_out[3].weight)
        if fused:
            # in_proj -> contraction over H -> contraction over W -> InstanceNorm, u_phi stays in shared memory
            u_n = ops.fablock_core(u, s, t, self.in_proj.weight, k_x, k_y, self.heads, inorm.eps)
            h = conv_layer(u_n, self.to_out[1], act=ops.ACT_GELU)
        else:
            u_phi = conv_layer(un, self.in_proj)
            u_phi = ops.axial_contract(u_phi, k_x, self.heads, axis=0)
            u_phi = ops.axial_contract(u_phi, k_y, self.heads, axis=1)
            s2, t2 = norm_affine(u_phi, inorm)
            h = conv_layer(u_phi, self.to_out[1], pro=(s2, t2, ops.ACT_NONE), act=ops.ACT_GELU)
        return conv_layer(h, self.to_out[3], residual=skip)
