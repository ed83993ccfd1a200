/*
 * lns_b200.h -- C-ABI of liblns_b200.so: the sm_100a kernels behind the LNS latent-rollout hot path
 * (autoencoder encode -> K latent-propagator steps -> decode).
 *
 * The reference (BaratiLab/LNS-Latent-Neural-PDE-Solver) is pure PyTorch and has no FFI layer of its
 * own: its boundary is the nn.Module API.  The Python drop-in modules in
 * /modules keep that API and call the functions below through ctypes (lns_b200/_C.py).  Every
 * function cites the reference call sites whose ATen library calls it replaces.
 *
 * Conventions
 *  - plain pointers + sizes, no torch types.  All pointers are DEVICE pointers owned by the caller
 *    (PyTorch allocates); the library never allocates or frees device memory and keeps no device
 *    state.  Exception: lns_last_error() returns a host string.
 *  - every call only ENQUEUES work on `stream` (a cudaStream_t passed as void*): no synchronisation,
 *    no allocation, no host callback -> legal inside CUDA-graph capture.
 *  - return value: 0 = OK, negative = error (LNS_E_*); text via lns_last_error().  An unsupported
 *    shape / dtype is an error, never a silent fallback.
 *  - activations are NHWC ("channel-last": [B][H][W][C], C contiguous) with an explicit batch stride in
 *    ELEMENTS; dtype LNS_F32, LNS_TF32, LNS_BF16 or LNS_F16.  NCHW fp32 exists only at the two ends of the path (the
 *    reference's tensors are NCHW fp32, modules/autoencoder2d.py:69-72).
 */
#ifndef LNS_B200_H
#define LNS_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define LNS_OK 0
#define LNS_E_INVALID (-1)     /* bad argument / unsupported shape */
#define LNS_E_CUDA (-2)        /* CUDA launch error, see lns_last_error() */
#define LNS_E_UNSUPPORTED (-3) /* combination not implemented by this path */

enum {
  LNS_F32 = 0,
  LNS_BF16 = 1,
  LNS_TF32 = 2 /* stored as fp32, every value rounded to nearest TF32 (10-bit mantissa) when written: the storage type of
                  the tf32 precision mode, whose convolutions run on tcgen05.mma.kind::tf32 (which ignores the low 13 bits) */
  ,
  LNS_F16 = 3 /* IEEE half (11-bit significand): the storage type of the fp16 precision mode -- the same tensor-core kernels
                 and HBM bytes as LNS_BF16 (tcgen05.mma.kind::f16 takes either format) with TF32-class rounding error;
                 conversions saturate to +-65504 */
};
enum { LNS_NHWC = 0, LNS_NCHW = 1 };
enum { LNS_ACT_NONE = 0, LNS_ACT_SILU = 1, LNS_ACT_GELU = 2 };
enum { LNS_PAD_ZEROS = 0, LNS_PAD_CIRCULAR = 1 };
/* weight formats produced by lns_pack_conv_weight */
enum {
  LNS_W_SIMT_F32 = 0, /* [tap][Cin][Cout] fp32 (CUDA-core validation path, any Cin/Cout)      */
  LNS_W_UMMA_BF16 = 1, /* [tap][Cin/64][Cout][64] bf16, K-major 128B-swizzled smem image for     */
                       /* tcgen05.mma (needs Cin % 64 == 0 and Cout % 16 == 0)                   */
  LNS_W_UMMA_TF32 = 2, /* [tap][Cin/32][Cout][32] fp32 rounded to TF32, same swizzled image, for   */
                       /* tcgen05.mma.kind::tf32 (needs Cin % 32 == 0 and Cout % 16 == 0)         */
  LNS_W_UMMA_F16 = 3,  /* LNS_W_UMMA_BF16's layout with IEEE-half elements (for LNS_F16 activations) */
  LNS_W_UMMA_F16X2 = 4 /* SPLIT filter: two LNS_W_UMMA_F16 images back to back, hi = rn_f16(w) and lo = rn_f16(w - hi)  */
                       /* (22 significant bits).  With LNS_F16 activations the conv issues 2 MMAs per K step        */
                       /* (filter rounding removed); with LNS_F32 activations the kernel splits them the same way in  */
                       /* its producer and issues 3 MMAs (fp32-class result on the f16 tensor path)                   */
};
/* which engine executes lns_conv2d */
enum {
  LNS_ENGINE_SIMT = 0, /* CUDA-core fp32 FMA (validation path; any shape)                                       */
  LNS_ENGINE_UMMA = 1, /* tcgen05 implicit GEMM, per-tap gather (any kernel/stride/dilation/resize, Cin%64==0)  */
  LNS_ENGINE_HALO = 2, /* tcgen05 implicit GEMM, shared-memory halo + resident filter: same-size 3x3 stride-1     */
                       /* convs with Cin == 64, Cout in {64,128} (the full-resolution layers)                     */
  LNS_ENGINE_LATENT = 3, /* tcgen05 implicit GEMM for the 8x8 circular latent grid, 128 -> 128 channels, dilation  */
                         /* 1|2: resident halos of 4 samples, streamed filter (the propagator's 3x3 convs)         */
  LNS_ENGINE_COARSE = 4  /* tcgen05 implicit GEMM for the coarse levels: same-size 3x3 stride-1 convs, Cin / Cout in */
                         /* {64,128}, dilation 1-3, any padding / nearest resize, 8x8 output blocks with resident    */
                         /* halos and a streamed filter.  Input 16-bit, or LNS_F32 = split into f16 hi + lo halo      */
                         /* planes in the producer (A.W = Ahi.W + Alo.W; with LNS_W_UMMA_F16X2 also + Ahi.Wlo)        */
};

const char* lns_version(void);
const char* lns_last_error(void);
/* sm count / compute capability of the current device; returns LNS_OK */
int lns_device_info(int* sm_count, int* cc_major, int* cc_minor);

/* ------------------------------------------------------------------------------------------------
 * lns_conv2d: one implicit-GEMM convolution  y = epilogue( conv( prologue( resize(x) ) ) )
 *
 * Replaces, in one kernel, the reference chains
 *   F.pad (circular / constant) + nn.Conv2d 3x3 / 1x1, stride 1|2, dilation 1..3
 *       modules/basics.py:249,252 (ResidualBlock), :326-328 (DownSampleBlock),
 *       train_stage2_ns2d.py:34-41 (DilatedResidualBlock), modules/autoencoder2d_half_periodic.py:36-52
 *   F.interpolate(scale_factor=2) / nn.Upsample(size) + conv   modules/basics.py:295-299,
 *       modules/autoencoder2d.py:134-136   (nearest resize folded into the gather, never materialised)
 *   nn.Linear / 1x1 conv                     modules/basics.py:345-348, modules/factorized_attention.py:114,140-142
 *   GroupNorm/InstanceNorm apply + Swish/GELU in front of a conv (prologue: per-(sample,channel) affine
 *       produced by lns_norm_finalize)      modules/basics.py:246-252
 *   bias, conditioning shift, GELU/Swish, residual add behind a conv (epilogue)
 *       train_stage2_ns2d.py:50-53, train_stage2_twophase_conditional.py:66-75
 *
 * index map: output pixel (yo,xo), tap (ky,kx) reads the VIRTUAL input (size Hv x Wv, the nearest-resized
 * x) at yv = yo*stride + ky*dil - pad_t (same for x).  pad mode circular wraps yv mod Hv, zeros gives 0
 * (after the prologue, exactly like padding the activated tensor).  source row = floor(yv*Hin/Hv).
 * epilogue: v = acc + bias[n] + sample_bias[b][n] + pre_add[b,yo,xo,n];  v = act(v);  v += residual.
 * ------------------------------------------------------------------------------------------------ */
typedef struct LnsConvDesc {
  /* input */
  const void* x;
  int32_t x_dtype, x_layout;
  int32_t B, Hin, Win, Cin;
  int64_t x_bstride;
  int32_t Hv, Wv; /* virtual (resized) input size; == Hin,Win when there is no resize */
  /* filter */
  int32_t KH, KW, stride, dil, pad_t, pad_l, pad_mode_h, pad_mode_w;
  const void* w; /* packed by lns_pack_conv_weight */
  int32_t w_format;
  int32_t engine;
  const float* bias;        /* [Cout] or NULL */
  const float* sample_bias; /* [B][Cout] or NULL */
  /* prologue: x' = act(x*scale[b][c] + shift[b][c]); NULL scale = identity */
  const float* pro_scale;
  const float* pro_shift;
  int32_t pro_act;
  /* epilogue */
  int32_t act;
  const void* pre_add; /* NHWC [B][Hout][Wout][Cout] or NULL, added before the activation */
  int32_t pre_add_dtype;
  int64_t pre_add_bstride;
  const void* residual; /* same shape as the output or NULL, added after the activation */
  int32_t res_dtype;
  int64_t res_bstride;
  /* output */
  void* y;
  int32_t y_dtype, y_layout;
  int32_t Hout, Wout, Cout;
  int64_t y_bstride;
  /* optional by-product (LNS_ENGINE_COARSE only, else must be NULL): per-channel CENTRED partial sums of the OUTPUT as stored,
   * stats[b][chunk][c] = (sum d, sum d^2, pivot, count) as float4 with d = y - pivot over the chunk's pixels, chunk =
   * 4 * (8x8 block of the sample) + epilogue warp, lns_conv_stats_chunks(Hout, Wout) chunks per sample -- the statistics of the
   * GroupNorm that follows (lns_norm_finalize_centred) without another read of the tensor (modules/basics.py:246-252:
   * GroupNorm -> Swish -> Conv, twice per ResidualBlock) */
  float* stats;
} LnsConvDesc;

int lns_conv2d(const LnsConvDesc* d, void* stream);
/* chunks per sample of LnsConvDesc.stats */
int lns_conv_stats_chunks(int Hout, int Wout);

/* Bytes of the packed image of an OIHW fp32 filter [Cout][Cin][KH][KW] in `format`. */
int64_t lns_packed_weight_bytes(int Cout, int Cin, int KH, int KW, int format);
/* Re-lay an OIHW fp32 device filter (nn.Conv2d.weight / nn.Linear.weight [out][in]) into `format`. */
int lns_pack_conv_weight(const float* w_oihw, int Cout, int Cin, int KH, int KW, int format, void* out,
                         void* stream);

/* ------------------------------------------------------------------------------------------------
 * Normalisation: nn.GroupNorm(32,C,eps=1e-6) modules/basics.py:18-24; nn.GroupNorm(1,C)
 * train_stage2_ns2d.py:33,45, modules/factorized_attention.py:113; nn.GroupNorm(8,C)
 * modules/autoencoder2d.py:149; nn.InstanceNorm2d modules/factorized_attention.py:139.
 * All of them become (1) per-(sample,channel) partial sums, (2) a tiny finalize that folds group
 * statistics, gamma/beta and an optional per-(sample,channel) pre-scale into y = x*scale + shift,
 * (3) the affine applied in the consuming conv's prologue or by lns_affine_act.
 * ------------------------------------------------------------------------------------------------ */
/* number of pixel chunks lns_chan_stats splits one sample into (deterministic, batch independent) */
int lns_chan_stats_chunks(int H, int W);
/* partial[b][chunk][c] = (sum, sum of squares) as float2 over the chunk's pixels of x */
int lns_chan_stats(const void* x, int dtype, int B, int H, int W, int C, int64_t bstride, float* partial,
                   void* stream);
/* scale/shift [B][C] of GroupNorm(G, C, eps) applied to (x * prescale) (prescale [B][C] or NULL);
 * gamma/beta [C] or NULL (no affine, InstanceNorm2d default). */
int lns_norm_finalize(const float* partial, int B, int nchunk, int C, int HW, int G, float eps,
                      const float* gamma, const float* beta, const float* prescale, float* scale,
                      float* shift, void* stream);
/* same from the centred partials of LnsConvDesc.stats ([b][chunk][c] float4 = sum d, sum d^2, pivot, count) */
int lns_norm_finalize_centred(const float* partial, int B, int nchunk, int C, int HW, int G, float eps,
                              const float* gamma, const float* beta, const float* prescale, float* scale,
                              float* shift, void* stream);
/* statistics + finalize in one call: one fused kernel when the sample has <= 1024 pixels (lns_chan_stats_chunks == 1),
 * otherwise lns_chan_stats + lns_norm_finalize through `partial_ws` (B*nchunk*C*2 floats; may be NULL when nchunk == 1) */
int lns_group_norm_affine(const void* x, int dtype, int B, int H, int W, int C, int64_t bstride, int G, float eps,
                          const float* gamma, const float* beta, const float* prescale, float* partial_ws,
                          float* scale, float* shift, void* stream);
/* GroupNorm statistics + apply + activation in ONE kernel for samples of <= 16384 elements, staged raw in shared memory (the
 * latent-grid layers: 8x8x128, 7x15x128 ...): y = act(GroupNorm(G)(x * prescale)); lns_group_norm_act_supported() tells. */
int lns_group_norm_act_supported(int H, int W, int C);
int lns_group_norm_act(const void* x, int dtype, int B, int H, int W, int C, int64_t bstride, int G, float eps,
                       const float* gamma, const float* beta, const float* prescale, int act, void* y, int y_dtype,
                       int64_t y_bstride, void* stream);
/* y = act(x*scale[b][c] + shift[b][c]) (+ NULL scale -> activation only); NHWC in/out */
int lns_affine_act(const void* x, int x_dtype, int64_t x_bstride, int B, int HW, int C, const float* scale,
                   const float* shift, int act, void* y, int y_dtype, int64_t y_bstride, void* stream);
/* nn.LayerNorm over C per token, then + pe[token index within the sample]
 * (SABlock: modules/basics.py:384-386 -- pe is added AFTER the norm; PoolingReducer LN
 * modules/factorized_attention.py:80).  x,y: [B][n][C] rows; pe: [>=n][C] fp32 or NULL. */
int lns_layernorm(const void* x, int x_dtype, int B, int n, int C, const float* gamma, const float* beta,
                  float eps, const float* pe, void* y, int y_dtype, void* stream);

/* ------------------------------------------------------------------------------------------------
 * SABlock attention core  softmax(q k^T * scale) v   modules/basics.py:395-398
 * qkv: [B][n][3*heads*dh] rows (q | k | v, each (head, dh) major); out: [B][n][heads*dh]
 * ------------------------------------------------------------------------------------------------ */
int lns_attention(const void* qkv, int dtype, int B, int n, int heads, int dh, float scale, void* out,
                  int out_dtype, void* stream);

/* SABlock in ONE kernel (csrc/sablock_fused.cu):  y = x + proj_out( softmax(q k^T * scale) v ),  q|k|v = Linear(LN(x) + pe)
 * -- modules/basics.py:384-404 including the LayerNorm, the positional table added after the norm, the three input linears,
 * the attention core, the output projection and the residual.  x, y: token rows [B][n][128], LNS_BF16 | LNS_F16; wqkv16
 * [3*heads*64][128] = to_q | to_k | to_v weights stacked, wproj16 [128][heads*64], both in the activation's 16-bit format;
 * bv / bproj fp32 biases or NULL; pe [>= n][128] fp32 or NULL.  Needs dim = 128, dim_head = 64, n <= 128. */
int lns_sablock_fused_supported(int n, int dim, int heads, int dim_head);
int lns_sablock_fused(const void* x, int dtype, int B, int n, int heads, const float* ln_g, const float* ln_b, float ln_eps,
                      const float* pe, const void* wqkv16, const float* bv, const void* wproj16, const float* bproj,
                      float scale, void* y, void* stream);

/* ------------------------------------------------------------------------------------------------
 * FABlock2D pieces  modules/factorized_attention.py:144-159
 * ------------------------------------------------------------------------------------------------ */
/* mean over one spatial axis: axis 0 -> out [B][W][C] (mean over H); axis 1 -> out [B][H][C] (PoolingReducer,
 * modules/factorized_attention.py:86-94; the bias-free linears in front of the mean commute with it) */
int lns_axis_mean(const void* x, int dtype, int B, int H, int W, int C, int64_t bstride, int axis, float* out,
                  void* stream);
/* LowRankKernel: qk [B][n][2*heads*d] -> rotary with the host-computed tables cos_tab/sin_tab [n][d/2]
 * (angle[i][f] = linspace(0,1,n)[i] * (scale/min_freq) * inv_freq[f]) ->
 * K[b][h][i][j] = <q_i, k_j> * scaling   modules/factorized_attention.py:43-69, modules/embedding.py:163-186 */
int lns_lowrank_kernel(const void* qk, int dtype, int B, int n, int heads, int d, const float* cos_tab,
                       const float* sin_tab, float scaling, float* K, void* stream);
/* axial contraction over axis 0 (H): out[b,i,m,(h,c)] = sum_j K[b,h,i,j] u[b,j,m,(h,c)]
 *                   over axis 1 (W): out[b,i,l,(h,c)] = sum_m K[b,h,l,m] u[b,i,m,(h,c)]
 * (the two einsums at modules/factorized_attention.py:157-158); u,out NHWC with C = heads*ch */
int lns_axial_contract(const void* u, int dtype, int B, int H, int W, int heads, int ch, const float* K,
                       int axis, void* out, int out_dtype, void* stream);

/* Fused FABlock2D core for the bf16 path (one CTA per (sample, head); u_phi never leaves shared memory):
 *   in_proj (GroupNorm(1) folded into a per-sample filter/bias) -> contraction over H with Kx -> contraction over W with
 *   Ky -> InstanceNorm2d (eps, no affine) -> out [B][H][W][heads*64] bf16, ready for to_out's 1x1 convs.
 * u: NHWC [B][H][W][64], dtype LNS_BF16 or LNS_F16 (the block's RAW input; out has the same dtype); gn_scale/gn_shift [B][64] from lns_group_norm_affine(G=1);
 * w_in_proj: the nn.Conv2d weight [heads*64][64] fp32 as stored; Kx [B][heads][H][H], Ky [B][heads][W][W] fp32.
 * modules/factorized_attention.py:146-158 + the InstanceNorm2d of :139.  lns_fablock_core_supported() tells whether the
 * shape fits (H, W <= 48 and H*W*144 B + ~50 KB of shared memory); otherwise use the unfused entry points above. */
int lns_fablock_core_supported(int H, int W, int dim, int dim_head);
/* FABlock2D pre-pass: one read of the block input u [B][H][W][C] -> GroupNorm(1,C) affine scale/shift [B][C] AND the two
 * pooled, normalised tensors pooled_x [B][H][C] = mean_W(GN(u)), pooled_y [B][W][C] = mean_H(GN(u)) (fp32) that
 * PoolingReducer consumes (modules/factorized_attention.py:86-94,113) -- no normalised copy of u is written. */
int lns_fablock_prepass(const void* u, int dtype, int B, int H, int W, int C, int64_t bstride, float eps,
                        const float* gamma, const float* beta, float* scale, float* shift, float* pooled_x,
                        float* pooled_y, void* stream);
int lns_fablock_core(const void* u, int dtype, int B, int H, int W, int heads, const float* gn_scale, const float* gn_shift,
                     const float* w_in_proj, const float* Kx, const float* Ky, float eps, void* out, void* stream);
/* lns_fablock_prepass that ALSO writes `staged` [B][H*W][64] (same 16-bit dtype as u): the normalised sample GN(u) as the
 * byte image of lns_fablock_full_staged's shared-memory tile (pixel row s = pixel s ^ ((s >> log2 W) & 7), 16-byte chunk ch
 * at ch ^ (s & 7)), so that kernel fetches a head's input with linear bulk copies.  C = 64, H, W in {16, 32}. */
int lns_fablock_prepass_staged(const void* u, int dtype, int B, int H, int W, int C, int64_t bstride, float eps,
                               const float* gamma, const float* beta, float* scale, float* shift, float* pooled_x,
                               float* pooled_y, void* staged, void* stream);

/* Propagator FFN in ONE kernel (csrc/ffn_fused.cu):  y = x + W2 . GELU( W1 . (x*scale[b] + shift[b]) )  -- GroupNorm(1,C) apply,
 * 1x1 conv, GELU, 1x1 conv and the residual add of train_stage2_ns2d.py:44-53 (both convs bias-free, C = hidden = 128).
 * x, y: NHWC rows [B][HW][128] LNS_BF16 | LNS_F16 with batch strides in elements; scale/shift [B][128] from
 * lns_group_norm_affine; w1_packed / w2_packed: lns_pack_conv_weight(..., LNS_W_UMMA_BF16 | _F16) images of the two 1x1
 * filters.  Both GEMMs run on tcgen05 with the hidden activation never leaving the SM. */
int lns_ffn_fused_supported(int C, int hidden);
int lns_ffn_fused(const void* x, int dtype, int B, int HW, int C, int64_t x_bstride, const float* scale, const float* shift,
                  const void* w1_packed, const void* w2_packed, void* y, int64_t y_bstride, void* stream);

/* FABlock2D pooled branch of one axis in ONE kernel (csrc/fa_axis.cu): pooled [B][n][64] fp32 (lns_fablock_prepass) ->
 * K [B][heads][n][n] fp32.  Replaces to_in (1x1 conv), PoolingReducer (Linear, LayerNorm, Linear+GELU, Linear+bias),
 * LowRankKernel.to_qk, the rotary embedding and q k^T (modules/factorized_attention.py:43-94,114-121).  Host-prepared
 * operands: w1t [64][64] = (reducer.to_in.weight @ to_in.weight)^T (the two bias-free linears composed), wf1t [64][128] and
 * wf2t [128][64] = out_ffn.1 / out_ffn.3 weights transposed, bf2 [64], wqk16 [2*heads*128][64] = to_qk.weight as LNS_BF16 or
 * LNS_F16 (dtype16), cos/sin tables [n][64] as for lns_lowrank_kernel.  The small layers run in fp32, to_qk and q k^T on
 * tensor cores with 16-bit operands.  Needs dim = hidden = latent = 64, dim_head*kernel_multiplier = 128, n <= 64. */
int lns_fa_axis_kernel_supported(int n, int dim, int hidden, int latent, int heads, int d);
int lns_fa_axis_kernel(const float* pooled, int dtype16, int B, int n, int heads, const float* w1t, const float* ln_g,
                       const float* ln_b, float ln_eps, const float* wf1t, const float* wf2t, const float* bf2,
                       const void* wqk16, const float* cos_tab, const float* sin_tab, float scaling, float* K, void* stream);

/* FABlock2D in ONE kernel per sample (csrc/fablock_full.cu): lns_fablock_core's phases with the InstanceNorm2d folded into
 * to_out[1] and BOTH 1x1 convolutions of to_out on tcgen05 -- the fp32 accumulator of all H*W pixels x 64 output channels
 * stays in tensor memory across the heads, the [H][W][heads*64] tensor never exists in HBM:
 *   out = Conv1x1_2( GELU( Conv1x1_1( InstanceNorm( Ky . Kx . in_proj(GN(u)) ) ) ) ) + u
 * (modules/factorized_attention.py:146-159 except the pooled branch that produces Kx, Ky).  w_out1: to_out.1.weight
 * [64][heads*64] fp32, w_out2: to_out.3.weight [64][64] fp32 (neither has a bias in the reference).  u/out NHWC
 * [B][H][W][64], LNS_BF16 or LNS_F16.  Needs H, W <= 32 (H * ceil8(W) <= 1024 pixel rows = 512 TMEM columns). */
int lns_fablock_full_supported(int H, int W, int dim, int dim_head, int dim_out);
int lns_fablock_full(const void* u, int dtype, int B, int H, int W, int heads, const float* gn_scale, const float* gn_shift,
                     const float* w_in_proj, const float* Kx, const float* Ky, float eps, const float* w_out1,
                     const float* w_out2, void* out, void* stream);
/* lns_fablock_full on pre-staged operands with a producer warp (csrc/fablock_full.cu, fablock_full2_kernel): same
 * mathematics (modules/factorized_attention.py:146-159), but the GroupNorm is applied by lns_fablock_prepass_staged (u_staged),
 * the in_proj slices arrive as 16-bit rows padded to 72 elements (w_in16 [heads][64][72], dtype of u, prepared once per
 * parameter version) and to_out.1.weight head-major (w1h [heads][64 out][64 in] fp32): every operand of a head is a bulk
 * copy issued by one producer lane, which also issues the tcgen05.mma of to_out[1]; u is the RAW input (skip connection).
 * H, W in {16, 32}. */
int lns_fablock_full_staged_supported(int H, int W, int dim, int dim_head, int dim_out);
int lns_fablock_full_staged(const void* u_staged, const void* u, int dtype, int B, int H, int W, int heads,
                            const void* w_in16, const float* Kx, const float* Ky, float eps, const float* w1h,
                            const float* w_out2, void* out, void* stream);

/* lns_fablock_full with EVERY contraction on tcgen05 (csrc/fablock_tc.cu): in_proj, both axial contractions (block-diagonal
 * kernel matrices x the pixel rows read as an MN-major operand) and both to_out convolutions are tcgen05.mma batches with
 * TMEM -> shared-memory drains in between; a 32x32 sample runs on a cluster of two CTAs (rows exchanged through distributed
 * shared memory), a 16x16 sample on one CTA.  Same arguments and result as lns_fablock_full; covers H == W in {16, 32}. */
int lns_fablock_tc_supported(int H, int W, int dim, int dim_head, int dim_out);
int lns_fablock_tc(const void* u, int dtype, int B, int H, int W, int heads, const float* gn_scale, const float* gn_shift,
                   const float* w_in_proj, const float* Kx, const float* Ky, float eps, const float* w_out1,
                   const float* w_out2, void* out, void* stream);

/* ------------------------------------------------------------------------------------------------
 * layout / misc
 * ------------------------------------------------------------------------------------------------ */
/* NCHW fp32 <-> NHWC (fp32|bf16); bstrides in elements */
int lns_nchw_to_nhwc(const float* x, int B, int C, int H, int W, int64_t x_bstride, void* y, int y_dtype,
                     int64_t y_bstride, void* stream);
int lns_nhwc_to_nchw(const void* x, int x_dtype, int B, int H, int W, int C, int64_t x_bstride, float* y,
                     int64_t y_bstride, void* stream);
/* decoder output projection: y[b][n][pix] (NCHW fp32) = bias[n] + sum_c w[n][c] * act(x[b][pix][c]*scale[b][c] + shift[b][c]),
 * Cout <= 4 -- the GroupNorm -> Swish -> Conv1x1(C -> in_channels) tail, modules/autoencoder2d.py:149-151.
 * x NHWC (fp32|bf16); w [Cout][C] fp32 (the nn.Conv2d weight as is); scale/shift [B][C] or NULL. */
int lns_pointwise_proj(const void* x, int dtype, int B, int HW, int C, int64_t x_bstride, const float* w,
                       const float* bias, int Cout, const float* scale, const float* shift, int act, float* y,
                       int64_t y_bstride, void* stream);
/* the same with a two-level output index for the rollout engine: the B samples are `B / group` rollout steps of `group`
 * trajectories each (step-major), sample s is written at y + (s % group) * y_bstride + (s / group) * y_gstride -- slot
 * (trajectory, step) of the [B, K, C, Ly, Lx] result of LatentDynamics.predict (torch.stack(dim=1), train_stage2_ns2d.py:157) */
int lns_pointwise_proj_steps(const void* x, int dtype, int B, int HW, int C, int64_t x_bstride, const float* w,
                             const float* bias, int Cout, const float* scale, const float* shift, int act, float* y,
                             int64_t y_bstride, int group, int64_t y_gstride, void* stream);
/* sinusoidal embedding cat(cos(p f), sin(p f)), f_i = exp(-ln(max_period) i / (dim/2))
 * modules/cond_utils.py:19-38 */
int lns_fourier_embedding(const float* param, int B, int dim, float max_period, float* out, void* stream);
/* y = x * (1 + gate[b][c]) -- conditional propagator gate, train_stage2_twophase_conditional.py:74 */
int lns_channel_gate(const void* x, int dtype, int B, int HW, int C, const float* gate, void* y, int y_dtype,
                     void* stream);

/* Validation metric right behind the path: per frame f = (trajectory, step, channel) of two fp32
 * tensors laid out like the rollout result [B][K][C][Ly][Lx] (P = Ly*Lx contiguous values per frame):
 *   out[f] = (sum (pred-target)^2, sum target^2, sum target)
 * from which relative_lp_loss (training_utils.py:9-23, reduce_dim (3,4) frame-wise and (1,3,4) sequence-wise,
 * train_stage2_ns2d.py:254-257) of the affinely de-normalised fields follows on the host without the fields ever leaving the
 * device.  One read of both tensors, deterministic. */
int lns_frame_sums(const float* pred, const float* target, int64_t frames, int P, float* out, void* stream);
/* The same three sums of the DE-NORMALISED frames for datasets whose de-normalisation is not one affine map -- the two-phase
 * dataset (dataset/twophase_flow_stage2.py:369-389): per channel c = f % C, v = x*scale[c] + shift[c]; flags[c] & 1 zeroes the
 * four border lines (Dirichlet walls of the velocity channels); flags[c] & 2 clamps to [clamp_lo, clamp_hi] (vof).  Applied to
 * prediction and target alike (train_stage2_twophase.py:251-252).  frames = B*K*C frames of H x W values. */
int lns_frame_sums_denorm(const float* pred, const float* target, int64_t frames, int H, int W, int C, const float* scale,
                          const float* shift, const int* flags, float clamp_lo, float clamp_hi, float* out, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Spectral convolution (FNO layer) as truncated DFTs with the complex mode-weight multiply fused:
 *   rfft2 -> corner blocks [:m1,:m2], [-m1:,:m2] x weights (x per-sample complex emb) -> irfft2
 * modules/basics.py:129-148 (SpectralConv2d), modules/fourier_cond.py:52-81 (conditioned variant).
 * x: NHWC [B][H][W][Ci]; w_modes: the two state_dict tensors weights1/weights2 [Ci][Co][m1][m2][2] re-laid
 * (host side, once) as [block 0|1][m1][m2][Ci][Co][re,im] fp32; emb: NULL or [B][m1][m2][block][re,im] fp32, which
 * is exactly FreqLinear's output (modules/fourier_cond.py:25-29) before view_as_complex; out: NHWC fp32
 * [B][H][W][Co].  work: scratch of lns_spectral_work_bytes() bytes.  Needs 2*m1 <= H and m2 <= W/2+1.
 * ------------------------------------------------------------------------------------------------ */
int64_t lns_spectral_work_bytes(int B, int H, int W, int Ci, int Co, int m1, int m2);
int lns_spectral_conv2d(const void* x, int x_dtype, int B, int H, int W, int Ci, int Co, int m1, int m2,
                        const float* w_modes, const float* emb, void* work, float* out, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Backward kernels of the latent propagator: the training rollout LatentDynamics.forward(z_in, z_out, loss_fn)
 * (train_stage2_ns2d.py:126-141; SW / two-phase scripts :126-142) back-propagates through t_out steps of SimpleCNN
 * (train_stage2_ns2d.py:25-87).  These replace what autograd derives from F.conv2d / F.group_norm / F.gelu there.
 * The DATA gradient of a same-size stride-1 convolution is lns_conv2d with the flipped, transposed filter; everything
 * else is below.  fp32 NHWC activations; all sums in a fixed order (split partials + ordered reduction, no atomics).
 *
 * lns_conv2d_wgrad: dW[o][i][ky][kx] (OIHW fp32, ACCUMULATED into) += sum_{b,y,x} dy[b][y][x][o] *
 *   pro(x)[b][src(y + ky*dil - pad_t, x + kx*dil - pad_l)][i], where src() is lns_conv2d's index map (zeros / circular per
 *   axis) and pro = the forward's gather prologue act(x*scale[b][i] + shift[b][i]) (scale/shift NULL: none).  Needs
 *   2*pad == dil*(k-1), Cin % 4 == 0, Cout % 4 == 0.  work: lns_conv2d_wgrad_work_bytes() bytes of scratch.  tensor_core = 1: the
 *   products run on mma.sync.m16n8k16 IEEE-half with both operands split hi + lo (3 MMAs, 22-bit operands; dy must be scaled
 *   into the half range), 0: CUDA-core fp32 FMA.  Every
 *   accumulating function multiplies its sum by out_scale first (1 / loss scale of a scaled backward pass).
 * lns_absmax: *out_bits (zero-initialised by the caller) = bit pattern of max |x| -- chooses the loss scale
 * lns_chan_sum_accum: grad[c] += sum_{b,pix} dy[b][pix][c]                       (bias gradient)
 * lns_act_bwd: dx = dy * act'(pre) elementwise (LNS_ACT_GELU exact erf, LNS_ACT_SILU)
 * lns_group_norm_bwd: x, dy -> dx (+ dskip if not NULL) for GroupNorm(G, C, eps) with weight gamma (NULL: ones); the
 *   statistics are recomputed from x.  dgamma_part / dbeta_part: [B][C] per-sample partials (both NULL: skipped).
 *   C must divide 256.
 * lns_batch_sum_accum: grad[c] += sum_b part[b][c]
 * ------------------------------------------------------------------------------------------------ */
int64_t lns_conv2d_wgrad_work_bytes(int B, int H, int W, int Cin, int Cout, int KH, int KW);
int lns_conv2d_wgrad(const float* x, int64_t x_bstride, const float* pro_scale, const float* pro_shift, int pro_act,
                     const float* dy, int64_t dy_bstride, int B, int H, int W, int Cin, int Cout, int KH, int KW, int dil,
                     int pad_t, int pad_l, int pad_mode_h, int pad_mode_w, int tensor_core, float out_scale,
                     const float* out_scale_dev, float* work, float* dW, void* stream);
int lns_chan_sum_slices(int B); /* work of lns_chan_sum_accum: lns_chan_sum_slices(B) * C floats */
int lns_chan_sum_accum(const float* dy, int64_t bstride, int B, int HW, int C, float out_scale, const float* out_scale_dev,
                       float* work, float* grad, void* stream);
int lns_act_bwd(const float* dy, const float* pre, int64_t n, int act, float* dx, void* stream);
int lns_group_norm_bwd(const float* x, int64_t x_bstride, const float* dy, int64_t dy_bstride, const float* dskip,
                       int64_t dskip_bstride, int B, int HW, int C, int G, float eps, const float* gamma, float* dx,
                       int64_t dx_bstride, float* dgamma_part, float* dbeta_part, void* stream);
int lns_batch_sum_accum(const float* part, int B, int C, float out_scale, const float* out_scale_dev, float* grad, void* stream);
int lns_absmax(const float* x, int64_t n, uint32_t* out_bits, void* stream);
/* device-side loss scale (no host synchronisation: the scaled backward pass can be captured in a CUDA graph): s2[0] = S, s2[1] =
 * 1 / S, S = the power of two that brings the absmax (bit pattern from lns_absmax) to about `target`; lns_scale_by: out = x * (*s).
 * The accumulating functions above multiply by out_scale * (out_scale_dev ? *out_scale_dev : 1). */
int lns_loss_scale(const uint32_t* absmax_bits, float target, float* s2, void* stream);
int lns_scale_by(const float* x, const float* scalar_dev, int64_t n, float* out, void* stream);
/* conditional propagator (train_stage2_twophase_conditional.py:66-75): out[b][c] (+)= sum_pix dy[b][pix][c] * (x ? x[b][pix][c] : 1)
 * -- gradient of the per-sample shift Linear(emb) (x NULL) and of the gate (1 + g) (x = the gated activation); C divides 256.
 * lns_scale_add: out = x * scale[b][c] + skip (scale / skip may be NULL), contiguous fp32 [B][HW][C] */
int lns_pixel_dot(const float* dy, int64_t dy_bstride, const float* x, int64_t x_bstride, int B, int HW, int C, int accumulate,
                  float* out, void* stream);
int lns_scale_add(const float* x, const float* scale, const float* skip, int B, int HW, int C, float* out, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* LNS_B200_H */
