// Backward kernels of the latent propagator (the latent-space TRAINING rollout,
// LatentDynamics.forward = back-propagation through t_out steps of SimpleCNN, train_stage2_ns2d.py:126-141).
//
// What lives here is everything the forward engines cannot do:
//   lns_conv2d_wgrad     dW[o][i][ky][kx] += sum_{b,y,x} dy[b,y,x,o] * pro(x)[b, src(y + ky*d - pt, x + kx*d - pl), i]
//                        (the filter gradient of nn.Conv2d, any padding mode of the index map, stride 1, same-size output;
//                        `pro` = the per-(sample, channel) affine + activation the forward applied in its gather, recomputed
//                        here so that neither GroupNorm outputs nor GELU outputs have to be kept)
//   lns_chan_sum_accum   db[c] += sum over all pixels of dy                     (the bias gradient)
//   lns_act_bwd          dx = dy * act'(pre)                                     (exact-erf GELU / Swish derivative)
//   lns_group_norm_bwd   dx = dskip + GroupNorm'(x)^T dy, per-sample partials of dgamma / dbeta
//   lns_batch_sum_accum  grad[c] += sum_b part[b][c]
// The data gradient of a stride-1 same-size convolution is a convolution of dy with the flipped, transposed filter in the same
// padding mode, so it runs on the FORWARD engines (lns_conv2d; on the 16-bit modes the split-operand tcgen05 engines).
// All sums are formed in a fixed order (split-K partials + an ordered reduction, no atomics): gradients are bit-reproducible.
#include "common.cuh"

namespace lns {
namespace {

constexpr int kTO = 64, kTI = 64, kTK = 16;

struct WgradParams {
  const float* x;
  int64_t x_bstride;
  const float* pro_scale;  // [B][Cin] or null
  const float* pro_shift;  // [B][Cin] or null
  int pro_act;
  const float* dy;
  int64_t dy_bstride;
  ConvGeom g;
  float* part;  // [nsplit][taps][Cout][Cin]
  int nsplit;
  int64_t pix_per_split, npix;
  int tiles_o;
};

// grid (tiles_o * tiles_i, taps, nsplit), 256 threads: a 64 x 64 (o x i) tile of one tap over one slice of the pixels
__global__ void __launch_bounds__(256) wgrad_kernel(const WgradParams p) {
  __shared__ __align__(16) float dys[kTK][kTO];
  __shared__ __align__(16) float xs[kTK][kTI];
  const ConvGeom& g = p.g;
  const int tid = threadIdx.x;
  const int tile_o = blockIdx.x % p.tiles_o, tile_i = blockIdx.x / p.tiles_o;
  const int tap = blockIdx.y, split = blockIdx.z;
  const int ky = tap / g.KW, kx = tap % g.KW;
  const int o0 = tile_o * kTO, i0 = tile_i * kTI;
  const int HW = g.Hout * g.Wout;
  const int64_t k_begin = (int64_t)split * p.pix_per_split;
  const int64_t k_end = min(p.npix, k_begin + p.pix_per_split);
  const int lrow = tid >> 4, lc4 = (tid & 15) * 4;  // loader: pixel row of the tile, first of 4 channels
  const int to4 = (tid >> 4) * 4, ti4 = (tid & 15) * 4;  // compute: 4 x 4 register tile
  float acc[4][4];
#pragma unroll
  for (int a = 0; a < 4; ++a)
#pragma unroll
    for (int b = 0; b < 4; ++b) acc[a][b] = 0.f;
  const bool o_ok = o0 + lc4 < g.Cout, i_ok = i0 + lc4 < g.Cin;
  for (int64_t k0 = k_begin; k0 < k_end; k0 += kTK) {
    const int64_t pix = k0 + lrow;
    float4 dv = make_float4(0.f, 0.f, 0.f, 0.f), xv = dv;
    if (pix < k_end) {
      const int b = (int)(pix / HW);
      const int rem = (int)(pix - (int64_t)b * HW);
      const int yo = rem / g.Wout, xo = rem - yo * g.Wout;
      if (o_ok) dv = __ldg(reinterpret_cast<const float4*>(p.dy + (int64_t)b * p.dy_bstride + (int64_t)rem * g.Cout + o0 + lc4));
      int ysrc, xsrc;
      if (i_ok && conv_src(g, yo + ky * g.dil - g.pad_t, xo + kx * g.dil - g.pad_l, ysrc, xsrc)) {
        xv = __ldg(reinterpret_cast<const float4*>(p.x + (int64_t)b * p.x_bstride + ((int64_t)ysrc * g.Win + xsrc) * g.Cin + i0 + lc4));
        if (p.pro_scale) {
          const float4 s = __ldg(reinterpret_cast<const float4*>(p.pro_scale + (int64_t)b * g.Cin + i0 + lc4));
          const float4 t = __ldg(reinterpret_cast<const float4*>(p.pro_shift + (int64_t)b * g.Cin + i0 + lc4));
          xv.x = fmaf(xv.x, s.x, t.x); xv.y = fmaf(xv.y, s.y, t.y); xv.z = fmaf(xv.z, s.z, t.z); xv.w = fmaf(xv.w, s.w, t.w);
        }
        if (p.pro_act != LNS_ACT_NONE) {
          xv.x = apply_act(xv.x, p.pro_act); xv.y = apply_act(xv.y, p.pro_act);
          xv.z = apply_act(xv.z, p.pro_act); xv.w = apply_act(xv.w, p.pro_act);
        }
      }
    }
    __syncthreads();  // the previous tile has been consumed
    *reinterpret_cast<float4*>(&dys[lrow][lc4]) = dv;
    *reinterpret_cast<float4*>(&xs[lrow][lc4]) = xv;
    __syncthreads();
#pragma unroll
    for (int k = 0; k < kTK; ++k) {
      const float4 a = *reinterpret_cast<const float4*>(&dys[k][to4]);
      const float4 b = *reinterpret_cast<const float4*>(&xs[k][ti4]);
      const float av[4] = {a.x, a.y, a.z, a.w}, bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
      for (int r = 0; r < 4; ++r)
#pragma unroll
        for (int c = 0; c < 4; ++c) acc[r][c] = fmaf(av[r], bv[c], acc[r][c]);
    }
  }
  const int taps = g.KH * g.KW;
  float* out = p.part + ((int64_t)split * taps + tap) * g.Cout * g.Cin;
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    const int o = o0 + to4 + r;
    if (o >= g.Cout) continue;
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      const int i = i0 + ti4 + c;
      if (i < g.Cin) out[(int64_t)o * g.Cin + i] = acc[r][c];
    }
  }
}

// Tensor-core flavour: mma.sync.m16n8k16 on IEEE-half operands, BOTH operands split into hi + lo halves (a = rn(v), lo =
// rn(v - hi): 22 significant bits; three MMAs per product block, fp32 accumulation -- the split of csrc/conv_coarse.cu).  The
// split happens ONCE per element when a stage is written to shared memory, as packed pairs of two consecutive pixels (the K
// dimension), so that a fragment register is one 32-bit shared-memory load.  128 x 128 (o x i) tile, 32 pixels per stage, the
// next stage's global loads in flight during the MMAs; 16 warps, warp w owns rows 32 (w & 3) .. + 32 and columns 32 (w >> 2) .. + 32.
// Callers scale dy into the half range (loss scaling, lns_b200/train.py).
constexpr int kTKc = 32, kTT = 128, kLdP = kTT + 8;  // row stride 136 words: conflict-free fragment loads (bank = 8 t + g)
__device__ __forceinline__ void split_pack(float a, float b, uint32_t& hi, uint32_t& lo) {
  hi = pack2_h16<true>(a, b);
  const float2 back = unpack2_h16<true>(hi);
  lo = pack2_h16<true>(a - back.x, b - back.y);
}
__device__ __forceinline__ void bw_mma(float* c, const uint32_t* a, uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__global__ void __launch_bounds__(512) wgrad_tc_kernel(const WgradParams p) {
  __shared__ __align__(16) uint32_t dyh[kTKc / 2][kLdP], dyl[kTKc / 2][kLdP], xh[kTKc / 2][kLdP], xl[kTKc / 2][kLdP];
  const ConvGeom& g = p.g;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int tile_o = blockIdx.x % p.tiles_o, tile_i = blockIdx.x / p.tiles_o;
  const int tap = blockIdx.y, split = blockIdx.z;
  const int ky = tap / g.KW, kx = tap % g.KW;
  const int o0 = tile_o * kTT, i0 = tile_i * kTT;
  const int HW = g.Hout * g.Wout;
  const int64_t k_begin = (int64_t)split * p.pix_per_split;
  const int64_t k_end = min(p.npix, k_begin + p.pix_per_split);
  // loader: item = (pixel pair, 4-channel column); 16 pairs x 32 columns = 512 items per operand, one per thread (warp w = pair w)
  const int lc4 = (tid & 31) * 4;
  const int gq = lane >> 2, t = lane & 3;
  const int m0 = (warp & 3) * 32, n0 = (warp >> 2) * 32;
  float acc[2][4][4];
#pragma unroll
  for (int a = 0; a < 2; ++a)
#pragma unroll
    for (int b = 0; b < 4; ++b) acc[a][b][0] = acc[a][b][1] = acc[a][b][2] = acc[a][b][3] = 0.f;
  const bool o_ok = o0 + lc4 < g.Cout, i_ok = i0 + lc4 < g.Cin;
  float4 dv[1][2], xv[1][2];  // [item][pixel of the pair], RAW: the prologue is applied when the stage is written (after the
  int bsel[1][2];             // MMAs of the previous stage), so that these loads stay in flight; bsel = sample index, -1 = zero
  const int kb32 = (int)k_begin, ke32 = (int)k_end;  // (npix < 2^31 is checked by the launcher: 32-bit index arithmetic)
  auto load = [&](int k0) {
    // A warp loads the two pixels of ONE pair per stage.  Lane j < 2 does the index arithmetic of pixel j (two divisions, the
    // padding map) and the warp reads the three results by shuffle (the ncu source page of the first version had 46 % of the
    // samples in that arithmetic, redone by every lane for every pixel).
    int my_b = -1, my_rem = 0, my_src = -1;
    {
      const int j = lane & 1;
      const int pix = k0 + (tid >> 5) * 2 + j;
      if (pix < ke32) {
        my_b = pix / HW;
        my_rem = pix - my_b * HW;
        const int yo = my_rem / g.Wout, xo = my_rem - yo * g.Wout;
        int ysrc, xsrc;
        if (conv_src(g, yo + ky * g.dil - g.pad_t, xo + kx * g.dil - g.pad_l, ysrc, xsrc)) my_src = ysrc * g.Win + xsrc;
      }
    }
#pragma unroll
    for (int it = 0; it < 1; ++it) {
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int b = __shfl_sync(0xffffffffu, my_b, it * 2 + h);
        const int rem = __shfl_sync(0xffffffffu, my_rem, it * 2 + h);
        const int src = __shfl_sync(0xffffffffu, my_src, it * 2 + h);
        dv[it][h] = xv[it][h] = make_float4(0.f, 0.f, 0.f, 0.f);
        bsel[it][h] = -1;
        if (b >= 0) {
          if (o_ok) dv[it][h] = __ldg(reinterpret_cast<const float4*>(p.dy + (int64_t)b * p.dy_bstride + (int64_t)rem * g.Cout + o0 + lc4));
          if (i_ok && src >= 0) {
            xv[it][h] = __ldg(reinterpret_cast<const float4*>(p.x + (int64_t)b * p.x_bstride + (int64_t)src * g.Cin + i0 + lc4));
            bsel[it][h] = b;
          }
        }
      }
    }
  };
  auto prologue = [&](float4 v, int b) -> float4 {
    if (b < 0) return v;  // zero padding / outside the slice: stays zero
    if (p.pro_scale) {
      const float4 sc = __ldg(reinterpret_cast<const float4*>(p.pro_scale + (int64_t)b * g.Cin + i0 + lc4));
      const float4 sh = __ldg(reinterpret_cast<const float4*>(p.pro_shift + (int64_t)b * g.Cin + i0 + lc4));
      v.x = fmaf(v.x, sc.x, sh.x); v.y = fmaf(v.y, sc.y, sh.y); v.z = fmaf(v.z, sc.z, sh.z); v.w = fmaf(v.w, sc.w, sh.w);
    }
    if (p.pro_act != LNS_ACT_NONE) {  // the forward engines on this path use the same fast forms (conv_coarse.cu)
      v.x = apply_act_fast(v.x, p.pro_act); v.y = apply_act_fast(v.y, p.pro_act);
      v.z = apply_act_fast(v.z, p.pro_act); v.w = apply_act_fast(v.w, p.pro_act);
    }
    return v;
  };
  if (kb32 < ke32) load(kb32);
  for (int k0 = kb32; k0 < ke32; k0 += kTKc) {
    __syncthreads();  // the previous stage has been consumed
#pragma unroll
    for (int it = 0; it < 1; ++it) {
      const int pair = tid >> 5;
      uint4 h4, l4;
      split_pack(dv[it][0].x, dv[it][1].x, h4.x, l4.x); split_pack(dv[it][0].y, dv[it][1].y, h4.y, l4.y);
      split_pack(dv[it][0].z, dv[it][1].z, h4.z, l4.z); split_pack(dv[it][0].w, dv[it][1].w, h4.w, l4.w);
      *reinterpret_cast<uint4*>(&dyh[pair][lc4]) = h4;
      *reinterpret_cast<uint4*>(&dyl[pair][lc4]) = l4;
      const float4 x0 = prologue(xv[it][0], bsel[it][0]), x1 = prologue(xv[it][1], bsel[it][1]);
      split_pack(x0.x, x1.x, h4.x, l4.x); split_pack(x0.y, x1.y, h4.y, l4.y);
      split_pack(x0.z, x1.z, h4.z, l4.z); split_pack(x0.w, x1.w, h4.w, l4.w);
      *reinterpret_cast<uint4*>(&xh[pair][lc4]) = h4;
      *reinterpret_cast<uint4*>(&xl[pair][lc4]) = l4;
    }
    __syncthreads();
    if (k0 + kTKc < ke32) load(k0 + kTKc);  // in flight during the MMAs below
#pragma unroll
    for (int ks = 0; ks < kTKc / 16; ++ks) {
      // A[m = o][k = pixel]: fragment registers (g, 2t..), (g + 8, 2t..), (g, 2t + 8..), (g + 8, 2t + 8..) = pair rows t, t + 4
      uint32_t ah[2][4], al[2][4];
#pragma unroll
      for (int mt = 0; mt < 2; ++mt) {
        const int m = m0 + mt * 16 + gq;
        ah[mt][0] = dyh[ks * 8 + t][m]; ah[mt][1] = dyh[ks * 8 + t][m + 8]; ah[mt][2] = dyh[ks * 8 + t + 4][m]; ah[mt][3] = dyh[ks * 8 + t + 4][m + 8];
        al[mt][0] = dyl[ks * 8 + t][m]; al[mt][1] = dyl[ks * 8 + t][m + 8]; al[mt][2] = dyl[ks * 8 + t + 4][m]; al[mt][3] = dyl[ks * 8 + t + 4][m + 8];
      }
#pragma unroll
      for (int nt = 0; nt < 4; ++nt) {
        const int n = n0 + nt * 8 + gq;
        const uint32_t bh0 = xh[ks * 8 + t][n], bh1 = xh[ks * 8 + t + 4][n], bl0 = xl[ks * 8 + t][n], bl1 = xl[ks * 8 + t + 4][n];
#pragma unroll
        for (int mt = 0; mt < 2; ++mt) {
          bw_mma(acc[mt][nt], al[mt], bh0, bh1);
          bw_mma(acc[mt][nt], ah[mt], bl0, bl1);
          bw_mma(acc[mt][nt], ah[mt], bh0, bh1);
        }
      }
    }
  }
  const int taps = g.KH * g.KW;
  float* out = p.part + ((int64_t)split * taps + tap) * g.Cout * g.Cin;
#pragma unroll
  for (int mt = 0; mt < 2; ++mt)
#pragma unroll
    for (int half = 0; half < 2; ++half) {
      const int o = o0 + m0 + mt * 16 + gq + half * 8;
      if (o >= g.Cout) continue;
#pragma unroll
      for (int nt = 0; nt < 4; ++nt) {
        const int i = i0 + n0 + nt * 8 + 2 * t;
        if (i < g.Cin) out[(int64_t)o * g.Cin + i] = acc[mt][nt][half * 2];
        if (i + 1 < g.Cin) out[(int64_t)o * g.Cin + i + 1] = acc[mt][nt][half * 2 + 1];
      }
    }
}

// dW[o][i][tap] += sum over the splits, in split order
__global__ void wgrad_reduce_kernel(const float* part, int nsplit, int taps, int Cout, int Cin, float out_scale,
                                    const float* out_scale_dev, float* dW) {
  const int64_t n = (int64_t)taps * Cout * Cin;
  if (out_scale_dev) out_scale *= __ldg(out_scale_dev);
  for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < n; e += (int64_t)gridDim.x * blockDim.x) {
    float s = 0.f;
    for (int sp = 0; sp < nsplit; ++sp) s += part[(int64_t)sp * n + e];
    const int tap = (int)(e / ((int64_t)Cout * Cin));
    const int64_t oi = e - (int64_t)tap * Cout * Cin;
    dW[oi * taps + tap] += s * out_scale;
  }
}

// grid (ceil(C / 32), NS), block 256 = 32 channels x 8 pixel lanes: part[slice][c] = sum over the slice's samples and all
// pixels of dy (fixed order); lns_chan_sum_accum then reduces the NS slices with batch_sum_kernel.  (The first version ran
// ceil(C / 32) CTAs in all: 1.7 ms per bias gradient at 1024 samples.)
__global__ void __launch_bounds__(256) chan_sum_kernel(const float* dy, int64_t bstride, int B, int HW, int C, float* part) {
  __shared__ float red[8][32];
  const int cl = threadIdx.x & 31, pl = threadIdx.x >> 5;
  const int c = blockIdx.x * 32 + cl;
  float s = 0.f;
  if (c < C)
    for (int b = blockIdx.y; b < B; b += gridDim.y) {
      const float* src = dy + (int64_t)b * bstride + c;
      float sb = 0.f;
      for (int pix = pl; pix < HW; pix += 8) sb += __ldg(src + (int64_t)pix * C);
      s += sb;
    }
  red[pl][cl] = s;
  __syncthreads();
  if (pl == 0 && c < C) {
    float t = 0.f;
#pragma unroll
    for (int j = 0; j < 8; ++j) t += red[j][cl];
    part[(int64_t)blockIdx.y * C + c] = t;
  }
}

__device__ __forceinline__ float act_grad(float x, int act) {
  if (act == LNS_ACT_GELU) {
    const float cdf = 0.5f * (1.0f + erff(x * 0.70710678118654752440f));
    return cdf + x * 0.39894228040143267794f * expf(-0.5f * x * x);
  }
  if (act == LNS_ACT_SILU) {
    const float sg = 1.0f / (1.0f + expf(-x));
    return sg * (1.0f + x * (1.0f - sg));
  }
  return 1.0f;
}

__global__ void act_bwd_kernel(const float4* dy, const float4* pre, int64_t n4, int act, float4* dx) {
  for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < n4; e += (int64_t)gridDim.x * blockDim.x) {
    const float4 d = __ldg(dy + e), x = __ldg(pre + e);
    dx[e] = make_float4(d.x * act_grad(x.x, act), d.y * act_grad(x.y, act), d.z * act_grad(x.z, act), d.w * act_grad(x.w, act));
  }
}

// One CTA (256 threads) per sample.  y = gamma_c * (x - mu_g) * rstd_g + beta_c over groups of C/G channels x HW pixels:
//   dx = rstd_g * (gamma_c dy - m1_g - xhat * m2_g),  m1_g = mean_g(gamma dy),  m2_g = mean_g(gamma dy xhat)
//   dgamma_c = sum_pix dy xhat,  dbeta_c = sum_pix dy      (per sample; summed over the batch by lns_batch_sum_accum)
// Thread t owns channel t % C and every (256 / C)-th pixel (needs 256 % C == 0): reads are coalesced over channels.
__global__ void __launch_bounds__(256) gn_bwd_kernel(const float* x, int64_t x_bstride, const float* dy, int64_t dy_bstride,
                                                      const float* dskip, int64_t ds_bstride, int HW, int C, int G, float eps,
                                                      const float* gamma, float* dx, int64_t dx_bstride, float* dgamma_part,
                                                      float* dbeta_part) {
  __shared__ double sred[256][2];
  __shared__ float ch[4][256];   // per channel: sum x, sum x^2 -> later A = sum dy, Bq = sum dy*x
  __shared__ float gstat[3][256];  // per group: mean, rstd ; then m1, m2 (reused)
  const int b = blockIdx.x, tid = threadIdx.x;
  const int c = tid % C, lane_p = tid / C, np = 256 / C;
  const int cpg = C / G;
  const float* xb = x + (int64_t)b * x_bstride;
  const float* dyb = dy + (int64_t)b * dy_bstride;
  // pass 1: per-channel sums of x, x^2 and of dy, dy*x
  double sx = 0.0, sxx = 0.0;
  float sa = 0.f, sb = 0.f;
  for (int pix = lane_p; pix < HW; pix += np) {
    const float xv = __ldg(xb + (int64_t)pix * C + c), dv = __ldg(dyb + (int64_t)pix * C + c);
    sx += (double)xv; sxx += (double)xv * (double)xv;
    sa += dv; sb = fmaf(dv, xv, sb);
  }
  sred[tid][0] = sx; sred[tid][1] = sxx;
  ch[2][tid] = sa; ch[3][tid] = sb;
  __syncthreads();
  if (tid < C) {
    double tx = 0.0, txx = 0.0;
    float ta = 0.f, tb = 0.f;
    for (int j = 0; j < np; ++j) {
      tx += sred[j * C + tid][0]; txx += sred[j * C + tid][1];
      ta += ch[2][j * C + tid]; tb += ch[3][j * C + tid];
    }
    // thread tid reads rows {tid, C + tid, ...} and writes row tid, which no other thread reads: no barrier in between
    sred[tid][0] = tx; sred[tid][1] = txx;
    ch[0][tid] = ta; ch[1][tid] = tb;
  }
  __syncthreads();
  // per group: mean, rstd, m1, m2
  if (tid < G) {
    double tx = 0.0, txx = 0.0;
    for (int j = 0; j < cpg; ++j) { tx += sred[tid * cpg + j][0]; txx += sred[tid * cpg + j][1]; }
    const double n = (double)cpg * (double)HW;
    const double mean = tx / n;
    double var = txx / n - mean * mean;
    if (var < 0.0) var = 0.0;
    const double rstd = 1.0 / sqrt(var + (double)eps);
    double s1 = 0.0, s2 = 0.0;
    for (int j = 0; j < cpg; ++j) {
      const int cc = tid * cpg + j;
      const double gm = gamma ? (double)gamma[cc] : 1.0;
      const double A = (double)ch[0][cc], Bq = (double)ch[1][cc];
      s1 += gm * A;
      s2 += gm * (Bq - mean * A) * rstd;
    }
    gstat[0][tid] = (float)mean; gstat[1][tid] = (float)rstd;
    gstat[2][tid] = (float)(s1 / n);
    ch[2][tid] = (float)(s2 / n);  // m2 (ch[2] is free after the channel reduction)
  }
  __syncthreads();
  const int gi = c / cpg;
  const float mean = gstat[0][gi], rstd = gstat[1][gi], m1 = gstat[2][gi], m2 = ch[2][gi];
  if (tid < C && dgamma_part) {
    dgamma_part[(int64_t)b * C + tid] = (ch[1][tid] - mean * ch[0][tid]) * rstd;
    dbeta_part[(int64_t)b * C + tid] = ch[0][tid];
  }
  const float gm = gamma ? __ldg(gamma + c) : 1.0f;
  float* dxb = dx + (int64_t)b * dx_bstride;
  const float* dsb = dskip ? dskip + (int64_t)b * ds_bstride : nullptr;
  for (int pix = lane_p; pix < HW; pix += np) {
    const int64_t e = (int64_t)pix * C + c;
    const float xhat = (__ldg(xb + e) - mean) * rstd;
    float v = rstd * (gm * __ldg(dyb + e) - m1 - xhat * m2);
    if (dsb) v += __ldg(dsb + e);
    dxb[e] = v;
  }
}

// grad[c] += out_scale * sum_b part[b][c]; grid ceil(C / 32), block 256 = 32 channels x 8 batch lanes, fixed order
__global__ void __launch_bounds__(256) batch_sum_kernel(const float* part, int B, int C, float out_scale, const float* out_scale_dev,
                                                        float* grad) {
  __shared__ float red[8][32];
  if (out_scale_dev) out_scale *= __ldg(out_scale_dev);
  const int cl = threadIdx.x & 31, bl = threadIdx.x >> 5;
  const int c = blockIdx.x * 32 + cl;
  float s = 0.f;
  if (c < C)
    for (int b = bl; b < B; b += 8) s += __ldg(part + (int64_t)b * C + c);
  red[bl][cl] = s;
  __syncthreads();
  if (bl == 0 && c < C) {
    float t = 0.f;
#pragma unroll
    for (int j = 0; j < 8; ++j) t += red[j][cl];
    grad[c] += t * out_scale;
  }
}

// s2[0] = S, s2[1] = 1 / S with S = the power of two that brings max |x| (given as its bit pattern) to about `target`; S = 1 for
// zero / non-finite inputs.  Device-side so that a loss-scaled backward pass needs no host synchronisation (CUDA-graph capture).
__global__ void loss_scale_kernel(const uint32_t* absmax_bits, float target, float* s2) {
  const float amax = __uint_as_float(*absmax_bits);
  float S = 1.0f;
  if (amax > 0.f && amax < 3.0e38f) {
    int e = (int)floorf(log2f(target / amax));
    e = e < -60 ? -60 : (e > 60 ? 60 : e);
    S = exp2f((float)e);
  }
  s2[0] = S;
  s2[1] = 1.0f / S;
}

// out = x * (*scalar), elementwise
__global__ void scale_by_kernel(const float4* x, const float* scalar, int64_t n4, float4* out) {
  const float sc = __ldg(scalar);
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
    const float4 v = __ldg(x + i);
    out[i] = make_float4(v.x * sc, v.y * sc, v.z * sc, v.w * sc);
  }
}

// out (one uint32 word, zero-initialised by the caller) = bit pattern of max |x| (non-negative floats order like integers)
__global__ void absmax_kernel(const float* x, int64_t n, uint32_t* out) {
  float m = 0.f;
  for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < n; e += (int64_t)gridDim.x * blockDim.x) {
    const float v = fabsf(x[e]);
    m = (v > m || v != v) ? v : m;  // NaN propagates (its bit pattern is above every finite value)
  }
  m = __uint_as_float(__reduce_max_sync(0xffffffffu, __float_as_uint(m)));
  if ((threadIdx.x & 31) == 0) atomicMax(out, __float_as_uint(m));
}

// grid B, block 256 (needs 256 % C == 0): out[b][c] (+)= sum_pix dy[b][pix][c] * (x ? x[b][pix][c] : 1), pixels in a fixed order.
// x == null: gradient of a per-sample bias broadcast over the pixels; x != null: gradient of a per-(sample, channel) gate.
__global__ void __launch_bounds__(256) pixel_dot_kernel(const float* dy, int64_t dy_bstride, const float* x, int64_t x_bstride, int HW,
                                                         int C, int accumulate, float* out) {
  __shared__ float red[256];
  const int b = blockIdx.x, c = threadIdx.x % C, lp = threadIdx.x / C, np = 256 / C;
  const float* dyb = dy + (int64_t)b * dy_bstride;
  const float* xb = x ? x + (int64_t)b * x_bstride : nullptr;
  float s = 0.f;
  for (int pix = lp; pix < HW; pix += np) {
    const float d = __ldg(dyb + (int64_t)pix * C + c);
    s = xb ? fmaf(d, __ldg(xb + (int64_t)pix * C + c), s) : s + d;
  }
  red[threadIdx.x] = s;
  __syncthreads();
  if (threadIdx.x < C) {
    float t = 0.f;
    for (int j = 0; j < np; ++j) t += red[j * C + threadIdx.x];
    float* o = out + (int64_t)b * C + threadIdx.x;
    *o = accumulate ? *o + t : t;
  }
}

// out = x * (scale ? scale[b][c] : 1) + (skip ? skip : 0), elementwise over [B][HW][C]
__global__ void scale_add_kernel(const float4* x, const float* scale, const float4* skip, int64_t per_sample4, int C, int64_t total4,
                                 float4* out) {
  const int c4n = C >> 2;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total4; i += (int64_t)gridDim.x * blockDim.x) {
    float4 v = __ldg(x + i);
    if (scale) {
      const int64_t b = i / per_sample4;
      const int c = (int)((i - b * per_sample4) % c4n) * 4;
      const float4 sc = __ldg(reinterpret_cast<const float4*>(scale + b * C + c));
      v.x *= sc.x; v.y *= sc.y; v.z *= sc.z; v.w *= sc.w;
    }
    if (skip) {
      const float4 k = __ldg(skip + i);
      v.x += k.x; v.y += k.y; v.z += k.z; v.w += k.w;
    }
    out[i] = v;
  }
}

int wgrad_splits(int64_t npix, int tiles, int taps) {
  // enough CTAs for ~4 per SM, at least 256 pixels per slice
  const int sms = device_sm_count();
  int64_t want = (4LL * sms + (int64_t)tiles * taps - 1) / ((int64_t)tiles * taps);
  const int64_t cap = (npix + 255) / 256;
  if (want > cap) want = cap;
  if (want < 1) want = 1;
  if (want > 64) want = 64;
  return (int)want;
}

}  // namespace
}  // namespace lns

extern "C" {

int64_t lns_conv2d_wgrad_work_bytes(int B, int H, int W, int Cin, int Cout, int KH, int KW) {
  if (B <= 0 || H <= 0 || W <= 0 || Cin <= 0 || Cout <= 0 || KH <= 0 || KW <= 0) return -1;
  return (int64_t)64 * KH * KW * Cout * Cin * 4;  // wgrad_splits never cuts the pixels into more than 64 slices
}

int lns_conv2d_wgrad(const float* x, int64_t x_bstride, const float* pro_scale, const float* pro_shift, int pro_act,
                     const float* dy, int64_t dy_bstride, int B, int H, int W, int Cin, int Cout, int KH, int KW, int dil,
                     int pad_t, int pad_l, int pad_mode_h, int pad_mode_w, int tensor_core, float out_scale,
                     const float* out_scale_dev, float* work, float* dW, void* stream) {
  LNS_REQUIRE(x && dy && work && dW && B > 0 && H > 0 && W > 0, "lns_conv2d_wgrad: bad arguments");
  LNS_REQUIRE(Cin % 4 == 0 && Cout % 4 == 0, "lns_conv2d_wgrad: Cin and Cout must be multiples of 4 (got %d, %d)", Cin, Cout);
  LNS_REQUIRE(x_bstride % 4 == 0 && dy_bstride % 4 == 0 && (reinterpret_cast<uintptr_t>(x) & 15) == 0 &&
                  (reinterpret_cast<uintptr_t>(dy) & 15) == 0,
              "lns_conv2d_wgrad: activations must be 16-byte aligned");
  LNS_REQUIRE((pro_scale == nullptr) == (pro_shift == nullptr), "lns_conv2d_wgrad: pro_scale and pro_shift go together");
  LNS_REQUIRE(KH >= 1 && KW >= 1 && dil >= 1 && 2 * pad_t == dil * (KH - 1) && 2 * pad_l == dil * (KW - 1),
              "lns_conv2d_wgrad: only same-size stride-1 convolutions (2*pad == dil*(k-1))");
  lns::WgradParams p;
  p.x = x; p.x_bstride = x_bstride; p.pro_scale = pro_scale; p.pro_shift = pro_shift; p.pro_act = pro_act;
  p.dy = dy; p.dy_bstride = dy_bstride;
  lns::ConvGeom& g = p.g;
  g.B = B; g.Hin = H; g.Win = W; g.Cin = Cin; g.Hv = H; g.Wv = W;
  g.KH = KH; g.KW = KW; g.stride = 1; g.dil = dil; g.pad_t = pad_t; g.pad_l = pad_l;
  g.circ_h = pad_mode_h == LNS_PAD_CIRCULAR; g.circ_w = pad_mode_w == LNS_PAD_CIRCULAR;
  g.Hout = H; g.Wout = W; g.Cout = Cout; g.x_bstride = x_bstride; g.y_bstride = dy_bstride;
  const int tile = tensor_core ? lns::kTT : lns::kTO;
  const int tiles_o = (Cout + tile - 1) / tile, tiles_i = (Cin + tile - 1) / tile;
  const int taps = KH * KW;
  p.npix = (int64_t)B * H * W;
  p.nsplit = lns::wgrad_splits(p.npix, tiles_o * tiles_i, taps);
  p.pix_per_split = ((p.npix + p.nsplit - 1) / p.nsplit + lns::kTK - 1) / lns::kTK * lns::kTK;
  p.part = work;
  p.tiles_o = tiles_o;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (tensor_core) {
    LNS_REQUIRE(p.npix < (1ll << 31) - lns::kTKc, "lns_conv2d_wgrad: too many pixels for the tensor-core path");
    p.pix_per_split = ((p.npix + p.nsplit - 1) / p.nsplit + lns::kTKc - 1) / lns::kTKc * lns::kTKc;
    lns::wgrad_tc_kernel<<<dim3(tiles_o * tiles_i, taps, p.nsplit), 512, 0, st>>>(p);
  } else {
    lns::wgrad_kernel<<<dim3(tiles_o * tiles_i, taps, p.nsplit), 256, 0, st>>>(p);
  }
  int rc = lns::check_launch("wgrad_kernel");
  if (rc != LNS_OK) return rc;
  const int64_t n = (int64_t)taps * Cout * Cin;
  lns::wgrad_reduce_kernel<<<lns::cdiv(n, 256), 256, 0, st>>>(work, p.nsplit, taps, Cout, Cin, out_scale, out_scale_dev, dW);
  return lns::check_launch("wgrad_reduce_kernel");
}

int lns_chan_sum_slices(int B) { return B < 128 ? B : 128; }

int lns_chan_sum_accum(const float* dy, int64_t bstride, int B, int HW, int C, float out_scale, const float* out_scale_dev,
                       float* work, float* grad, void* stream) {
  LNS_REQUIRE(dy && grad && work && B > 0 && HW > 0 && C > 0, "lns_chan_sum_accum: bad arguments");
  const int ns = lns_chan_sum_slices(B);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  lns::chan_sum_kernel<<<dim3((C + 31) / 32, ns), 256, 0, st>>>(dy, bstride, B, HW, C, work);
  int rc = lns::check_launch("chan_sum_kernel");
  if (rc != LNS_OK) return rc;
  lns::batch_sum_kernel<<<(C + 31) / 32, 256, 0, st>>>(work, ns, C, out_scale, out_scale_dev, grad);
  return lns::check_launch("batch_sum_kernel");
}

int lns_act_bwd(const float* dy, const float* pre, int64_t n, int act, float* dx, void* stream) {
  LNS_REQUIRE(dy && pre && dx && n > 0 && n % 4 == 0, "lns_act_bwd: bad arguments (n must be a multiple of 4)");
  LNS_REQUIRE(((reinterpret_cast<uintptr_t>(dy) | reinterpret_cast<uintptr_t>(pre) | reinterpret_cast<uintptr_t>(dx)) & 15) == 0,
              "lns_act_bwd: pointers must be 16-byte aligned");
  const int64_t n4 = n / 4;
  int blocks = lns::cdiv(n4, 256);
  const int cap = lns::device_sm_count() * 16;
  if (blocks > cap) blocks = cap;
  lns::act_bwd_kernel<<<blocks, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      reinterpret_cast<const float4*>(dy), reinterpret_cast<const float4*>(pre), n4, act, reinterpret_cast<float4*>(dx));
  return lns::check_launch("act_bwd_kernel");
}

int lns_group_norm_bwd(const float* x, int64_t x_bstride, const float* dy, int64_t dy_bstride, const float* dskip,
                       int64_t dskip_bstride, int B, int HW, int C, int G, float eps, const float* gamma, float* dx,
                       int64_t dx_bstride, float* dgamma_part, float* dbeta_part, void* stream) {
  LNS_REQUIRE(x && dy && dx && B > 0 && HW > 0 && C > 0 && G > 0 && C % G == 0, "lns_group_norm_bwd: bad arguments");
  LNS_REQUIRE(C <= 256 && 256 % C == 0, "lns_group_norm_bwd: C must divide 256 (got %d)", C);
  LNS_REQUIRE((dgamma_part == nullptr) == (dbeta_part == nullptr), "lns_group_norm_bwd: dgamma_part and dbeta_part go together");
  lns::gn_bwd_kernel<<<B, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(x, x_bstride, dy, dy_bstride, dskip, dskip_bstride, HW, C,
                                                                              G, eps, gamma, dx, dx_bstride, dgamma_part, dbeta_part);
  return lns::check_launch("gn_bwd_kernel");
}

int lns_pixel_dot(const float* dy, int64_t dy_bstride, const float* x, int64_t x_bstride, int B, int HW, int C, int accumulate,
                  float* out, void* stream) {
  LNS_REQUIRE(dy && out && B > 0 && HW > 0 && C > 0 && C <= 256 && 256 % C == 0, "lns_pixel_dot: bad arguments (C must divide 256)");
  lns::pixel_dot_kernel<<<B, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(dy, dy_bstride, x, x_bstride, HW, C, accumulate, out);
  return lns::check_launch("pixel_dot_kernel");
}

int lns_scale_add(const float* x, const float* scale, const float* skip, int B, int HW, int C, float* out, void* stream) {
  LNS_REQUIRE(x && out && B > 0 && HW > 0 && C > 0 && C % 4 == 0, "lns_scale_add: bad arguments");
  LNS_REQUIRE(((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(out) | reinterpret_cast<uintptr_t>(skip) |
                reinterpret_cast<uintptr_t>(scale)) & 15) == 0, "lns_scale_add: pointers must be 16-byte aligned");
  const int64_t per4 = (int64_t)HW * C / 4, total4 = per4 * B;
  int blocks = lns::cdiv(total4, 256);
  const int cap = lns::device_sm_count() * 16;
  if (blocks > cap) blocks = cap;
  lns::scale_add_kernel<<<blocks, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      reinterpret_cast<const float4*>(x), scale, reinterpret_cast<const float4*>(skip), per4, C, total4, reinterpret_cast<float4*>(out));
  return lns::check_launch("scale_add_kernel");
}

int lns_loss_scale(const uint32_t* absmax_bits, float target, float* s2, void* stream) {
  LNS_REQUIRE(absmax_bits && s2 && target > 0.f, "lns_loss_scale: bad arguments");
  lns::loss_scale_kernel<<<1, 1, 0, reinterpret_cast<cudaStream_t>(stream)>>>(absmax_bits, target, s2);
  return lns::check_launch("loss_scale_kernel");
}

int lns_scale_by(const float* x, const float* scalar_dev, int64_t n, float* out, void* stream) {
  LNS_REQUIRE(x && scalar_dev && out && n > 0 && n % 4 == 0, "lns_scale_by: bad arguments (n must be a multiple of 4)");
  LNS_REQUIRE(((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(out)) & 15) == 0, "lns_scale_by: pointers must be 16-byte aligned");
  int blocks = lns::cdiv(n / 4, 256);
  const int cap = lns::device_sm_count() * 16;
  if (blocks > cap) blocks = cap;
  lns::scale_by_kernel<<<blocks, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(reinterpret_cast<const float4*>(x), scalar_dev, n / 4,
                                                                                  reinterpret_cast<float4*>(out));
  return lns::check_launch("scale_by_kernel");
}

int lns_absmax(const float* x, int64_t n, uint32_t* out_bits, void* stream) {
  LNS_REQUIRE(x && out_bits && n > 0, "lns_absmax: bad arguments");
  int blocks = lns::cdiv(n, 1024);
  const int cap = lns::device_sm_count() * 4;
  if (blocks > cap) blocks = cap;
  lns::absmax_kernel<<<blocks, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(x, n, out_bits);
  return lns::check_launch("absmax_kernel");
}

int lns_batch_sum_accum(const float* part, int B, int C, float out_scale, const float* out_scale_dev, float* grad, void* stream) {
  LNS_REQUIRE(part && grad && B > 0 && C > 0, "lns_batch_sum_accum: bad arguments");
  lns::batch_sum_kernel<<<(C + 31) / 32, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(part, B, C, out_scale, out_scale_dev, grad);
  return lns::check_launch("batch_sum_kernel");
}

}  // extern "C"
