// CUDA-core implicit-GEMM convolution: the fp32 validation engine (LNS_ENGINE_SIMT) and the engine for
// the tiny-channel ends of the path (Cin or Cout in {1,3,4,16}) where a tensor-core tile would be >75% padding.
// Also: filter re-layout kernels (lns_pack_conv_weight).
//
// GEMM view: M = B*Hout*Wout output pixels, N = Cout, K = KH*KW*Cin.  64 x BN tile per CTA, 16-deep K chunks,
// 4 x (BN/16) accumulators per thread, fp32 FMA accumulation in a fixed order (deterministic, batch independent).
#include <stdlib.h>

#include "common.cuh"

namespace lns {

struct SimtParams {
  ConvGeom g;
  const void* x;
  int x_dtype, x_layout;
  const float* w;
  const float* bias;
  const float* sample_bias;
  const float* pro_scale;
  const float* pro_shift;
  int pro_act;
  int act;
  const void* pre_add;
  int pre_add_dtype;
  int64_t pre_add_bstride;
  const void* residual;
  int res_dtype;
  int64_t res_bstride;
  void* y;
  int y_dtype, y_layout;
  int M;
};

template <int BN>
__global__ void __launch_bounds__(256) conv_simt_kernel(const SimtParams p) {
  constexpr int BM = 64, BK = 16, TN = BN / 16;
  __shared__ __align__(16) float As[BK][BM + 4];
  __shared__ __align__(16) float Bs[BK][BN];
  const ConvGeom& g = p.g;
  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;
  const int m0 = blockIdx.x * BM, n0 = blockIdx.y * BN;
  const int HWo = g.Hout * g.Wout;
  const int64_t chan_stride = (p.x_layout == LNS_NCHW) ? (int64_t)g.Hin * g.Win : 1;

  // rows this thread gathers for the A tile: k_l fixed, 4 pixel rows
  const int k_l = tid & 15;
  int lb[4], ly[4], lx[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    int m = m0 + (tid >> 4) + 16 * i;
    if (m < p.M) {
      lb[i] = m / HWo;
      int r = m - lb[i] * HWo;
      ly[i] = r / g.Wout;
      lx[i] = r - ly[i] * g.Wout;
    } else {
      lb[i] = -1; ly[i] = 0; lx[i] = 0;
    }
  }

  float acc[4][TN];
#pragma unroll
  for (int r = 0; r < 4; ++r)
#pragma unroll
    for (int j = 0; j < TN; ++j) acc[r][j] = 0.f;

  const int taps = g.KH * g.KW;
  for (int tap = 0; tap < taps; ++tap) {
    const int ky = tap / g.KW, kx = tap - ky * g.KW;
    int64_t soff[4];
    bool sval[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      sval[i] = false;
      soff[i] = 0;
      if (lb[i] >= 0) {
        int ys, xs;
        if (conv_src(g, ly[i] * g.stride + ky * g.dil - g.pad_t, lx[i] * g.stride + kx * g.dil - g.pad_l, ys, xs)) {
          sval[i] = true;
          soff[i] = (int64_t)lb[i] * g.x_bstride +
                    ((p.x_layout == LNS_NCHW) ? ((int64_t)ys * g.Win + xs) : ((int64_t)ys * g.Win + xs) * g.Cin);
        }
      }
    }
    for (int c0 = 0; c0 < g.Cin; c0 += BK) {
      const int c = c0 + k_l;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        float v = 0.f;
        if (sval[i] && c < g.Cin) {
          v = ld_as_float(p.x, p.x_dtype, soff[i] + (int64_t)c * chan_stride);
          if (p.pro_scale) {
            int64_t sc = (int64_t)lb[i] * g.Cin + c;
            v = fmaf(v, __ldg(p.pro_scale + sc), __ldg(p.pro_shift + sc));
          }
          v = apply_act(v, p.pro_act);
        }
        As[k_l][(tid >> 4) + 16 * i] = v;
      }
      for (int e = tid; e < BK * BN; e += 256) {
        int k = e / BN, n = e - k * BN;
        int cc = c0 + k;
        float v = 0.f;
        if (cc < g.Cin && n0 + n < g.Cout) v = __ldg(p.w + ((int64_t)tap * g.Cin + cc) * g.Cout + n0 + n);
        Bs[k][n] = v;
      }
      __syncthreads();
#pragma unroll
      for (int k = 0; k < BK; ++k) {
        float4 a4 = *reinterpret_cast<const float4*>(&As[k][ty * 4]);
        float a[4] = {a4.x, a4.y, a4.z, a4.w};
        float b[TN];
#pragma unroll
        for (int j = 0; j < TN; ++j) b[j] = Bs[k][tx * TN + j];
#pragma unroll
        for (int r = 0; r < 4; ++r)
#pragma unroll
          for (int j = 0; j < TN; ++j) acc[r][j] = fmaf(a[r], b[j], acc[r][j]);
      }
      __syncthreads();
    }
  }

  // epilogue
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    int m = m0 + ty * 4 + r;
    if (m >= p.M) continue;
    int b = m / HWo;
    int pix = m - b * HWo;
#pragma unroll
    for (int j = 0; j < TN; ++j) {
      int n = n0 + tx * TN + j;
      if (n >= g.Cout) continue;
      float v = acc[r][j];
      if (p.bias) v += __ldg(p.bias + n);
      if (p.sample_bias) v += __ldg(p.sample_bias + (int64_t)b * g.Cout + n);
      if (p.pre_add) v += ld_as_float(p.pre_add, p.pre_add_dtype, (int64_t)b * p.pre_add_bstride + (int64_t)pix * g.Cout + n);
      v = apply_act(v, p.act);
      if (p.residual) v += ld_as_float(p.residual, p.res_dtype, (int64_t)b * p.res_bstride + (int64_t)pix * g.Cout + n);
      int64_t o = (int64_t)b * g.y_bstride + ((p.y_layout == LNS_NCHW) ? ((int64_t)n * HWo + pix) : ((int64_t)pix * g.Cout + n));
      st_from_float(p.y, p.y_dtype, o, v);
    }
  }
}

// ---- channel lift: 1x1 conv from a few channels (Cin <= 16) to many (Cout % 8 == 0) -----------------------------------
// The encoder's first layer (1|3|4 -> 64 channels at full resolution, reading the reference's NCHW fp32 input) and the
// propagator's in_proj (16 -> 128 on the fp32 latent) are pure HBM streams: K is too small for any GEMM tiling (the generic
// 64 x BN tile above spent 1.7 ms on a layer whose output takes 83 us to write).  Thread = (pixel, 8 output channels): the
// pixel's Cin values are read once (a broadcast within the pixel's thread group), the filter sits in shared memory, every
// thread writes one 16-byte (16-bit output) or 32-byte chunk, a pixel's channels are contiguous across its threads.
// Same accumulation order as conv_simt_kernel (c ascending, then bias): bit-identical on the fp32 path.
__global__ void __launch_bounds__(256) lift1x1_kernel(const SimtParams p, int pix_per_block) {
  extern __shared__ float wl_s[];  // [Cin][Cout] + bias [Cout]
  const ConvGeom& g = p.g;
  const int Cin = g.Cin, Cout = g.Cout;
  for (int e = threadIdx.x; e < Cin * Cout; e += 256) wl_s[e] = __ldg(p.w + e);
  float* bias_s = wl_s + Cin * Cout;
  for (int e = threadIdx.x; e < Cout; e += 256) bias_s[e] = p.bias ? __ldg(p.bias + e) : 0.f;
  __syncthreads();
  const int CG = Cout >> 3, ppp = 256 / CG;  // channel groups per pixel, pixels per pass
  const int cg = threadIdx.x % CG, pl = threadIdx.x / CG;
  const int HW = g.Hout * g.Wout;
  const int64_t m_end = min((int64_t)p.M, (int64_t)(blockIdx.x + 1) * pix_per_block);
  const bool fast = is_h16(p.y_dtype);
  if (pl >= ppp) return;
  for (int64_t m = (int64_t)blockIdx.x * pix_per_block + pl; m < m_end; m += ppp) {
    const int b = (int)(m / HW), pix = (int)(m - (int64_t)b * HW);
    float xv[16];
    if (p.x_layout == LNS_NCHW) {
#pragma unroll
      for (int c = 0; c < 16; ++c)
        if (c < Cin) xv[c] = __ldg(reinterpret_cast<const float*>(p.x) + (int64_t)b * g.x_bstride + (int64_t)c * HW + pix);
    } else {
#pragma unroll
      for (int c = 0; c < 16; c += 4)
        if (c < Cin) {
          const float4 t4 = ld4_as_float(p.x, p.x_dtype, (int64_t)b * g.x_bstride + (int64_t)pix * Cin + c);
          xv[c] = t4.x; xv[c + 1] = t4.y; xv[c + 2] = t4.z; xv[c + 3] = t4.w;
        }
    }
    float acc[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] = 0.f;
#pragma unroll
    for (int c = 0; c < 16; ++c) {
      if (c < Cin) {
        const float4 w0 = *reinterpret_cast<const float4*>(wl_s + c * Cout + cg * 8);
        const float4 w1 = *reinterpret_cast<const float4*>(wl_s + c * Cout + cg * 8 + 4);
        acc[0] = fmaf(xv[c], w0.x, acc[0]); acc[1] = fmaf(xv[c], w0.y, acc[1]);
        acc[2] = fmaf(xv[c], w0.z, acc[2]); acc[3] = fmaf(xv[c], w0.w, acc[3]);
        acc[4] = fmaf(xv[c], w1.x, acc[4]); acc[5] = fmaf(xv[c], w1.y, acc[5]);
        acc[6] = fmaf(xv[c], w1.z, acc[6]); acc[7] = fmaf(xv[c], w1.w, acc[7]);
      }
    }
    const int64_t o = (int64_t)b * g.y_bstride + (int64_t)pix * Cout + cg * 8;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      float v = acc[j] + bias_s[cg * 8 + j];
      if (p.sample_bias) v += __ldg(p.sample_bias + (int64_t)b * Cout + cg * 8 + j);
      acc[j] = fast ? apply_act_fast(v, p.act) : apply_act(v, p.act);
    }
    if (p.residual) {
      const int64_t ro = (int64_t)b * p.res_bstride + (int64_t)pix * Cout + cg * 8;
      const float4 r0 = ld4_as_float(p.residual, p.res_dtype, ro), r1 = ld4_as_float(p.residual, p.res_dtype, ro + 4);
      acc[0] += r0.x; acc[1] += r0.y; acc[2] += r0.z; acc[3] += r0.w;
      acc[4] += r1.x; acc[5] += r1.y; acc[6] += r1.z; acc[7] += r1.w;
    }
    st4_from_float(p.y, p.y_dtype, o, make_float4(acc[0], acc[1], acc[2], acc[3]));
    st4_from_float(p.y, p.y_dtype, o + 4, make_float4(acc[4], acc[5], acc[6], acc[7]));
  }
}

// v2 of the channel lift for Cout % 4 == 0 (the encoder's NCHW fp32 input lift 1|3|4 -> 64, the propagator's in_proj 16 -> 128
// and the decoder's 16 -> 128 on the fp32 NHWC latent).  The ncu profile of the kernel above (18.1 M warp instructions, 65 % of the stall
// samples on the shared-memory scoreboard for 75 776 pixels) showed it re-loading its 16 x 8 filter values from shared memory
// for EVERY pixel -- one FMA per loaded value -- and paying one global round trip per pixel pass.  Here a thread owns 4
// output channels for the whole kernel and keeps their Cin x 4 filter values (and bias) in registers; the CTA's 128 input
// pixels arrive once, coalesced, through shared memory (one round trip per CTA; the per-pixel reads are warp broadcasts); a
// pixel's Cout / 4 threads write its channels as one contiguous run.
// Same accumulation order as conv_simt_kernel (c ascending, then bias): bit-identical on the fp32 path.
constexpr int kLiftPix = 128;  // pixels per CTA
// Q4 = input channels / 4, rounded up (compile time: the filter registers of absent channels do not exist -- 1 | 3 | 4 input
// channels need 16 instead of 64 of them, which more than doubles the resident CTAs of the encoder's full-resolution lift)
template <int Q4>
__global__ void __launch_bounds__(256) lift1x1_v2_kernel(const SimtParams p) {
  __shared__ __align__(16) float xs[kLiftPix * 16];
  const ConvGeom& g = p.g;
  const int Cin = g.Cin, Cout = g.Cout;
  const int HW = g.Hout * g.Wout;
  const int64_t m0 = (int64_t)blockIdx.x * kLiftPix;
  const int npx = (int)min((int64_t)kLiftPix, (int64_t)p.M - m0);
  constexpr int q4 = Q4;
  if (p.x_layout == LNS_NCHW) {
    // the reference's NCHW fp32 input (encoder lift, Cin = 1 | 3 | 4): consecutive threads -> consecutive pixels of one plane;
    // channels up to the next multiple of 4 are zero (their filter registers are zero too: fma(0, 0, a) == a)
    for (int e = threadIdx.x; e < npx * q4 * 4; e += 256) {
      const int c = e / npx, pi = e - c * npx;
      const int64_t m = m0 + pi;
      const int b = (int)(m / HW), pix = (int)(m - (int64_t)b * HW);
      xs[pi * 16 + c] = c < Cin ? __ldg(reinterpret_cast<const float*>(p.x) + (int64_t)b * g.x_bstride + (int64_t)c * HW + pix) : 0.f;
    }
  } else {
    // NHWC: Cin / 4 float4 per pixel, consecutive threads -> consecutive 16-byte pieces
    for (int e = threadIdx.x; e < npx * q4; e += 256) {
      const int pi = e / q4, c = (e - pi * q4) * 4;
      const int64_t m = m0 + pi;
      const int b = (int)(m / HW), pix = (int)(m - (int64_t)b * HW);
      const float4 t4 = ld4_as_float(p.x, p.x_dtype, (int64_t)b * g.x_bstride + (int64_t)pix * Cin + c);
      *reinterpret_cast<float4*>(xs + pi * 16 + c) = t4;
    }
  }
  const int TPP = Cout >> 2, ppp = 256 / TPP;  // threads per pixel, pixels per pass
  const int cq = threadIdx.x % TPP, pl = threadIdx.x / TPP;
  float4 wr[Q4 * 4];
#pragma unroll
  for (int c = 0; c < Q4 * 4; ++c) wr[c] = (c < Cin) ? __ldg(reinterpret_cast<const float4*>(p.w + c * Cout + cq * 4)) : make_float4(0.f, 0.f, 0.f, 0.f);
  const float4 bv = p.bias ? __ldg(reinterpret_cast<const float4*>(p.bias + cq * 4)) : make_float4(0.f, 0.f, 0.f, 0.f);
  const bool fast = is_h16(p.y_dtype);
  __syncthreads();
  if (pl >= ppp) return;
  // (sample, pixel) of this thread's first pixel, then advanced incrementally: no division in the loop
  int b = (int)((m0 + pl) / HW), pix = (int)((m0 + pl) - (int64_t)b * HW);
  const int adv_b = ppp / HW, adv_p = ppp - adv_b * HW;
  for (int pi = pl; pi < npx; pi += ppp) {
    float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
#pragma unroll
    for (int c4 = 0; c4 < Q4 * 4; c4 += 4) {
      {
        const float4 xv = *reinterpret_cast<const float4*>(xs + pi * 16 + c4);
        const float xa[4] = {xv.x, xv.y, xv.z, xv.w};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          a0 = fmaf(xa[j], wr[c4 + j].x, a0); a1 = fmaf(xa[j], wr[c4 + j].y, a1);
          a2 = fmaf(xa[j], wr[c4 + j].z, a2); a3 = fmaf(xa[j], wr[c4 + j].w, a3);
        }
      }
    }
    float v[4] = {a0 + bv.x, a1 + bv.y, a2 + bv.z, a3 + bv.w};
    if (p.sample_bias) {
      const float4 sb = __ldg(reinterpret_cast<const float4*>(p.sample_bias + (int64_t)b * Cout + cq * 4));
      v[0] += sb.x; v[1] += sb.y; v[2] += sb.z; v[3] += sb.w;
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) v[j] = fast ? apply_act_fast(v[j], p.act) : apply_act(v[j], p.act);
    if (p.residual) {
      const float4 r0 = ld4_as_float(p.residual, p.res_dtype, (int64_t)b * p.res_bstride + (int64_t)pix * Cout + cq * 4);
      v[0] += r0.x; v[1] += r0.y; v[2] += r0.z; v[3] += r0.w;
    }
    st4_from_float(p.y, p.y_dtype, (int64_t)b * g.y_bstride + (int64_t)pix * Cout + cq * 4, make_float4(v[0], v[1], v[2], v[3]));
    b += adv_b;
    pix += adv_p;
    if (pix >= HW) {
      pix -= HW;
      ++b;
    }
  }
}

static bool lift_ok(const LnsConvDesc* d) {
  const bool nhwc_in = d->x_layout == LNS_NHWC;
  return d->KH == 1 && d->KW == 1 && d->stride == 1 && d->pad_t == 0 && d->pad_l == 0 && d->Hv == d->Hin && d->Wv == d->Win &&
         d->Hout == d->Hin && d->Wout == d->Win && d->Cin <= 16 && d->Cout % 8 == 0 && d->Cout >= 8 && d->Cout <= 256 &&
         d->y_layout == LNS_NHWC && d->pro_scale == nullptr && d->pro_act == LNS_ACT_NONE && d->pre_add == nullptr &&
         d->y_bstride % 4 == 0 && (d->residual == nullptr || d->res_bstride % 4 == 0) &&
         (nhwc_in ? (d->Cin % 4 == 0 && d->x_bstride % 4 == 0) : !is_h16_host(d->x_dtype));
}

int conv2d_simt(const LnsConvDesc* d, cudaStream_t stream) {
  LNS_REQUIRE(d->w_format == LNS_W_SIMT_F32, "lns_conv2d(simt): weights must be packed as LNS_W_SIMT_F32");
  SimtParams p;
  p.g = make_geom(d);
  p.x = d->x; p.x_dtype = d->x_dtype; p.x_layout = d->x_layout;
  p.w = reinterpret_cast<const float*>(d->w);
  p.bias = d->bias; p.sample_bias = d->sample_bias;
  p.pro_scale = d->pro_scale; p.pro_shift = d->pro_shift; p.pro_act = d->pro_act;
  p.act = d->act;
  p.pre_add = d->pre_add; p.pre_add_dtype = d->pre_add_dtype; p.pre_add_bstride = d->pre_add_bstride;
  p.residual = d->residual; p.res_dtype = d->res_dtype; p.res_bstride = d->res_bstride;
  p.y = d->y; p.y_dtype = d->y_dtype; p.y_layout = d->y_layout;
  int64_t M = (int64_t)d->B * d->Hout * d->Wout;
  LNS_REQUIRE(M < (1ll << 31), "lns_conv2d(simt): too many output pixels");
  LNS_REQUIRE(!(d->pro_scale && !d->pro_shift), "lns_conv2d: pro_scale without pro_shift");
  p.M = (int)M;
  if (lift_ok(d)) {
    const int ppp = 256 / (d->Cout / 8);
    static int passes = 0;
    if (!passes) {
      const char* c = getenv("LNS_LIFT_PASSES");
      passes = c ? atoi(c) : 8;
      if (passes < 1) passes = 1;
    }
    int ppb = ppp * passes;                                // 8 passes per block ...
    while ((M + ppb - 1) / ppb > 148 * 16) ppb *= 2;       // ... more when the grid would exceed 16 blocks per SM
    const int tpp = d->Cout / 4;
    if (d->Cout % 4 == 0 && tpp <= 256 && 256 % tpp == 0 && (reinterpret_cast<uintptr_t>(d->w) & 15) == 0 &&
        (!d->bias || (reinterpret_cast<uintptr_t>(d->bias) & 15) == 0) &&
        (!d->sample_bias || (reinterpret_cast<uintptr_t>(d->sample_bias) & 15) == 0)) {
      const unsigned nblk = (unsigned)((M + kLiftPix - 1) / kLiftPix);
      switch ((d->Cin + 3) / 4) {
        case 1: lift1x1_v2_kernel<1><<<nblk, 256, 0, stream>>>(p); break;
        case 2: lift1x1_v2_kernel<2><<<nblk, 256, 0, stream>>>(p); break;
        case 3: lift1x1_v2_kernel<3><<<nblk, 256, 0, stream>>>(p); break;
        default: lift1x1_v2_kernel<4><<<nblk, 256, 0, stream>>>(p); break;
      }
      return check_launch("lift1x1_v2_kernel");
    }
    const size_t smem = ((size_t)d->Cin * d->Cout + d->Cout) * sizeof(float);
    lift1x1_kernel<<<(unsigned)((M + ppb - 1) / ppb), 256, smem, stream>>>(p, ppb);
    return check_launch("lift1x1_kernel");
  }
  if (d->Cout <= 16) {
    dim3 grid(cdiv(M, 64), cdiv(d->Cout, 16));
    conv_simt_kernel<16><<<grid, 256, 0, stream>>>(p);
  } else {
    dim3 grid(cdiv(M, 64), cdiv(d->Cout, 64));
    conv_simt_kernel<64><<<grid, 256, 0, stream>>>(p);
  }
  return check_launch("conv_simt_kernel");
}

// ---- filter re-layout ------------------------------------------------------------------------------
__global__ void pack_simt_kernel(const float* __restrict__ w, int Cout, int Cin, int KH, int KW, float* __restrict__ out) {
  int64_t total = (int64_t)Cout * Cin * KH * KW;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    int n = (int)(i % Cout);
    int64_t r = i / Cout;
    int c = (int)(r % Cin);
    int tap = (int)(r / Cin);
    out[i] = w[((int64_t)n * Cin + c) * KH * KW + tap];
  }
}

// [tap][Cin/64][Cout][64] bf16; inside each 128-byte row the eight 16-byte chunks are XOR-swizzled with (row & 7),
// i.e. the image is byte-for-byte what tcgen05's SWIZZLE_128B K-major shared-memory layout expects.
// (f16 != 0: IEEE-half elements instead of bf16 -- LNS_W_UMMA_F16)
// (f16 == 2: LNS_W_UMMA_F16X2 -- the hi image followed by the image of lo = rn_f16(w - hi))
__global__ void pack_umma_kernel(const float* __restrict__ w, int Cout, int Cin, int KH, int KW, int f16,
                                 uint16_t* __restrict__ out) {
  int64_t total = (int64_t)Cout * Cin * KH * KW;
  int slabs = Cin / 64;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    int kk = (int)(i % 64);
    int64_t r = i / 64;
    int n = (int)(r % Cout);
    r /= Cout;
    int slab = (int)(r % slabs);
    int tap = (int)(r / slabs);
    int c = slab * 64 + kk;
    float v = w[((int64_t)n * Cin + c) * KH * KW + tap];
    int chunk = (kk >> 3) ^ (n & 7);
    int64_t o = (((int64_t)tap * slabs + slab) * Cout + n) * 64 + chunk * 8 + (kk & 7);
    out[o] = f16 ? to_h16<true>(v) : to_h16<false>(v);
    if (f16 == 2) out[total + o] = to_h16<true>(v - from_h16<true>(out[o]));
  }
}

// [tap][Cin/32][Cout][32] fp32 rounded to nearest TF32; 16-byte chunks (4 values) XOR-swizzled with (row & 7)
__global__ void pack_umma_tf32_kernel(const float* __restrict__ w, int Cout, int Cin, int KH, int KW, float* __restrict__ out) {
  int64_t total = (int64_t)Cout * Cin * KH * KW;
  int slabs = Cin / 32;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    int kk = (int)(i % 32);
    int64_t r = i / 32;
    int n = (int)(r % Cout);
    r /= Cout;
    int slab = (int)(r % slabs);
    int tap = (int)(r / slabs);
    int c = slab * 32 + kk;
    float v = round_tf32(w[((int64_t)n * Cin + c) * KH * KW + tap]);
    int chunk = (kk >> 2) ^ (n & 7);
    int64_t o = (((int64_t)tap * slabs + slab) * Cout + n) * 32 + chunk * 4 + (kk & 3);
    out[o] = v;
  }
}

}  // namespace lns

extern "C" {

int64_t lns_packed_weight_bytes(int Cout, int Cin, int KH, int KW, int format) {
  int64_t n = (int64_t)Cout * Cin * KH * KW;
  if (format == LNS_W_SIMT_F32) return n * 4;
  if (format == LNS_W_UMMA_BF16 || format == LNS_W_UMMA_F16) return (Cin % 64 == 0 && Cout % 16 == 0) ? n * 2 : -1;
  if (format == LNS_W_UMMA_TF32) return (Cin % 32 == 0 && Cout % 16 == 0) ? n * 4 : -1;
  if (format == LNS_W_UMMA_F16X2) return (Cin % 64 == 0 && Cout % 16 == 0) ? n * 4 : -1;
  return -1;
}

int lns_pack_conv_weight(const float* w, int Cout, int Cin, int KH, int KW, int format, void* out, void* stream) {
  LNS_REQUIRE(w && out && Cout > 0 && Cin > 0 && KH > 0 && KW > 0, "lns_pack_conv_weight: bad arguments");
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  int64_t total = (int64_t)Cout * Cin * KH * KW;
  int blocks = (int)((total + 255) / 256);
  if (blocks > 4096) blocks = 4096;
  if (format == LNS_W_SIMT_F32) {
    lns::pack_simt_kernel<<<blocks, 256, 0, s>>>(w, Cout, Cin, KH, KW, reinterpret_cast<float*>(out));
  } else if (format == LNS_W_UMMA_BF16 || format == LNS_W_UMMA_F16 || format == LNS_W_UMMA_F16X2) {
    LNS_REQUIRE(Cin % 64 == 0 && Cout % 16 == 0, "lns_pack_conv_weight: UMMA format needs Cin%%64==0, Cout%%16==0 (got %d,%d)",
                Cin, Cout);
    lns::pack_umma_kernel<<<blocks, 256, 0, s>>>(w, Cout, Cin, KH, KW,
                                                 format == LNS_W_UMMA_F16X2 ? 2 : (format == LNS_W_UMMA_F16 ? 1 : 0),
                                                 reinterpret_cast<uint16_t*>(out));
  } else if (format == LNS_W_UMMA_TF32) {
    LNS_REQUIRE(Cin % 32 == 0 && Cout % 16 == 0, "lns_pack_conv_weight: TF32 format needs Cin%%32==0, Cout%%16==0 (got %d,%d)",
                Cin, Cout);
    lns::pack_umma_tf32_kernel<<<blocks, 256, 0, s>>>(w, Cout, Cin, KH, KW, reinterpret_cast<float*>(out));
  } else {
    lns::set_error("lns_pack_conv_weight: unknown format %d", format);
    return LNS_E_INVALID;
  }
  return lns::check_launch("pack kernel");
}

}  // extern "C"
