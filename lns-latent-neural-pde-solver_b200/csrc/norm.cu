// Normalisation kernels: per-(sample,channel) statistics, GroupNorm/InstanceNorm finalize -> per-(sample,channel)
// affine, standalone affine+activation, LayerNorm(+pe).  All HBM-bound: 8/16-byte vector accesses, coalesced over the
// contiguous channel axis of NHWC, warp-shuffle / fixed-order shared-memory reductions (deterministic, and the
// result for one sample does not depend on the batch it is in).
#include "common.cuh"

namespace lns {

constexpr int kStatsChunkPixels = 1024;

// grid (nchunk, B), block 256.  thread -> (channel quad q, pixel lane); C % 4 == 0, C <= 1024.
__global__ void __launch_bounds__(256) chan_stats_kernel(const void* __restrict__ x, int dtype, int HW, int C,
                                                          int64_t bstride, float2* __restrict__ partial, int nchunk) {
  extern __shared__ float red[];  // [rows][C][2]
  const int cg = C >> 2;
  const int rows = 256 / cg;
  const int q = threadIdx.x % cg, lane = threadIdx.x / cg;
  const int chunk = blockIdx.x, b = blockIdx.y;
  const int p0 = chunk * kStatsChunkPixels;
  const int p1 = min(HW, p0 + kStatsChunkPixels);
  float s[4] = {0, 0, 0, 0}, ss[4] = {0, 0, 0, 0};
  if (lane < rows) {
    // 8 independent loads in flight per thread, all issued before the first use (accumulation order unchanged)
    for (int pb = p0 + lane; pb < p1; pb += 8 * rows) {
      int64_t off[8];
      bool ok[8];
      float4 v[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const int p = pb + u * rows;
        ok[u] = p < p1;
        off[u] = (int64_t)b * bstride + (int64_t)(ok[u] ? p : pb) * C + q * 4;
      }
      ld4n_as_float<8>(x, dtype, off, ok, v);
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        s[0] += v[u].x; ss[0] = fmaf(v[u].x, v[u].x, ss[0]);
        s[1] += v[u].y; ss[1] = fmaf(v[u].y, v[u].y, ss[1]);
        s[2] += v[u].z; ss[2] = fmaf(v[u].z, v[u].z, ss[2]);
        s[3] += v[u].w; ss[3] = fmaf(v[u].w, v[u].w, ss[3]);
      }
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      red[((lane * C) + q * 4 + j) * 2 + 0] = s[j];
      red[((lane * C) + q * 4 + j) * 2 + 1] = ss[j];
    }
  }
  __syncthreads();
  for (int c = threadIdx.x; c < C; c += 256) {
    double a = 0.0, a2 = 0.0;
    for (int l = 0; l < rows; ++l) {
      a += (double)red[((l * C) + c) * 2 + 0];
      a2 += (double)red[((l * C) + c) * 2 + 1];
    }
    partial[((int64_t)b * nchunk + chunk) * C + c] = make_float2((float)a, (float)a2);
  }
}

// grid B, block 256.  Pass 1: thread per channel sums its chunks (coalesced float2 loads, fixed order, fp64).  Pass 2: one thread
// per group combines the group's channels in channel order (fp64) into mean / rstd.  Pass 3: thread per channel writes the
// folded affine.  (The first version gave a whole warp to each group: with 4 channels per group 28 of 32 lanes idled through
// dependent loads and fp64 shuffles -- 52 us for 4736 samples; this one is bandwidth-trivial.)
template <bool PIV>
__global__ void __launch_bounds__(256) norm_finalize_kernel(const float2* __restrict__ partial, int nchunk, int C, int HW,
                                                             int G, float eps, const float* __restrict__ gamma,
                                                             const float* __restrict__ beta,
                                                             const float* __restrict__ prescale,
                                                             float* __restrict__ scale, float* __restrict__ shift) {
  extern __shared__ double fin_sm[];  // [C][2] channel sums | [G][2] mean, rstd
  double* csum = fin_sm;
  double* gst = fin_sm + 2 * C;
  const int b = blockIdx.x;
  const int cpg = C / G;
  for (int c = threadIdx.x; c < C; c += 256) {
    double cs = 0.0, css = 0.0;
    for (int k = 0; k < nchunk; ++k) {
      if (PIV) {
        // centred partials (sum d, sum d^2, pivot p, count n), x = p + d: exact recombination in fp64
        const float4 v = reinterpret_cast<const float4*>(partial)[((int64_t)b * nchunk + k) * C + c];
        const double sd = (double)v.x, pp = (double)v.z, n = (double)v.w;
        cs += sd + n * pp;
        css += (double)v.y + 2.0 * pp * sd + n * pp * pp;
      } else {
        const float2 v = partial[((int64_t)b * nchunk + k) * C + c];
        cs += (double)v.x;
        css += (double)v.y;
      }
    }
    const double ps = prescale ? (double)prescale[(int64_t)b * C + c] : 1.0;
    csum[2 * c] = ps * cs;
    csum[2 * c + 1] = ps * ps * css;
  }
  __syncthreads();
  for (int gi = threadIdx.x; gi < G; gi += 256) {
    double sum = 0.0, sumsq = 0.0;
    for (int j = 0; j < cpg; ++j) {
      sum += csum[2 * (gi * cpg + j)];
      sumsq += csum[2 * (gi * cpg + j) + 1];
    }
    const double n = (double)cpg * (double)HW;
    const double mean = sum / n;
    double var = sumsq / n - mean * mean;  // fp64: the cancellation costs E[x^2] / var * 2^-53
    if (var < 0.0) var = 0.0;
    gst[2 * gi] = mean;
    gst[2 * gi + 1] = 1.0 / sqrt(var + (double)eps);
  }
  __syncthreads();
  for (int c = threadIdx.x; c < C; c += 256) {
    const int gi = c / cpg;
    const double mean = gst[2 * gi], rstd = gst[2 * gi + 1];
    const double ga = gamma ? (double)gamma[c] : 1.0;
    const double be = beta ? (double)beta[c] : 0.0;
    const double ps = prescale ? (double)prescale[(int64_t)b * C + c] : 1.0;
    scale[(int64_t)b * C + c] = (float)(ps * rstd * ga);
    shift[(int64_t)b * C + c] = (float)(be - mean * rstd * ga);
  }
}

// elementwise y = act(x*scale[b][c] + shift[b][c]); 4 channels per thread
__global__ void __launch_bounds__(256) affine_act_kernel(const void* __restrict__ x, int x_dtype, int64_t x_bstride,
                                                          int64_t per_sample4, int C, const float* __restrict__ scale,
                                                          const float* __restrict__ shift, int act, void* __restrict__ y,
                                                          int y_dtype, int64_t y_bstride, int64_t total4) {
  const int c4n = C >> 2;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total4; i += (int64_t)gridDim.x * blockDim.x) {
    int64_t b = i / per_sample4;
    int64_t r = i - b * per_sample4;
    int c = (int)(r % c4n) * 4;
    float4 v = ld4_as_float(x, x_dtype, b * x_bstride + r * 4);
    if (scale) {
      float4 sc = __ldg(reinterpret_cast<const float4*>(scale + b * C + c));
      float4 sh = __ldg(reinterpret_cast<const float4*>(shift + b * C + c));
      v.x = fmaf(v.x, sc.x, sh.x); v.y = fmaf(v.y, sc.y, sh.y);
      v.z = fmaf(v.z, sc.z, sh.z); v.w = fmaf(v.w, sc.w, sh.w);
    }
    v.x = apply_act_for(v.x, act, y_dtype); v.y = apply_act_for(v.y, act, y_dtype);
    v.z = apply_act_for(v.z, act, y_dtype); v.w = apply_act_for(v.w, act, y_dtype);
    st4_from_float(y, y_dtype, b * y_bstride + r * 4, v);
  }
}

// Same op for the common case 256 % (C/4) == 0 (every thread keeps ONE channel quad for the whole sample: the per-sample
// affine is loaded once) with U independent loads in flight per thread and 32-bit index arithmetic.  The grid-stride version
// above has one 8-byte load in flight per thread and two 64-bit divisions per element: 2.7 TB/s in the round-1 timeline.
// grid (ceil(per_sample4 / (256*U)), min(B, 65535)), block 256.
template <int U>
__global__ void __launch_bounds__(256) affine_act2_kernel(const void* __restrict__ x, int x_dtype, int64_t x_bstride, int per_sample4,
                                                           int C, int B, const float* __restrict__ scale, const float* __restrict__ shift,
                                                           int act, void* __restrict__ y, int y_dtype, int64_t y_bstride, int tanh_silu) {
  const int c = (threadIdx.x % (C >> 2)) * 4;
  const int r0 = blockIdx.x * (256 * U) + threadIdx.x;
  const bool fast = is_h16(y_dtype);
  // Swish on 16-bit outputs as h + h*tanh(h), h = x/2: ONE special-function op per element instead of two (ex2 + rcp).  At 16 SFU
  // results per clock per SM the two-op form cost 0.13 ms of a 0.27 ms launch at 4736 x 32x32 x 64 (265 -> 221 us with tanh); the
  // approximation error (2^-11 relative on tanh) is of the size of the 16-bit storage rounding that follows: per-stage parity
  // moves by < 1 % of its value, the 20-step drift not at all.
  const bool ts = fast && tanh_silu && act == LNS_ACT_SILU;
  const float hs = ts ? 0.5f : 1.f;
  for (int b = blockIdx.y; b < B; b += gridDim.y) {
    float4 sc = make_float4(1.f, 1.f, 1.f, 1.f), sh = make_float4(0.f, 0.f, 0.f, 0.f);
    if (scale) {
      sc = __ldg(reinterpret_cast<const float4*>(scale + (int64_t)b * C + c));
      sh = __ldg(reinterpret_cast<const float4*>(shift + (int64_t)b * C + c));
    }
    int64_t off[U];
    bool ok[U];
    float4 v[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int r = r0 + u * 256;
      ok[u] = r < per_sample4;
      off[u] = (int64_t)b * x_bstride + (int64_t)(ok[u] ? r : 0) * 4;
    }
    ld4n_as_float<U>(x, x_dtype, off, ok, v);
#pragma unroll
    for (int u = 0; u < U; ++u) {
      float4 t = v[u];
      t.x = fmaf(t.x, sc.x, sh.x); t.y = fmaf(t.y, sc.y, sh.y); t.z = fmaf(t.z, sc.z, sh.z); t.w = fmaf(t.w, sc.w, sh.w);
      if (ts) {
        float th;
        t.x *= hs; asm("tanh.approx.f32 %0, %1;" : "=f"(th) : "f"(t.x)); t.x = fmaf(t.x, th, t.x);
        t.y *= hs; asm("tanh.approx.f32 %0, %1;" : "=f"(th) : "f"(t.y)); t.y = fmaf(t.y, th, t.y);
        t.z *= hs; asm("tanh.approx.f32 %0, %1;" : "=f"(th) : "f"(t.z)); t.z = fmaf(t.z, th, t.z);
        t.w *= hs; asm("tanh.approx.f32 %0, %1;" : "=f"(th) : "f"(t.w)); t.w = fmaf(t.w, th, t.w);
      } else if (fast) {
        t.x = apply_act_fast(t.x, act); t.y = apply_act_fast(t.y, act); t.z = apply_act_fast(t.z, act); t.w = apply_act_fast(t.w, act);
      } else {
        t.x = apply_act(t.x, act); t.y = apply_act(t.y, act); t.z = apply_act(t.z, act); t.w = apply_act(t.w, act);
      }
      if (ok[u]) st4_from_float(y, y_dtype, (int64_t)b * y_bstride + (int64_t)(r0 + u * 256) * 4, t);
    }
  }
}

// y = x * (1 + gate[b][c])
__global__ void __launch_bounds__(256) channel_gate_kernel(const void* __restrict__ x, int dtype, int64_t per_sample4,
                                                            int C, const float* __restrict__ gate,
                                                            void* __restrict__ y, int y_dtype, int64_t total4) {
  const int c4n = C >> 2;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total4; i += (int64_t)gridDim.x * blockDim.x) {
    int64_t b = i / per_sample4;
    int64_t r = i - b * per_sample4;
    int c = (int)(r % c4n) * 4;
    float4 v = ld4_as_float(x, dtype, i * 4);
    float4 gt = __ldg(reinterpret_cast<const float4*>(gate + b * C + c));
    v.x *= (1.f + gt.x); v.y *= (1.f + gt.y); v.z *= (1.f + gt.z); v.w *= (1.f + gt.w);
    st4_from_float(y, y_dtype, i * 4, v);
  }
}

// one warp per token row; C <= 1024, two-pass (mean, then centred variance) in registers
template <int MAXV>
__global__ void __launch_bounds__(256) layernorm_kernel(const void* __restrict__ x, int x_dtype, int64_t rows, int n, int C,
                                                         const float* __restrict__ gamma, const float* __restrict__ beta,
                                                         float eps, const float* __restrict__ pe, void* __restrict__ y,
                                                         int y_dtype) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t row = blockIdx.x * (int64_t)(blockDim.x >> 5) + warp;
  if (row >= rows) return;
  float v[MAXV];
  float s = 0.f;
#pragma unroll
  for (int j = 0; j < MAXV; ++j) {
    int c = lane + 32 * j;
    v[j] = (c < C) ? ld_as_float(x, x_dtype, row * C + c) : 0.f;
    s += v[j];
  }
  const float mean = warp_sum(s) / (float)C;
  float s2 = 0.f;
#pragma unroll
  for (int j = 0; j < MAXV; ++j) {
    int c = lane + 32 * j;
    float dlt = (c < C) ? (v[j] - mean) : 0.f;
    s2 = fmaf(dlt, dlt, s2);
  }
  const float rstd = rsqrtf(warp_sum(s2) / (float)C + eps);
  const int tok = (int)(row % n);
#pragma unroll
  for (int j = 0; j < MAXV; ++j) {
    int c = lane + 32 * j;
    if (c < C) {
      float o = (v[j] - mean) * rstd;
      o = o * (gamma ? __ldg(gamma + c) : 1.f) + (beta ? __ldg(beta + c) : 0.f);
      if (pe) o += __ldg(pe + (int64_t)tok * C + c);
      st_from_float(y, y_dtype, row * C + c, o);
    }
  }
}

}  // namespace lns

extern "C" {

int lns_chan_stats_chunks(int H, int W) { return lns::cdiv((int64_t)H * W, lns::kStatsChunkPixels); }

int lns_chan_stats(const void* x, int dtype, int B, int H, int W, int C, int64_t bstride, float* partial, void* stream) {
  LNS_REQUIRE(x && partial && B > 0 && H > 0 && W > 0, "lns_chan_stats: bad arguments");
  int cg = C / 4;
  LNS_REQUIRE(C % 4 == 0 && C >= 4 && C <= 1024 && (cg & (cg - 1)) == 0,
              "lns_chan_stats: C must be a power of two in [4,1024] (got %d)", C);
  int nchunk = lns_chan_stats_chunks(H, W);
  int rows = 256 / cg;
  size_t smem = (size_t)rows * C * 2 * sizeof(float);
  dim3 grid(nchunk, B);
  lns::chan_stats_kernel<<<grid, 256, smem, reinterpret_cast<cudaStream_t>(stream)>>>(
      x, dtype, H * W, C, bstride, reinterpret_cast<float2*>(partial), nchunk);
  return lns::check_launch("chan_stats_kernel");
}

int lns_norm_finalize(const float* partial, int B, int nchunk, int C, int HW, int G, float eps, const float* gamma,
                      const float* beta, const float* prescale, float* scale, float* shift, void* stream) {
  LNS_REQUIRE(partial && scale && shift && B > 0 && C > 0 && G > 0 && C % G == 0, "lns_norm_finalize: bad arguments");
  LNS_REQUIRE((size_t)(2 * C + 2 * G) * sizeof(double) <= 48 * 1024, "lns_norm_finalize: C = %d too large", C);
  lns::norm_finalize_kernel<false><<<B, 256, (size_t)(2 * C + 2 * G) * sizeof(double), reinterpret_cast<cudaStream_t>(stream)>>>(
      reinterpret_cast<const float2*>(partial), nchunk, C, HW, G, eps, gamma, beta, prescale, scale, shift);
  return lns::check_launch("norm_finalize_kernel");
}

int lns_norm_finalize_centred(const float* partial, int B, int nchunk, int C, int HW, int G, float eps, const float* gamma,
                              const float* beta, const float* prescale, float* scale, float* shift, void* stream) {
  LNS_REQUIRE(partial && scale && shift && B > 0 && C > 0 && G > 0 && C % G == 0, "lns_norm_finalize_centred: bad arguments");
  LNS_REQUIRE((size_t)(2 * C + 2 * G) * sizeof(double) <= 48 * 1024, "lns_norm_finalize_centred: C = %d too large", C);
  LNS_REQUIRE((reinterpret_cast<uintptr_t>(partial) & 15) == 0, "lns_norm_finalize_centred: partials must be 16-byte aligned");
  lns::norm_finalize_kernel<true><<<B, 256, (size_t)(2 * C + 2 * G) * sizeof(double), reinterpret_cast<cudaStream_t>(stream)>>>(
      reinterpret_cast<const float2*>(partial), nchunk, C, HW, G, eps, gamma, beta, prescale, scale, shift);
  return lns::check_launch("norm_finalize_kernel");
}

int lns_affine_act(const void* x, int x_dtype, int64_t x_bstride, int B, int HW, int C, const float* scale,
                   const float* shift, int act, void* y, int y_dtype, int64_t y_bstride, void* stream) {
  LNS_REQUIRE(x && y && B > 0 && HW > 0 && C > 0 && C % 4 == 0, "lns_affine_act: bad arguments (C=%d)", C);
  LNS_REQUIRE(x_bstride % 4 == 0 && y_bstride % 4 == 0, "lns_affine_act: batch strides must be multiples of 4");
  LNS_REQUIRE(!(scale && !shift), "lns_affine_act: scale without shift");
  int64_t per4 = (int64_t)HW * C / 4;
  int64_t total4 = per4 * B;
  if (256 % (C / 4) == 0 && per4 < (1ll << 29)) {
    constexpr int U = 8;
    dim3 grid((unsigned)((per4 + 256 * U - 1) / (256 * U)), (unsigned)(B < 65535 ? B : 65535));
    lns::affine_act2_kernel<U><<<grid, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
        x, x_dtype, x_bstride, (int)per4, C, B, scale, shift, act, y, y_dtype, y_bstride, getenv("LNS_AFFINE_SFU2") ? 0 : 1);  // Swish on 16-bit outputs: one tanh.approx (default) or ex2 + rcp
    return lns::check_launch("affine_act2_kernel");
  }
  int blocks = (int)((total4 + 255) / 256);
  if (blocks > 148 * 16) blocks = 148 * 16;
  lns::affine_act_kernel<<<blocks, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      x, x_dtype, x_bstride, per4, C, scale, shift, act, y, y_dtype, y_bstride, total4);
  return lns::check_launch("affine_act_kernel");
}

int lns_channel_gate(const void* x, int dtype, int B, int HW, int C, const float* gate, void* y, int y_dtype,
                     void* stream) {
  LNS_REQUIRE(x && y && gate && B > 0 && HW > 0 && C % 4 == 0, "lns_channel_gate: bad arguments");
  int64_t per4 = (int64_t)HW * C / 4;
  int64_t total4 = per4 * B;
  int blocks = (int)((total4 + 255) / 256);
  if (blocks > 148 * 16) blocks = 148 * 16;
  lns::channel_gate_kernel<<<blocks, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(x, dtype, per4, C, gate, y,
                                                                                        y_dtype, total4);
  return lns::check_launch("channel_gate_kernel");
}

int lns_layernorm(const void* x, int x_dtype, int B, int n, int C, const float* gamma, const float* beta, float eps,
                  const float* pe, void* y, int y_dtype, void* stream) {
  LNS_REQUIRE(x && y && B > 0 && n > 0 && C > 0 && C <= 1024, "lns_layernorm: bad arguments (C=%d)", C);
  int64_t rows = (int64_t)B * n;
  int blocks = (int)((rows + 7) / 8);
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  if (C <= 128)
    lns::layernorm_kernel<4><<<blocks, 256, 0, s>>>(x, x_dtype, rows, n, C, gamma, beta, eps, pe, y, y_dtype);
  else
    lns::layernorm_kernel<32><<<blocks, 256, 0, s>>>(x, x_dtype, rows, n, C, gamma, beta, eps, pe, y, y_dtype);
  return lns::check_launch("layernorm_kernel");
}

}  // extern "C"
