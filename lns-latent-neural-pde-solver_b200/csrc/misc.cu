// Layout conversion at the two ends of the path (the reference's tensors are NCHW fp32) and small helpers.
#include "common.cuh"

namespace lns {

// 32x32 (channel x pixel) tile transpose through shared memory: both sides coalesced.
// grid (ceil(P/32), ceil(C/32), B), block (32, 8)
__global__ void nchw_to_nhwc_kernel(const float* __restrict__ x, int C, int P, int64_t x_bstride, void* __restrict__ y,
                                    int y_dtype, int64_t y_bstride) {
  __shared__ float tile[32][33];
  const int b = blockIdx.z, p0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
  for (int r = threadIdx.y; r < 32; r += 8) {
    int c = c0 + r, p = p0 + threadIdx.x;
    tile[r][threadIdx.x] = (c < C && p < P) ? __ldg(x + (int64_t)b * x_bstride + (int64_t)c * P + p) : 0.f;
  }
  __syncthreads();
  for (int r = threadIdx.y; r < 32; r += 8) {
    int p = p0 + r, c = c0 + threadIdx.x;
    if (p < P && c < C) st_from_float(y, y_dtype, (int64_t)b * y_bstride + (int64_t)p * C + c, tile[threadIdx.x][r]);
  }
}

__global__ void nhwc_to_nchw_kernel(const void* __restrict__ x, int x_dtype, int C, int P, int64_t x_bstride,
                                    float* __restrict__ y, int64_t y_bstride) {
  __shared__ float tile[32][33];
  const int b = blockIdx.z, p0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
  for (int r = threadIdx.y; r < 32; r += 8) {
    int p = p0 + r, c = c0 + threadIdx.x;
    tile[r][threadIdx.x] = (p < P && c < C) ? ld_as_float(x, x_dtype, (int64_t)b * x_bstride + (int64_t)p * C + c) : 0.f;
  }
  __syncthreads();
  for (int r = threadIdx.y; r < 32; r += 8) {
    int c = c0 + r, p = p0 + threadIdx.x;
    if (c < C && p < P) y[(int64_t)b * y_bstride + (int64_t)c * P + p] = tile[threadIdx.x][r];
  }
}

__global__ void fourier_embedding_kernel(const float* __restrict__ param, int B, int dim, float max_period,
                                         float* __restrict__ out) {
  const int half = dim / 2;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < B * dim; i += gridDim.x * blockDim.x) {
    int b = i / dim, j = i - b * dim;
    float v = 0.f;
    if (j < 2 * half) {
      int f = j < half ? j : j - half;
      float freq = expf(-logf(max_period) * (float)f / (float)half);
      float a = param[b] * freq;
      v = j < half ? cosf(a) : sinf(a);
    }
    out[i] = v;
  }
}

// Validation metric partial sums, one read of prediction and target: for every frame f (a contiguous run of P fp32 values =
// one (trajectory, step, channel) image) out[f] = (sum (a-b)^2, sum b^2, sum b).  relative_lp_loss (training_utils.py:9-23)
// after the datasets' affine de-normalisation x*std + mean (dataset/ns2d_fno_stage2_simpleae.py:78-79) follows on the host:
//   sum (a'-b')^2 = std^2 sum (a-b)^2,   sum b'^2 = std^2 sum b^2 + 2 std mean sum b + P mean^2.
// grid F (frames), block 256; fixed-order reduction (deterministic).
__global__ void __launch_bounds__(256) frame_sums_kernel(const float* __restrict__ a, const float* __restrict__ b, int P,
                                                         float* __restrict__ out) {
  __shared__ double red[8][3];
  const int64_t base = (int64_t)blockIdx.x * P;
  float d2 = 0.f, b2 = 0.f, b1 = 0.f;
  if ((P & 3) == 0 && ((reinterpret_cast<uintptr_t>(a + base) | reinterpret_cast<uintptr_t>(b + base)) & 15) == 0) {
    const float4* a4 = reinterpret_cast<const float4*>(a + base);
    const float4* b4 = reinterpret_cast<const float4*>(b + base);
    for (int i = threadIdx.x; i < P / 4; i += 256 * 4) {  // 4 pairs of 16-byte loads in flight per thread
      float4 va[4], vb[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int j = i + u * 256;
        const bool ok = j < P / 4;
        va[u] = ok ? __ldg(a4 + j) : make_float4(0.f, 0.f, 0.f, 0.f);
        vb[u] = ok ? __ldg(b4 + j) : make_float4(0.f, 0.f, 0.f, 0.f);
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const float dx = va[u].x - vb[u].x, dy = va[u].y - vb[u].y, dz = va[u].z - vb[u].z, dw = va[u].w - vb[u].w;
        d2 = fmaf(dx, dx, fmaf(dy, dy, fmaf(dz, dz, fmaf(dw, dw, d2))));
        b2 = fmaf(vb[u].x, vb[u].x, fmaf(vb[u].y, vb[u].y, fmaf(vb[u].z, vb[u].z, fmaf(vb[u].w, vb[u].w, b2))));
        b1 += (vb[u].x + vb[u].y) + (vb[u].z + vb[u].w);
      }
    }
  } else {
    for (int i = threadIdx.x; i < P; i += 256) {
      const float x = __ldg(a + base + i), y = __ldg(b + base + i);
      d2 = fmaf(x - y, x - y, d2);
      b2 = fmaf(y, y, b2);
      b1 += y;
    }
  }
  double s0 = warp_sum_d((double)d2), s1 = warp_sum_d((double)b2), s2 = warp_sum_d((double)b1);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane == 0) {
    red[warp][0] = s0;
    red[warp][1] = s1;
    red[warp][2] = s2;
  }
  __syncthreads();
  if (threadIdx.x < 3) {
    double t = 0.0;
    for (int w = 0; w < 8; ++w) t += red[w][threadIdx.x];
    out[(int64_t)blockIdx.x * 3 + threadIdx.x] = (float)t;
  }
}

// The same metric behind a NON-affine de-normalisation: the two-phase dataset (dataset/twophase_flow_stage2.py:369-389) scales
// velocity and pressure with their own statistics, zeroes the velocity on the four closed walls (Dirichlet) and clamps the vof
// channel to [0, 1 + 1e-8] -- applied to prediction AND target (train_stage2_twophase.py:251-252) before relative_lp_loss.
// Per frame f (channel c = f % C):  v = x * scale[c] + shift[c];  flags[c] & 1: v = 0 on the border;  flags[c] & 2: clamp(v, lo, hi).
// out[f] = (sum (p' - t')^2, sum t'^2, sum t') of the de-normalised values.  grid F, block 256, fixed-order reduction.
__global__ void __launch_bounds__(256) frame_sums_denorm_kernel(const float* __restrict__ a, const float* __restrict__ b, int H, int W, int C,
                                                                const float* __restrict__ scale, const float* __restrict__ shift,
                                                                const int* __restrict__ flags, float lo, float hi,
                                                                float* __restrict__ out) {
  __shared__ double red[8][3];
  const int P = H * W;
  const int64_t base = (int64_t)blockIdx.x * P;
  const int c = (int)(blockIdx.x % (unsigned)C);
  const float sc = __ldg(scale + c), sh = __ldg(shift + c);
  const int fl = __ldg(flags + c);
  float d2 = 0.f, b2 = 0.f, b1 = 0.f;
  for (int i = threadIdx.x; i < P; i += 256) {
    float x = fmaf(__ldg(a + base + i), sc, sh), y = fmaf(__ldg(b + base + i), sc, sh);
    if (fl & 1) {
      const int r = i / W, col = i - r * W;
      if (r == 0 || r == H - 1 || col == 0 || col == W - 1) x = y = 0.f;
    }
    if (fl & 2) {
      x = fminf(fmaxf(x, lo), hi);
      y = fminf(fmaxf(y, lo), hi);
    }
    d2 = fmaf(x - y, x - y, d2);
    b2 = fmaf(y, y, b2);
    b1 += y;
  }
  double s0 = warp_sum_d((double)d2), s1 = warp_sum_d((double)b2), s2 = warp_sum_d((double)b1);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane == 0) {
    red[warp][0] = s0;
    red[warp][1] = s1;
    red[warp][2] = s2;
  }
  __syncthreads();
  if (threadIdx.x < 3) {
    double t = 0.0;
    for (int w = 0; w < 8; ++w) t += red[w][threadIdx.x];
    out[(int64_t)blockIdx.x * 3 + threadIdx.x] = (float)t;
  }
}

}  // namespace lns

extern "C" {

int lns_frame_sums_denorm(const float* pred, const float* target, int64_t frames, int H, int W, int C, const float* scale,
                          const float* shift, const int* flags, float clamp_lo, float clamp_hi, float* out, void* stream) {
  LNS_REQUIRE(pred && target && out && scale && shift && flags && frames > 0 && H > 0 && W > 0 && C > 0 && frames < (1ll << 31) &&
                  frames % C == 0, "lns_frame_sums_denorm: bad arguments");
  lns::frame_sums_denorm_kernel<<<(unsigned)frames, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(pred, target, H, W, C, scale, shift,
                                                                                                     flags, clamp_lo, clamp_hi, out);
  return lns::check_launch("frame_sums_denorm_kernel");
}

int lns_frame_sums(const float* pred, const float* target, int64_t frames, int P, float* out, void* stream) {
  LNS_REQUIRE(pred && target && out && frames > 0 && P > 0 && frames < (1ll << 31), "lns_frame_sums: bad arguments");
  lns::frame_sums_kernel<<<(unsigned)frames, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(pred, target, P, out);
  return lns::check_launch("frame_sums_kernel");
}

int lns_nchw_to_nhwc(const float* x, int B, int C, int H, int W, int64_t x_bstride, void* y, int y_dtype,
                     int64_t y_bstride, void* stream) {
  LNS_REQUIRE(x && y && B > 0 && C > 0 && H > 0 && W > 0 && B <= 65535, "lns_nchw_to_nhwc: bad arguments");
  int P = H * W;
  dim3 grid(lns::cdiv(P, 32), lns::cdiv(C, 32), B), block(32, 8);
  lns::nchw_to_nhwc_kernel<<<grid, block, 0, reinterpret_cast<cudaStream_t>(stream)>>>(x, C, P, x_bstride, y, y_dtype,
                                                                                        y_bstride);
  return lns::check_launch("nchw_to_nhwc_kernel");
}

int lns_nhwc_to_nchw(const void* x, int x_dtype, int B, int H, int W, int C, int64_t x_bstride, float* y,
                     int64_t y_bstride, void* stream) {
  LNS_REQUIRE(x && y && B > 0 && C > 0 && H > 0 && W > 0 && B <= 65535, "lns_nhwc_to_nchw: bad arguments");
  int P = H * W;
  dim3 grid(lns::cdiv(P, 32), lns::cdiv(C, 32), B), block(32, 8);
  lns::nhwc_to_nchw_kernel<<<grid, block, 0, reinterpret_cast<cudaStream_t>(stream)>>>(x, x_dtype, C, P, x_bstride, y,
                                                                                        y_bstride);
  return lns::check_launch("nhwc_to_nchw_kernel");
}

int lns_fourier_embedding(const float* param, int B, int dim, float max_period, float* out, void* stream) {
  LNS_REQUIRE(param && out && B > 0 && dim > 0, "lns_fourier_embedding: bad arguments");
  int blocks = lns::cdiv((int64_t)B * dim, 256);
  lns::fourier_embedding_kernel<<<blocks, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(param, B, dim, max_period,
                                                                                             out);
  return lns::check_launch("fourier_embedding_kernel");
}

}  // extern "C"
