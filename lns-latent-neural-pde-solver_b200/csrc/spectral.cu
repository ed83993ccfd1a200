// Spectral convolution (FNO layer) as shared-memory-staged truncated DFTs with the complex mode-weight multiply fused.
//
//   reference:  x_ft = rfft2(x); out_ft[:, :, :m1, :m2]  = einsum(x_ft[:, :, :m1, :m2]  (* emb1), w1)
//                                out_ft[:, :, -m1:, :m2] = einsum(x_ft[:, :, -m1:, :m2] (* emb2), w2); irfft2(out_ft, s=(H,W))
//   (modules/basics.py:129-148, modules/fourier_cond.py:52-81)
//
// Only 2*m1 x m2 of the H x (W/2+1) spectrum is ever used, so instead of full FFTs (which also cannot do H=61, W=121
// in radix form) three kernels compute exactly the retained modes:
//   A  rows:    X1[b,y,kx,ci]  = sum_x x[b,y,x,ci] e^{-2 pi i kx x / W}                         (real -> complex)
//   BC columns: X2[r,ci] = sum_y X1[y,ci] e^{-2 pi i ky_r y / H};  X2 *= emb;  Y[r,co] = sum_ci X2[r,ci] Wm[r,kx,ci,co];
//               Z[b,y,kx,co] = 1/H sum_r Y[r,co] e^{+2 pi i ky_r y / H}                          (all in one CTA)
//   D  rows:    out[b,y,x,co] = 1/W sum_kx f_kx Re(Z[b,y,kx,co] e^{+2 pi i kx x / W}),  f_0 = f_Nyquist = 1, else 2
//               (the c2r transform ignores Im of the kx = 0 / Nyquist columns; sin() is 0 there, so it drops out).
// ky_r = r for r < m1, H - 2*m1 + r otherwise.  Twiddles come from sincospif of an exactly reduced integer phase.
#include "common.cuh"

namespace lns {

__device__ __forceinline__ float2 cmul(float2 a, float2 b) { return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x); }

// grid (H, B), block 256.  smem: row [W][Ci] floats, tw [W] float2 (cos, sin of 2 pi t / W)
__global__ void __launch_bounds__(256) spectral_rows_fwd(const void* __restrict__ x, int dtype, int H, int W, int Ci, int m2,
                                                          float2* __restrict__ X1) {
  extern __shared__ float sm[];
  float* row = sm;
  float2* tw = reinterpret_cast<float2*>(sm + (size_t)W * Ci);
  const int y = blockIdx.x, b = blockIdx.y;
  const int64_t base = ((int64_t)b * H + y) * W * Ci;
  for (int e = threadIdx.x; e < W * Ci; e += blockDim.x) row[e] = ld_as_float(x, dtype, base + e);
  for (int t = threadIdx.x; t < W; t += blockDim.x) {
    float s, c;
    sincospif(2.0f * (float)t / (float)W, &s, &c);
    tw[t] = make_float2(c, s);
  }
  __syncthreads();
  for (int e = threadIdx.x; e < m2 * Ci; e += blockDim.x) {
    int ci = e % Ci, kx = e / Ci;
    float re = 0.f, im = 0.f;
    int ph = 0;
    for (int xx = 0; xx < W; ++xx) {
      float v = row[xx * Ci + ci];
      float2 t = tw[ph];
      re = fmaf(v, t.x, re);
      im = fmaf(-v, t.y, im);
      ph += kx;
      if (ph >= W) ph -= W;
    }
    X1[(((int64_t)b * H + y) * m2 + kx) * Ci + ci] = make_float2(re, im);
  }
}

// grid (m2, B), block 256.
// smem: col [H][Ci] float2, X2 [2 m1][Ci] float2, Y [2 m1][Co] float2, tw [H] float2
__global__ void __launch_bounds__(256) spectral_cols_mix(const float2* __restrict__ X1, int H, int Ci, int Co, int m1, int m2,
                                                          const float2* __restrict__ Wm, const float2* __restrict__ emb,
                                                          float2* __restrict__ Z) {
  extern __shared__ float sm[];
  float2* col = reinterpret_cast<float2*>(sm);
  float2* X2 = col + (size_t)H * Ci;
  float2* Y = X2 + (size_t)2 * m1 * Ci;
  float2* tw = Y + (size_t)2 * m1 * Co;
  const int kx = blockIdx.x, b = blockIdx.y;
  const int R = 2 * m1;
  for (int e = threadIdx.x; e < H * Ci; e += blockDim.x) {
    int yy = e / Ci, ci = e - yy * Ci;
    col[e] = X1[(((int64_t)b * H + yy) * m2 + kx) * Ci + ci];
  }
  for (int t = threadIdx.x; t < H; t += blockDim.x) {
    float s, c;
    sincospif(2.0f * (float)t / (float)H, &s, &c);
    tw[t] = make_float2(c, s);
  }
  __syncthreads();
  for (int e = threadIdx.x; e < R * Ci; e += blockDim.x) {
    int r = e / Ci, ci = e - r * Ci;
    int ky = r < m1 ? r : H - 2 * m1 + r;
    float2 acc = make_float2(0.f, 0.f);
    int ph = 0;
    for (int yy = 0; yy < H; ++yy) {
      float2 t = tw[ph];  // e^{-i th} = (cos, -sin)
      float2 v = col[yy * Ci + ci];
      acc.x += v.x * t.x + v.y * t.y;
      acc.y += v.y * t.x - v.x * t.y;
      ph += ky;
      if (ph >= H) ph -= H;
    }
    if (emb) {
      // emb [B][m1][m2][block][re,im]   (FreqLinear output, modules/fourier_cond.py:25-29)
      int blk = r < m1 ? 0 : 1, rr = r < m1 ? r : r - m1;
      acc = cmul(acc, emb[(((int64_t)b * m1 + rr) * m2 + kx) * 2 + blk]);
    }
    X2[e] = acc;
  }
  __syncthreads();
  for (int e = threadIdx.x; e < R * Co; e += blockDim.x) {
    int r = e / Co, co = e - r * Co;
    int blk = r < m1 ? 0 : 1, rr = r < m1 ? r : r - m1;
    const float2* wp = Wm + ((((int64_t)blk * m1 + rr) * m2 + kx) * Ci) * Co + co;
    float2 acc = make_float2(0.f, 0.f);
    for (int ci = 0; ci < Ci; ++ci) {
      float2 a = X2[r * Ci + ci];
      float2 w = __ldg(wp + (int64_t)ci * Co);
      acc.x += a.x * w.x - a.y * w.y;
      acc.y += a.x * w.y + a.y * w.x;
    }
    Y[e] = acc;
  }
  __syncthreads();
  const float invH = 1.0f / (float)H;
  for (int e = threadIdx.x; e < H * Co; e += blockDim.x) {
    int yy = e / Co, co = e - yy * Co;
    float2 acc = make_float2(0.f, 0.f);
    for (int r = 0; r < R; ++r) {
      int ky = r < m1 ? r : H - 2 * m1 + r;
      int ph = (int)(((int64_t)ky * yy) % H);
      float2 t = tw[ph];  // e^{+i th} = (cos, +sin)
      float2 v = Y[r * Co + co];
      acc.x += v.x * t.x - v.y * t.y;
      acc.y += v.x * t.y + v.y * t.x;
    }
    Z[(((int64_t)b * H + yy) * m2 + kx) * Co + co] = make_float2(acc.x * invH, acc.y * invH);
  }
}

// grid (H, B), block 256.  smem: Zrow [m2][Co] float2, tw [W] float2
__global__ void __launch_bounds__(256) spectral_rows_inv(const float2* __restrict__ Z, int H, int W, int Co, int m2,
                                                          float* __restrict__ out) {
  extern __shared__ float sm[];
  float2* zr = reinterpret_cast<float2*>(sm);
  float2* tw = zr + (size_t)m2 * Co;
  const int y = blockIdx.x, b = blockIdx.y;
  for (int e = threadIdx.x; e < m2 * Co; e += blockDim.x) zr[e] = Z[((int64_t)b * H + y) * m2 * Co + e];
  for (int t = threadIdx.x; t < W; t += blockDim.x) {
    float s, c;
    sincospif(2.0f * (float)t / (float)W, &s, &c);
    tw[t] = make_float2(c, s);
  }
  __syncthreads();
  const float invW = 1.0f / (float)W;
  for (int e = threadIdx.x; e < W * Co; e += blockDim.x) {
    int xx = e / Co, co = e - xx * Co;
    float acc = 0.f;
    int ph = 0;
    for (int kx = 0; kx < m2; ++kx) {
      float2 t = tw[ph];
      float2 v = zr[kx * Co + co];
      float f = (kx == 0 || 2 * kx == W) ? 1.f : 2.f;
      acc += f * (v.x * t.x - v.y * t.y);
      ph += xx;
      if (ph >= W) ph -= W;
    }
    out[(((int64_t)b * H + y) * W + xx) * Co + co] = acc * invW;
  }
}

}  // namespace lns

extern "C" {

int64_t lns_spectral_work_bytes(int B, int H, int W, int Ci, int Co, int m1, int m2) {
  (void)W; (void)m1;
  return (int64_t)B * H * m2 * ((int64_t)Ci + Co) * 8;
}

int lns_spectral_conv2d(const void* x, int x_dtype, int B, int H, int W, int Ci, int Co, int m1, int m2,
                        const float* w_modes, const float* emb, void* work, float* out, void* stream) {
  LNS_REQUIRE(x && w_modes && work && out && B > 0 && H > 0 && W > 0 && Ci > 0 && Co > 0, "lns_spectral_conv2d: bad arguments");
  LNS_REQUIRE(B <= 65535, "lns_spectral_conv2d: batch %d exceeds grid limit, chunk the call", B);
  LNS_REQUIRE(m1 >= 1 && 2 * m1 <= H, "lns_spectral_conv2d: needs 2*modes1 <= H (got %d, H=%d): overlapping row blocks "
              "(reference quirk modules/basics.py:143-145) are not supported", m1, H);
  LNS_REQUIRE(m2 >= 1 && m2 <= W / 2 + 1, "lns_spectral_conv2d: needs modes2 <= W/2+1 (got %d, W=%d)", m2, W);
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  float2* X1 = reinterpret_cast<float2*>(work);
  float2* Z = X1 + (int64_t)B * H * m2 * Ci;
  size_t smA = ((size_t)W * Ci + 2 * (size_t)W) * sizeof(float);
  size_t smB = ((size_t)H * Ci + 2 * (size_t)m1 * Ci + 2 * (size_t)m1 * Co + H) * sizeof(float2);
  size_t smD = ((size_t)m2 * Co + W) * sizeof(float2);
  LNS_REQUIRE(smA <= 227 * 1024 && smB <= 227 * 1024 && smD <= 227 * 1024,
              "lns_spectral_conv2d: shape needs %zu/%zu/%zu B shared memory", smA, smB, smD);
  LNS_OPT_IN_SMEM((lns::spectral_rows_fwd), 227 * 1024, "spectral");
  LNS_OPT_IN_SMEM((lns::spectral_cols_mix), 227 * 1024, "spectral");
  LNS_OPT_IN_SMEM((lns::spectral_rows_inv), 227 * 1024, "spectral");
  lns::spectral_rows_fwd<<<dim3(H, B), 256, smA, s>>>(x, x_dtype, H, W, Ci, m2, X1);
  int rc = lns::check_launch("spectral_rows_fwd");
  if (rc) return rc;
  lns::spectral_cols_mix<<<dim3(m2, B), 256, smB, s>>>(X1, H, Ci, Co, m1, m2, reinterpret_cast<const float2*>(w_modes),
                                                       reinterpret_cast<const float2*>(emb), Z);
  rc = lns::check_launch("spectral_cols_mix");
  if (rc) return rc;
  lns::spectral_rows_inv<<<dim3(H, B), 256, smD, s>>>(Z, H, W, Co, m2, out);
  return lns::check_launch("spectral_rows_inv");
}

}  // extern "C"
