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

}  // namespace lns

extern "C" {

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
