// SABlock softmax attention core and the FABlock2D pieces (axis mean, low-rank kernel with rotary embedding,
// axial contractions).  CUDA-core fp32 arithmetic with operands staged in shared memory; sequences here are tiny
// (n <= 288 tokens for SABlock, n <= 96 per axis for FABlock2D), one (sample, head) fits in one CTA.
#include "common.cuh"

namespace lns {

// ---- softmax(q k^T * scale) v -------------------------------------------------------------------------
// grid (heads, B, qsplit), block 256 (8 warps, one query per warp at a time).
// smem: K_s [n][dh+1], V_s [n][dh], q_s [8][dh], p_s [8][n]
__global__ void __launch_bounds__(256) attention_kernel(const void* __restrict__ qkv, int dtype, int n, int heads, int dh,
                                                         float scale, void* __restrict__ out, int out_dtype) {
  extern __shared__ float sm[];
  float* K_s = sm;
  float* V_s = K_s + (size_t)n * (dh + 1);
  float* q_s = V_s + (size_t)n * dh;
  float* p_s = q_s + 8 * dh;
  const int h = blockIdx.x, b = blockIdx.y;
  const int hd = heads * dh;
  const int64_t row_stride = 3 * (int64_t)hd;
  const int64_t base = (int64_t)b * n * row_stride;
  for (int e = threadIdx.x; e < n * dh; e += blockDim.x) {
    int j = e / dh, d = e - j * dh;
    K_s[j * (dh + 1) + d] = ld_as_float(qkv, dtype, base + j * row_stride + hd + h * dh + d);
    V_s[j * dh + d] = ld_as_float(qkv, dtype, base + j * row_stride + 2 * hd + h * dh + d);
  }
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float* q_w = q_s + warp * dh;
  float* p_w = p_s + warp * n;
  const int qper = (n + gridDim.z - 1) / gridDim.z;
  const int q0 = blockIdx.z * qper, q1 = min(n, q0 + qper);
  for (int i = q0 + warp; i < q1; i += 8) {
    for (int d = lane; d < dh; d += 32) q_w[d] = ld_as_float(qkv, dtype, base + i * row_stride + h * dh + d);
    __syncwarp();
    float mx = -INFINITY;
    for (int j = lane; j < n; j += 32) {
      float s = 0.f;
      const float* kr = K_s + j * (dh + 1);
      for (int d = 0; d < dh; ++d) s = fmaf(q_w[d], kr[d], s);
      s *= scale;
      p_w[j] = s;
      mx = fmaxf(mx, s);
    }
    mx = warp_max(mx);
    float sum = 0.f;
    for (int j = lane; j < n; j += 32) {
      float e = expf(p_w[j] - mx);
      p_w[j] = e;
      sum += e;
    }
    sum = warp_sum(sum);
    __syncwarp();
    const float inv = 1.f / sum;
    for (int d = lane; d < dh; d += 32) {
      float o = 0.f;
      for (int j = 0; j < n; ++j) o = fmaf(p_w[j], V_s[j * dh + d], o);
      st_from_float(out, out_dtype, ((int64_t)b * n + i) * hd + h * dh + d, o * inv);
    }
    __syncwarp();
  }
}

// ---- mean over one spatial axis -----------------------------------------------------------------------
// grid (keep, B), block 256: thread -> (channel c, lane); fixed-order reduction.
__global__ void __launch_bounds__(256) axis_mean_kernel(const void* __restrict__ x, int dtype, int H, int W, int C,
                                                         int64_t bstride, int axis, float* __restrict__ out) {
  extern __shared__ float red[];  // [rows][C]
  const int rows = 256 / C;
  const int c = threadIdx.x % C, lane = threadIdx.x / C;
  const int keep = blockIdx.x, b = blockIdx.y;
  const int L = axis == 0 ? H : W;
  float s = 0.f;
  if (lane < rows) {
    for (int r = lane; r < L; r += rows) {
      int64_t pix = axis == 0 ? ((int64_t)r * W + keep) : ((int64_t)keep * W + r);
      s += ld_as_float(x, dtype, (int64_t)b * bstride + pix * C + c);
    }
    red[lane * C + c] = s;
  }
  __syncthreads();
  if (threadIdx.x < C) {
    float a = 0.f;
    for (int l = 0; l < rows; ++l) a += red[l * C + threadIdx.x];
    int K = axis == 0 ? W : H;
    out[((int64_t)b * K + keep) * C + threadIdx.x] = a / (float)L;
  }
}

// ---- LowRankKernel: rotary + q k^T ----------------------------------------------------------------------
// grid (heads, B), block 256. smem q_s, k_s: [n][d+1]
__global__ void __launch_bounds__(256) lowrank_kernel(const void* __restrict__ qk, int dtype, int n, int heads, int d,
                                                       const float* __restrict__ cos_t, const float* __restrict__ sin_t,
                                                       float scaling, float* __restrict__ Kout) {
  extern __shared__ float sm[];
  float* q_s = sm;
  float* k_s = sm + (size_t)n * (d + 1);
  const int h = blockIdx.x, b = blockIdx.y;
  const int hd = heads * d, half = d >> 1;
  const int64_t row_stride = 2 * (int64_t)hd;
  const int64_t base = (int64_t)b * n * row_stride;
  // each work item rotates one (token, frequency) pair of q and of k
  for (int e = threadIdx.x; e < n * half; e += blockDim.x) {
    int i = e / half, f = e - i * half;
    float cs = __ldg(cos_t + i * half + f), sn = __ldg(sin_t + i * half + f);
    int64_t o = base + i * row_stride + h * d;
    float q1 = ld_as_float(qk, dtype, o + f), q2 = ld_as_float(qk, dtype, o + f + half);
    float k1 = ld_as_float(qk, dtype, o + hd + f), k2 = ld_as_float(qk, dtype, o + hd + f + half);
    // t*cos + rotate_half(t)*sin with rotate_half(t) = cat(-t2, t1)   (modules/embedding.py:179-186)
    q_s[i * (d + 1) + f] = q1 * cs - q2 * sn;
    q_s[i * (d + 1) + f + half] = q2 * cs + q1 * sn;
    k_s[i * (d + 1) + f] = k1 * cs - k2 * sn;
    k_s[i * (d + 1) + f + half] = k2 * cs + k1 * sn;
  }
  __syncthreads();
  for (int e = threadIdx.x; e < n * n; e += blockDim.x) {
    int i = e / n, j = e - i * n;
    const float* qr = q_s + i * (d + 1);
    const float* kr = k_s + j * (d + 1);
    float s = 0.f;
    for (int t = 0; t < d; ++t) s = fmaf(qr[t], kr[t], s);
    Kout[(((int64_t)b * heads + h) * n + i) * n + j] = s * scaling;
  }
}

// ---- axial contraction ---------------------------------------------------------------------------------
// grid (lines, heads, B), block 256; ch <= 64.  smem: slab [n][ch], K_s [n][n]
__global__ void __launch_bounds__(256) axial_contract_kernel(const void* __restrict__ u, int dtype, int H, int W,
                                                              int heads, int ch, const float* __restrict__ Kmat,
                                                              int axis, void* __restrict__ out, int out_dtype) {
  extern __shared__ float sm[];
  const int n = axis == 0 ? H : W;
  float* slab = sm;
  float* K_s = sm + (size_t)n * ch;
  const int line = blockIdx.x, h = blockIdx.y, b = blockIdx.z;
  const int C = heads * ch;
  const int64_t sbase = (int64_t)b * H * W * C + h * ch;
  for (int e = threadIdx.x; e < n * ch; e += blockDim.x) {
    int j = e / ch, c = e - j * ch;
    int64_t pix = axis == 0 ? ((int64_t)j * W + line) : ((int64_t)line * W + j);
    slab[e] = ld_as_float(u, dtype, sbase + pix * C + c);
  }
  const float* Kg = Kmat + ((int64_t)b * heads + h) * n * n;
  for (int e = threadIdx.x; e < n * n; e += blockDim.x) K_s[e] = __ldg(Kg + e);
  __syncthreads();
  const int c = threadIdx.x % ch, ig = threadIdx.x / ch, ngrp = blockDim.x / ch;
  for (int i = ig; i < n; i += ngrp) {
    float acc = 0.f;
    const float* kr = K_s + i * n;
    for (int j = 0; j < n; ++j) acc = fmaf(kr[j], slab[j * ch + c], acc);
    int64_t pix = axis == 0 ? ((int64_t)i * W + line) : ((int64_t)line * W + i);
    st_from_float(out, out_dtype, sbase + pix * C + c, acc);
  }
}


// ---- axial contraction on tensor cores (bf16 in / bf16 out) -----------------------------------------------------------
// out_line[i][c] = sum_j K[i][j] * slab_line[j][c]  is a GEMM with M = K = n (padded to 16), N = 64 channels per line.
// Memory-bound (16 FLOP/B): the point of the tensor cores here is to keep the SM out of the way of the HBM stream.
// grid (ceil(lines/8), heads, B), block 128: a CTA stages K (bf16) and 8 line slabs in shared memory with cp.async,
// each warp owns 2 lines; mma.sync.m16n8k16 (bf16 x bf16 -> fp32) with ldmatrix / ldmatrix.trans operand fetch; results
// go through a per-warp staging tile so that every global store is a full 128-byte channel row.
constexpr int kAxLines = 8;      // lines per CTA (LPC = 4 for long lines: three CTAs per SM instead of one)
constexpr int kAxSlabStride = 72;  // bf16 elements per slab row (64 + 8 pad: conflict-free ldmatrix)

__device__ __forceinline__ void ldmatrix_x4(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr));
}
__device__ __forceinline__ void ldmatrix_x2_trans(uint32_t addr, uint32_t& r0, uint32_t& r1) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x2.trans.shared.b16 {%0,%1}, [%2];" : "=r"(r0), "=r"(r1) : "r"(addr));
}
// F16 = false: bf16 operands, true: IEEE-half operands (same rate, same fragment layout)
template <bool F16>
__device__ __forceinline__ void mma_h16_16816(float* c, const uint32_t* a, uint32_t b0, uint32_t b1) {
  if (F16)
    asm volatile(
        "mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
        : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
  else
    asm volatile(
        "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
        : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

template <bool F16, int LPC>
__global__ void __launch_bounds__(128) axial_contract_mma_kernel(const __nv_bfloat16* __restrict__ u, int H, int W, int heads,
                                                                  const float* __restrict__ Kmat, int axis,
                                                                  __nv_bfloat16* __restrict__ out) {
  extern __shared__ __align__(16) uint8_t smraw[];
  const int n = axis == 0 ? H : W;
  const int lines = axis == 0 ? W : H;
  const int n16 = (n + 15) & ~15;
  const int kstride = n16 + 8;
  __nv_bfloat16* K_s = reinterpret_cast<__nv_bfloat16*>(smraw);
  __nv_bfloat16* slab = K_s + (size_t)n16 * kstride;                             // [kAxLines][n16][72]
  __nv_bfloat16* stage = slab + (size_t)LPC * n16 * kAxSlabStride;            // [4 warps][16][72]
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int line0 = blockIdx.x * LPC, h = blockIdx.y, b = blockIdx.z;
  const int C = heads * 64;
  const int64_t sbase = (int64_t)b * H * W * C + h * 64;

  // K (fp32) -> bf16, zero padded to n16 x n16
  const float* Kg = Kmat + ((int64_t)b * heads + h) * n * n;
  if ((n & 3) == 0 && (reinterpret_cast<uintptr_t>(Kg) & 15) == 0) {
    // 16-byte loads, four in flight per thread (the scalar loop below is one dependent 4-byte load per element: at n = 96 that
    // was 72 exposed global-memory latencies per CTA, most of the kernel's 3.7 ms per launch on the 48x96 shallow-water level)
    const int q16 = n16 >> 2, total = n16 * q16;
    for (int e0 = tid; e0 < total; e0 += 4 * 128) {
      float4 v[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int e = e0 + u * 128;
        const int i = e / q16, j = (e - i * q16) * 4;
        v[u] = (e < total && i < n && j < n) ? __ldg(reinterpret_cast<const float4*>(Kg + i * n + j)) : make_float4(0.f, 0.f, 0.f, 0.f);
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int e = e0 + u * 128;
        if (e < total) {
          const int i = e / q16, j = (e - i * q16) * 4;
          *reinterpret_cast<uint2*>(K_s + i * kstride + j) = make_uint2(pack2_h16<F16>(v[u].x, v[u].y), pack2_h16<F16>(v[u].z, v[u].w));
        }
      }
    }
  } else {
    for (int e = tid; e < n16 * n16; e += 128) {
      int i = e / n16, j = e - i * n16;
      float v = (i < n && j < n) ? __ldg(Kg + i * n + j) : 0.f;
      reinterpret_cast<uint16_t*>(K_s)[i * kstride + j] = to_h16<F16>(v);
    }
  }
  // slabs: 8 x 16-byte chunks per (line, j) row
  const int nl = min(LPC, lines - line0);
  for (int e = tid; e < nl * n16 * 8; e += 128) {
    int ch = e & 7;
    int r = e >> 3;
    int l = r / n16, j = r - l * n16;
    uint32_t dst = (uint32_t)__cvta_generic_to_shared(slab + ((size_t)l * n16 + j) * kAxSlabStride + ch * 8);
    if (j < n) {
      int line = line0 + l;
      int64_t pix = axis == 0 ? ((int64_t)j * W + line) : ((int64_t)line * W + j);
      const __nv_bfloat16* src = u + sbase + pix * C + ch * 8;
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
    } else {
      *reinterpret_cast<uint4*>(slab + ((size_t)l * n16 + j) * kAxSlabStride + ch * 8) = make_uint4(0, 0, 0, 0);
    }
  }
  asm volatile("cp.async.commit_group;" ::: "memory");
  asm volatile("cp.async.wait_group 0;" ::: "memory");
  __syncthreads();

  __nv_bfloat16* my_stage = stage + (size_t)warp * 16 * kAxSlabStride;
  const int ktiles = n16 >> 4;
  for (int l = warp * (LPC / 4); l < (warp + 1) * (LPC / 4) && l < nl; ++l) {
    const __nv_bfloat16* sl = slab + (size_t)l * n16 * kAxSlabStride;
    const int line = line0 + l;
    for (int mt = 0; mt < ktiles; ++mt) {
      float acc[8][4];
#pragma unroll
      for (int nt = 0; nt < 8; ++nt) acc[nt][0] = acc[nt][1] = acc[nt][2] = acc[nt][3] = 0.f;
      for (int kt = 0; kt < ktiles; ++kt) {
        uint32_t a[4];
        ldmatrix_x4((uint32_t)__cvta_generic_to_shared(K_s + (size_t)(mt * 16 + (lane & 15)) * kstride + kt * 16 + (lane >> 4) * 8),
                    a[0], a[1], a[2], a[3]);
#pragma unroll
        for (int nt = 0; nt < 8; ++nt) {
          uint32_t b0, b1;
          ldmatrix_x2_trans((uint32_t)__cvta_generic_to_shared(sl + (size_t)(kt * 16 + (lane & 15)) * kAxSlabStride + nt * 8), b0, b1);
          mma_h16_16816<F16>(acc[nt], a, b0, b1);
        }
      }
      // fragment -> staging tile [16][64] (bf16)
      __syncwarp();
#pragma unroll
      for (int nt = 0; nt < 8; ++nt) {
        int r = lane >> 2, c = nt * 8 + (lane & 3) * 2;
        *reinterpret_cast<uint32_t*>(my_stage + r * kAxSlabStride + c) = pack2_h16<F16>(acc[nt][0], acc[nt][1]);
        *reinterpret_cast<uint32_t*>(my_stage + (r + 8) * kAxSlabStride + c) = pack2_h16<F16>(acc[nt][2], acc[nt][3]);
      }
      __syncwarp();
      // 4 rows per pass, 8 lanes x 16 bytes per row
#pragma unroll
      for (int pass = 0; pass < 4; ++pass) {
        int r = pass * 4 + (lane >> 3);
        int i = mt * 16 + r;
        if (i < n) {
          int64_t pix = axis == 0 ? ((int64_t)i * W + line) : ((int64_t)line * W + i);
          uint4 v = *reinterpret_cast<const uint4*>(my_stage + r * kAxSlabStride + (lane & 7) * 8);
          *reinterpret_cast<uint4*>(out + sbase + pix * C + (lane & 7) * 8) = v;
        }
      }
    }
  }
}


// ---- SABlock attention on tensor cores (bf16 in / bf16 out, head dim 64) ---------------------------------------------
// grid (heads, B), block 128.  Q, K, V head slices are staged in shared memory (bf16, 72-element rows: conflict-free
// ldmatrix); each warp owns 16-query tiles and walks the keys in blocks of 64 with an online softmax (exp2, fp32
// statistics): S = Q K^T and O += P V are mma.sync.m16n8k16, P never leaves registers (S accumulator fragments are
// re-packed as the A operand of the second GEMM).
constexpr int kAttStride = 72;

template <bool F16>
__global__ void __launch_bounds__(128) attention_mma_kernel(const __nv_bfloat16* __restrict__ qkv, int n, int heads,
                                                             float scale_log2e, __nv_bfloat16* __restrict__ out) {
  extern __shared__ __align__(16) uint8_t smraw[];
  const int n16 = (n + 15) & ~15;
  const int nk = (n + 63) & ~63;  // keys padded to whole 64-key blocks (zero rows, masked below)
  __nv_bfloat16* Q_s = reinterpret_cast<__nv_bfloat16*>(smraw);
  __nv_bfloat16* K_s = Q_s + (size_t)n16 * kAttStride;
  __nv_bfloat16* V_s = K_s + (size_t)nk * kAttStride;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int h = blockIdx.x, b = blockIdx.y;
  const int hd = heads * 64;
  const int64_t row_stride = 3 * (int64_t)hd;
  const __nv_bfloat16* base = qkv + (int64_t)b * n * row_stride + h * 64;
  // stage Q | K | V : 8 x 16-byte chunks per token row
  for (int e = tid; e < (n16 + 2 * nk) * 8; e += 128) {
    int ch = e & 7, r = e >> 3;
    int which, tok;
    __nv_bfloat16* dst;
    if (r < n16) { which = 0; tok = r; dst = Q_s + (size_t)tok * kAttStride; }
    else if (r < n16 + nk) { which = 1; tok = r - n16; dst = K_s + (size_t)tok * kAttStride; }
    else { which = 2; tok = r - n16 - nk; dst = V_s + (size_t)tok * kAttStride; }
    if (tok < n) {
      uint32_t d32 = (uint32_t)__cvta_generic_to_shared(dst + ch * 8);
      const __nv_bfloat16* src = base + (int64_t)tok * row_stride + which * hd + ch * 8;
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d32), "l"(src) : "memory");
    } else {
      *reinterpret_cast<uint4*>(dst + ch * 8) = make_uint4(0, 0, 0, 0);
    }
  }
  asm volatile("cp.async.commit_group;" ::: "memory");
  asm volatile("cp.async.wait_group 0;" ::: "memory");
  __syncthreads();

  const int g = lane >> 2, t = lane & 3;
  for (int mt = warp; mt < (n16 >> 4); mt += 4) {
    // Q fragments for the 4 k-steps of d = 64
    uint32_t qa[4][4];
#pragma unroll
    for (int ks = 0; ks < 4; ++ks)
      ldmatrix_x4((uint32_t)__cvta_generic_to_shared(Q_s + (size_t)(mt * 16 + (lane & 15)) * kAttStride + ks * 16 + (lane >> 4) * 8),
                  qa[ks][0], qa[ks][1], qa[ks][2], qa[ks][3]);
    float o[8][4];
#pragma unroll
    for (int i = 0; i < 8; ++i) o[i][0] = o[i][1] = o[i][2] = o[i][3] = 0.f;
    float m0 = -INFINITY, m1 = -INFINITY, l0 = 0.f, l1 = 0.f;  // running max / sum of rows g and g + 8
    for (int kb = 0; kb < nk; kb += 64) {
      float sacc[8][4];
#pragma unroll
      for (int nt = 0; nt < 8; ++nt) {
        sacc[nt][0] = sacc[nt][1] = sacc[nt][2] = sacc[nt][3] = 0.f;
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) {
          uint32_t b0, b1;
          // K rows are keys, columns d: the (non-transposed) 8x8 tiles are exactly the col-major B fragments
          asm volatile("ldmatrix.sync.aligned.m8n8.x2.shared.b16 {%0,%1}, [%2];" : "=r"(b0), "=r"(b1)
                       : "r"((uint32_t)__cvta_generic_to_shared(K_s + (size_t)(kb + nt * 8 + (lane & 7)) * kAttStride + ks * 16 + ((lane >> 3) & 1) * 8)));
          mma_h16_16816<F16>(sacc[nt], qa[ks], b0, b1);
        }
      }
      // scale, mask padded keys, block max
      float bm0 = -INFINITY, bm1 = -INFINITY;
#pragma unroll
      for (int nt = 0; nt < 8; ++nt) {
        int key = kb + nt * 8 + t * 2;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          float v = sacc[nt][j] * scale_log2e;
          if (key + (j & 1) >= n) v = -INFINITY;
          sacc[nt][j] = v;
        }
        bm0 = fmaxf(bm0, fmaxf(sacc[nt][0], sacc[nt][1]));
        bm1 = fmaxf(bm1, fmaxf(sacc[nt][2], sacc[nt][3]));
      }
      bm0 = fmaxf(bm0, __shfl_xor_sync(0xffffffffu, bm0, 1)); bm0 = fmaxf(bm0, __shfl_xor_sync(0xffffffffu, bm0, 2));
      bm1 = fmaxf(bm1, __shfl_xor_sync(0xffffffffu, bm1, 1)); bm1 = fmaxf(bm1, __shfl_xor_sync(0xffffffffu, bm1, 2));
      const float nm0 = fmaxf(m0, bm0), nm1 = fmaxf(m1, bm1);
      const float c0 = exp2f(m0 - nm0), c1 = exp2f(m1 - nm1);  // first block: exp2(-inf) = 0
      m0 = nm0; m1 = nm1;
      l0 *= c0; l1 *= c1;
#pragma unroll
      for (int i = 0; i < 8; ++i) { o[i][0] *= c0; o[i][1] *= c0; o[i][2] *= c1; o[i][3] *= c1; }
      float rs0 = 0.f, rs1 = 0.f;
#pragma unroll
      for (int nt = 0; nt < 8; ++nt) {
        sacc[nt][0] = exp2f(sacc[nt][0] - m0); sacc[nt][1] = exp2f(sacc[nt][1] - m0);
        sacc[nt][2] = exp2f(sacc[nt][2] - m1); sacc[nt][3] = exp2f(sacc[nt][3] - m1);
        rs0 += sacc[nt][0] + sacc[nt][1];
        rs1 += sacc[nt][2] + sacc[nt][3];
      }
      l0 += rs0; l1 += rs1;
      // O += P V over this key block: 4 k-steps of 16 keys
#pragma unroll
      for (int kk = 0; kk < 4; ++kk) {
        uint32_t pa[4];
        pa[0] = pack2_h16<F16>(sacc[2 * kk][0], sacc[2 * kk][1]);
        pa[1] = pack2_h16<F16>(sacc[2 * kk][2], sacc[2 * kk][3]);
        pa[2] = pack2_h16<F16>(sacc[2 * kk + 1][0], sacc[2 * kk + 1][1]);
        pa[3] = pack2_h16<F16>(sacc[2 * kk + 1][2], sacc[2 * kk + 1][3]);
#pragma unroll
        for (int nt = 0; nt < 8; ++nt) {
          uint32_t b0, b1;
          ldmatrix_x2_trans((uint32_t)__cvta_generic_to_shared(V_s + (size_t)(kb + kk * 16 + (lane & 15)) * kAttStride + nt * 8), b0, b1);
          mma_h16_16816<F16>(o[nt], pa, b0, b1);
        }
      }
    }
    l0 += __shfl_xor_sync(0xffffffffu, l0, 1); l0 += __shfl_xor_sync(0xffffffffu, l0, 2);
    l1 += __shfl_xor_sync(0xffffffffu, l1, 1); l1 += __shfl_xor_sync(0xffffffffu, l1, 2);
    const float i0 = 1.f / l0, i1 = 1.f / l1;
    const int r0 = mt * 16 + g, r1 = r0 + 8;
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
      int col = h * 64 + nt * 8 + t * 2;
      if (r0 < n) *reinterpret_cast<uint32_t*>(out + ((int64_t)b * n + r0) * hd + col) = pack2_h16<F16>(o[nt][0] * i0, o[nt][1] * i0);
      if (r1 < n) *reinterpret_cast<uint32_t*>(out + ((int64_t)b * n + r1) * hd + col) = pack2_h16<F16>(o[nt][2] * i1, o[nt][3] * i1);
    }
  }
}


// ---- LowRankKernel on tensor cores (bf16 q|k rows) ---------------------------------------------------------------------
// K[b][h] = rot(q) rot(k)^T  is an n x n x d GEMM per (sample, head) with n <= 96, d = 128.  grid (heads, B), block 128:
// the rotary embedding is applied while staging q and k (bf16) in shared memory, each warp owns 16-row tiles of K.
template <bool F16>
__global__ void __launch_bounds__(128) lowrank_mma_kernel(const __nv_bfloat16* __restrict__ qk, int n, int heads, int d,
                                                           const float* __restrict__ cos_t, const float* __restrict__ sin_t,
                                                           float scaling, float* __restrict__ Kout) {
  extern __shared__ __align__(16) uint8_t smraw[];
  const int n16 = (n + 15) & ~15;
  const int stride = d + 8;
  __nv_bfloat16* q_s = reinterpret_cast<__nv_bfloat16*>(smraw);
  __nv_bfloat16* k_s = q_s + (size_t)n16 * stride;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int h = blockIdx.x, b = blockIdx.y;
  const int hd = heads * d, half = d >> 1;
  const int64_t row_stride = 2 * (int64_t)hd;
  const __nv_bfloat16* base = qk + (int64_t)b * n * row_stride + h * d;
  // rotate pairs (f, f + d/2); 2 consecutive f per work item so that loads / stores are 4 bytes
  for (int e = tid; e < n16 * (half >> 1); e += 128) {
    const int i = e / (half >> 1), f = (e - i * (half >> 1)) * 2;
    uint32_t q1 = 0u, q2 = 0u, k1 = 0u, k2 = 0u;
    if (i < n) {
      const __nv_bfloat16* r = base + (int64_t)i * row_stride;
      float2 a1 = unpack2_h16<F16>(*reinterpret_cast<const uint32_t*>(r + f));
      float2 a2 = unpack2_h16<F16>(*reinterpret_cast<const uint32_t*>(r + f + half));
      float2 b1 = unpack2_h16<F16>(*reinterpret_cast<const uint32_t*>(r + hd + f));
      float2 b2 = unpack2_h16<F16>(*reinterpret_cast<const uint32_t*>(r + hd + f + half));
      const float c0 = __ldg(cos_t + i * half + f), c1 = __ldg(cos_t + i * half + f + 1);
      const float s0 = __ldg(sin_t + i * half + f), s1 = __ldg(sin_t + i * half + f + 1);
      q1 = pack2_h16<F16>(a1.x * c0 - a2.x * s0, a1.y * c1 - a2.y * s1);
      q2 = pack2_h16<F16>(a2.x * c0 + a1.x * s0, a2.y * c1 + a1.y * s1);
      k1 = pack2_h16<F16>(b1.x * c0 - b2.x * s0, b1.y * c1 - b2.y * s1);
      k2 = pack2_h16<F16>(b2.x * c0 + b1.x * s0, b2.y * c1 + b1.y * s1);
    }
    *reinterpret_cast<uint32_t*>(q_s + (size_t)i * stride + f) = q1;
    *reinterpret_cast<uint32_t*>(q_s + (size_t)i * stride + f + half) = q2;
    *reinterpret_cast<uint32_t*>(k_s + (size_t)i * stride + f) = k1;
    *reinterpret_cast<uint32_t*>(k_s + (size_t)i * stride + f + half) = k2;
  }
  __syncthreads();
  const int g = lane >> 2, t = lane & 3;
  float* Kg = Kout + ((int64_t)b * heads + h) * n * n;
  for (int mt = warp; mt < (n16 >> 4); mt += 4) {
    for (int nb = 0; nb < n16; nb += 64) {
      float acc[8][4];
#pragma unroll
      for (int nt = 0; nt < 8; ++nt) acc[nt][0] = acc[nt][1] = acc[nt][2] = acc[nt][3] = 0.f;
      for (int ks = 0; ks < (d >> 4); ++ks) {
        uint32_t a[4];
        ldmatrix_x4((uint32_t)__cvta_generic_to_shared(q_s + (size_t)(mt * 16 + (lane & 15)) * stride + ks * 16 + (lane >> 4) * 8),
                    a[0], a[1], a[2], a[3]);
#pragma unroll
        for (int nt = 0; nt < 8; ++nt) {
          if (nb + nt * 8 < n16) {
            uint32_t b0, b1;
            asm volatile("ldmatrix.sync.aligned.m8n8.x2.shared.b16 {%0,%1}, [%2];" : "=r"(b0), "=r"(b1)
                         : "r"((uint32_t)__cvta_generic_to_shared(k_s + (size_t)(nb + nt * 8 + (lane & 7)) * stride + ks * 16 + ((lane >> 3) & 1) * 8)));
            mma_h16_16816<F16>(acc[nt], a, b0, b1);
          }
        }
      }
      const int i0 = mt * 16 + g, i1 = i0 + 8;
#pragma unroll
      for (int nt = 0; nt < 8; ++nt) {
        const int j = nb + nt * 8 + t * 2;
        if (i0 < n && j < n) Kg[i0 * n + j] = acc[nt][0] * scaling;
        if (i0 < n && j + 1 < n) Kg[i0 * n + j + 1] = acc[nt][1] * scaling;
        if (i1 < n && j < n) Kg[i1 * n + j] = acc[nt][2] * scaling;
        if (i1 < n && j + 1 < n) Kg[i1 * n + j + 1] = acc[nt][3] * scaling;
      }
    }
  }
}

}  // namespace lns

extern "C" {

int lns_attention(const void* qkv, int dtype, int B, int n, int heads, int dh, float scale, void* out, int out_dtype,
                  void* stream) {
  LNS_REQUIRE(qkv && out && B > 0 && n > 0 && heads > 0 && dh > 0, "lns_attention: bad arguments");
  LNS_REQUIRE(B <= 65535, "lns_attention: batch %d exceeds grid limit, chunk the call", B);
  if (lns::is_h16_host(dtype) && out_dtype == dtype && dh == 64) {
    // tensor-core path (the bf16 rollout)
    int n16 = (n + 15) & ~15, nk = (n + 63) & ~63;
    size_t smem_mma = ((size_t)n16 + 2 * (size_t)nk) * lns::kAttStride * sizeof(__nv_bfloat16);
    if (smem_mma <= 227 * 1024) {
      { LNS_OPT_IN_SMEM((lns::attention_mma_kernel<false>), 227 * 1024, "attention"); LNS_OPT_IN_SMEM((lns::attention_mma_kernel<true>), 227 * 1024, "attention"); }
      dim3 grid(heads, B);
      auto kern = dtype == LNS_F16 ? lns::attention_mma_kernel<true> : lns::attention_mma_kernel<false>;
      kern<<<grid, 128, smem_mma, reinterpret_cast<cudaStream_t>(stream)>>>(
          reinterpret_cast<const __nv_bfloat16*>(qkv), n, heads, scale * 1.4426950408889634f,
          reinterpret_cast<__nv_bfloat16*>(out));
      return lns::check_launch("attention_mma_kernel");
    }
  }
  size_t smem = ((size_t)n * (dh + 1) + (size_t)n * dh + 8 * dh + 8 * (size_t)n) * sizeof(float);
  LNS_REQUIRE(smem <= 227 * 1024, "lns_attention: n=%d dh=%d needs %zu B shared memory", n, dh, smem);
  { LNS_OPT_IN_SMEM((lns::attention_kernel), 227 * 1024, "attention"); }
  int qsplit = 1;
  while ((int64_t)B * heads * qsplit < 296 && qsplit * 8 < n) qsplit *= 2;
  dim3 grid(heads, B, qsplit);
  lns::attention_kernel<<<grid, 256, smem, reinterpret_cast<cudaStream_t>(stream)>>>(qkv, dtype, n, heads, dh, scale, out,
                                                                                      out_dtype);
  return lns::check_launch("attention_kernel");
}

int lns_axis_mean(const void* x, int dtype, int B, int H, int W, int C, int64_t bstride, int axis, float* out,
                  void* stream) {
  LNS_REQUIRE(x && out && B > 0 && H > 0 && W > 0 && C > 0 && C <= 256 && (axis == 0 || axis == 1),
              "lns_axis_mean: bad arguments (C=%d)", C);
  LNS_REQUIRE(B <= 65535, "lns_axis_mean: batch %d exceeds grid limit, chunk the call", B);
  int rows = 256 / C;
  dim3 grid(axis == 0 ? W : H, B);
  lns::axis_mean_kernel<<<grid, 256, (size_t)rows * C * sizeof(float), reinterpret_cast<cudaStream_t>(stream)>>>(
      x, dtype, H, W, C, bstride, axis, out);
  return lns::check_launch("axis_mean_kernel");
}

int lns_lowrank_kernel(const void* qk, int dtype, int B, int n, int heads, int d, const float* cos_tab,
                       const float* sin_tab, float scaling, float* K, void* stream) {
  LNS_REQUIRE(qk && K && cos_tab && sin_tab && B > 0 && n > 0 && heads > 0 && d > 0 && d % 2 == 0,
              "lns_lowrank_kernel: bad arguments");
  LNS_REQUIRE(B <= 65535, "lns_lowrank_kernel: batch %d exceeds grid limit, chunk the call", B);
  if (lns::is_h16_host(dtype) && d % 16 == 0) {
    // tensor-core path (bf16 q|k from the tcgen05 to_qk GEMM)
    int n16 = (n + 15) & ~15;
    size_t smem_mma = 2 * (size_t)n16 * (d + 8) * sizeof(__nv_bfloat16);
    if (smem_mma <= 227 * 1024) {
      { LNS_OPT_IN_SMEM((lns::lowrank_mma_kernel<false>), 227 * 1024, "attention"); LNS_OPT_IN_SMEM((lns::lowrank_mma_kernel<true>), 227 * 1024, "attention"); }
      dim3 grid(heads, B);
      auto kern = dtype == LNS_F16 ? lns::lowrank_mma_kernel<true> : lns::lowrank_mma_kernel<false>;
      kern<<<grid, 128, smem_mma, reinterpret_cast<cudaStream_t>(stream)>>>(
          reinterpret_cast<const __nv_bfloat16*>(qk), n, heads, d, cos_tab, sin_tab, scaling, K);
      return lns::check_launch("lowrank_mma_kernel");
    }
  }
  size_t smem = 2 * (size_t)n * (d + 1) * sizeof(float);
  LNS_REQUIRE(smem <= 227 * 1024, "lns_lowrank_kernel: n=%d d=%d needs %zu B shared memory", n, d, smem);
  { LNS_OPT_IN_SMEM((lns::lowrank_kernel), 227 * 1024, "attention"); }
  dim3 grid(heads, B);
  lns::lowrank_kernel<<<grid, 256, smem, reinterpret_cast<cudaStream_t>(stream)>>>(qk, dtype, n, heads, d, cos_tab,
                                                                                    sin_tab, scaling, K);
  return lns::check_launch("lowrank_kernel");
}

int lns_axial_contract(const void* u, int dtype, int B, int H, int W, int heads, int ch, const float* K, int axis,
                       void* out, int out_dtype, void* stream) {
  LNS_REQUIRE(u && out && K && B > 0 && H > 0 && W > 0 && heads > 0 && ch > 0 && ch <= 256 && 256 % ch == 0 &&
                  (axis == 0 || axis == 1),
              "lns_axial_contract: bad arguments (ch=%d)", ch);
  LNS_REQUIRE(B <= 65535, "lns_axial_contract: batch %d exceeds grid limit, chunk the call", B);
  int n = axis == 0 ? H : W;
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  if (lns::is_h16_host(dtype) && out_dtype == dtype && ch == 64) {
    // tensor-core path (the bf16 rollout)
    int n16 = (n + 15) & ~15;
    int lines = axis == 0 ? W : H;
    // long lines (n16 >= 64: the 48x96 shallow-water level): 4 lines per CTA so that three CTAs share an SM -- with 8 the 130 KB
    // slab left ONE 4-warp CTA per SM and every load / ldmatrix latency exposed (3.4 ms per launch at 320 x 48x96 x 512)
    const int lpc = n16 >= 64 ? 4 : lns::kAxLines;
    size_t smem_mma = ((size_t)n16 * (n16 + 8) + (size_t)lpc * n16 * lns::kAxSlabStride +
                       4 * 16 * (size_t)lns::kAxSlabStride) * sizeof(__nv_bfloat16);
    LNS_REQUIRE(smem_mma <= 227 * 1024, "lns_axial_contract: n=%d needs %zu B shared memory", n, smem_mma);
    {
      LNS_OPT_IN_SMEM((lns::axial_contract_mma_kernel<false, 8>), 227 * 1024, "attention");
      LNS_OPT_IN_SMEM((lns::axial_contract_mma_kernel<true, 8>), 227 * 1024, "attention");
      LNS_OPT_IN_SMEM((lns::axial_contract_mma_kernel<false, 4>), 227 * 1024, "attention");
      LNS_OPT_IN_SMEM((lns::axial_contract_mma_kernel<true, 4>), 227 * 1024, "attention");
    }
    dim3 grid(lns::cdiv(lines, lpc), heads, B);
    const bool f16 = dtype == LNS_F16;
    auto kern = lpc == 4 ? (f16 ? lns::axial_contract_mma_kernel<true, 4> : lns::axial_contract_mma_kernel<false, 4>)
                         : (f16 ? lns::axial_contract_mma_kernel<true, 8> : lns::axial_contract_mma_kernel<false, 8>);
    kern<<<grid, 128, smem_mma, s>>>(reinterpret_cast<const __nv_bfloat16*>(u), H, W, heads, K, axis,
                                                               reinterpret_cast<__nv_bfloat16*>(out));
    return lns::check_launch("axial_contract_mma_kernel");
  }
  size_t smem = ((size_t)n * ch + (size_t)n * n) * sizeof(float);
  LNS_REQUIRE(smem <= 227 * 1024, "lns_axial_contract: n=%d needs %zu B shared memory", n, smem);
  { LNS_OPT_IN_SMEM((lns::axial_contract_kernel), 227 * 1024, "attention"); }
  dim3 grid(axis == 0 ? W : H, heads, B);
  lns::axial_contract_kernel<<<grid, 256, smem, s>>>(u, dtype, H, W, heads, ch, K, axis, out, out_dtype);
  return lns::check_launch("axial_contract_kernel");
}

}  // extern "C"
