// FABlock2D pooled branch in ONE kernel per axis:  pooled rows -> axial attention kernel  K[b][h] (n x n)
//   (modules/factorized_attention.py: to_in :114, PoolingReducer :72-94, LowRankKernel :43-69, rotary modules/embedding.py:163-186)
//     t  = pooled x (W_reducer_in . W_to_in)^T              the two bias-free 64x64 linears, composed on the host (fp64)
//     t  = LayerNorm(t);  t = GELU(t x W_f1^T);  z = t x W_f2^T + b_f2          [n][64]
//     qk = z x W_qk^T   [n][2*heads*d]  ->  q_h, k_h = rotary(qk slices)  ->  K[b][h] = q_h k_h^T * scaling
// Unfused this was 8 launches per axis on [B*n] x 64 rows -- half of them CUDA-core GEMMs on tiny matrices -- plus a
// 2048-channel intermediate written and re-read (537 MB per 4096 samples of 32 rows): 13 ms of the 128 ms round-1 rollout.
// One CTA takes S = 64/n samples (<= 64 rows): the small layers run on 3xTF32 mma.sync tiles out of shared memory (fp32-class
// accuracy, fp32 LayerNorm / GELU; weights arrive transposed [k][out]), the to_qk GEMM and q k^T run on mma.sync with
// 16-bit operands (the per-head 32 KB weight slice is streamed with cp.async and shared by the CTA's samples), q|k never
// leave shared memory.
#include "common.cuh"

namespace lns {
namespace {
constexpr int kAxRows = 64;            // rows (samples x positions) per CTA
constexpr int kAxRowsAlloc = 64 + 16;  // a sample's last 16-row MMA tile may overhang
constexpr int kXS = 68;                // fp32 row stride of the 64-wide activations (= 4 mod 32: conflict-free tf32 fragments)
constexpr int kYS = 132;               // fp32 row stride of the 128-wide hidden layer
constexpr int kWPad = 8;               // filter rows [k][N + 8] floats: the 4 k rows of a B fragment hit different banks
constexpr int kHS = 72;                // 16-bit row stride of the 64-wide MMA operands
constexpr int kQS = 136;               // 16-bit row stride of q / k (d = 128)

__device__ __forceinline__ uint32_t s32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void ldsm4(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr));
}
__device__ __forceinline__ void ldsm2(uint32_t addr, uint32_t& r0, uint32_t& r1) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x2.shared.b16 {%0,%1}, [%2];" : "=r"(r0), "=r"(r1) : "r"(addr));
}
template <bool F16>
__device__ __forceinline__ void mma16816(float* c, const uint32_t* a, uint32_t b0, uint32_t b1) {
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

struct AxisParams {
  const float* pooled;  // [B][n][64]
  int B, n, heads, S;   // S samples per CTA
  const float* w1t;     // [64 k][64 o]   composed (reducer.to_in . to_in), transposed
  const float* ln_g;    // [64]
  const float* ln_b;    // [64]
  float ln_eps;
  const float* wf1t;    // [64 k][128 o]  out_ffn.1 transposed
  const float* wf2t;    // [128 k][64 o]  out_ffn.3 transposed
  const float* bf2;     // [64]
  const uint16_t* wqk;  // [2*heads*128][64] 16-bit (bf16 or f16), rows: q (head, d) | k (head, d)
  const float* cos_t;   // [n][64]
  const float* sin_t;   // [n][64]
  float scaling;
  float* K;             // [B][heads][n][n]
};

__device__ __forceinline__ void mma_tf32(float* c, const uint32_t* a, uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

// One warp computes NTW 8-column tiles of  y[16 rows mt][N] = x[16 rows][Kdim] x wt[Kdim][N]  on mma.sync m16n8k8 with the
// 3xTF32 split (a = a_hi + a_lo, b = b_hi + b_lo, acc += a_lo b_hi + a_hi b_lo + a_hi b_hi: fp32-class accuracy, the
// kernels Kx / Ky scale everything downstream).  x: fp32 rows with stride xs (= 4 mod 32), wt: [Kdim][ws] fp32 (ws = N + 8).
// acc[j] = fragment of column tile nt0 + j.
// (The first version of this front end was scalar fp32 FMAs out of shared memory: 6.6k instructions per warp, half of the
//  kernel; as tensor-core tiles it is ~2k.)
__device__ __forceinline__ void split_tf32(float v, uint32_t& hi, uint32_t& lo) {
  const float h = round_tf32(v);
  hi = __float_as_uint(h);
  lo = __float_as_uint(round_tf32(v - h));
}
template <int NTW>
__device__ __forceinline__ void tile_tf32(const float* x, int xs, const float* wt, int ws, int Kdim, int mt, int nt0, int lane,
                                          float (&acc)[NTW][4]) {
  const int g = lane >> 2, t = lane & 3;
#pragma unroll
  for (int j = 0; j < NTW; ++j) acc[j][0] = acc[j][1] = acc[j][2] = acc[j][3] = 0.f;
  const float* xa = x + (mt * 16 + g) * xs + t;
  const float* wb = wt + t * ws + nt0 * 8 + g;
  for (int k = 0; k < Kdim; k += 8) {
    uint32_t ah[4], al[4];
    split_tf32(xa[k], ah[0], al[0]);
    split_tf32(xa[8 * xs + k], ah[1], al[1]);
    split_tf32(xa[k + 4], ah[2], al[2]);
    split_tf32(xa[8 * xs + k + 4], ah[3], al[3]);
#pragma unroll
    for (int j = 0; j < NTW; ++j) {
      uint32_t bh0, bl0, bh1, bl1;
      split_tf32(wb[k * ws + j * 8], bh0, bl0);
      split_tf32(wb[(k + 4) * ws + j * 8], bh1, bl1);
      mma_tf32(acc[j], al, bh0, bh1);
      mma_tf32(acc[j], ah, bl0, bl1);
      mma_tf32(acc[j], ah, bh0, bh1);
    }
  }
}

// [rows][N] fp32 global (16-byte aligned, N % 4 == 0) -> shared rows of pitch N + kWPad
__device__ __forceinline__ void cp_filter(float* dst, const float* src, int rows, int N, int tid) {
  const int cpr = N >> 2;  // 16-byte chunks per row
  for (int e = tid; e < rows * cpr; e += 256) {
    const int rr = e / cpr, c = e - rr * cpr;
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(s32(dst + rr * (N + kWPad) + c * 4)), "l"(src + rr * N + c * 4) : "memory");
  }
}
}  // namespace

// grid ceil(B / S), block 256
template <bool F16>
__global__ void __launch_bounds__(256, 2) fa_axis_kernel(const AxisParams p) {
  extern __shared__ __align__(16) uint8_t sm[];
  // persistent region
  uint16_t* Hs = reinterpret_cast<uint16_t*>(sm);                               // [kAxRows][kHS]  z as 16-bit MMA operand
  float2* cs_s = reinterpret_cast<float2*>(sm + kAxRows * kHS * 2);        // [n][64] (cos, sin) of position i, frequency f
  uint8_t* un = sm + kAxRows * kHS * 2 + p.n * 64 * 8;
  // front-end view of the union
  float* X = reinterpret_cast<float*>(un);                                      // [64][kXS]
  float* Y = X + kAxRows * kXS;                                                 // [64][kYS]
  float* Wb = Y + kAxRows * kYS;                                                // 128 x 72 or 64 x 136 floats
  // back-end view of the union
  uint16_t* Wq = reinterpret_cast<uint16_t*>(un);                               // [256][kHS]  q | k weight slice of one head
  uint16_t* Qs = Wq + 256 * kHS;                                                // [kAxRowsAlloc][kQS]
  uint16_t* Ks = Qs + kAxRowsAlloc * kQS;                                       // [kAxRowsAlloc][kQS]
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int n = p.n, heads = p.heads;
  const int b0 = blockIdx.x * p.S;
  const int ns = min(p.S, p.B - b0);  // samples of this CTA
  const int R = ns * n;               // valid rows
  const int g = lane >> 2, t = lane & 3;

  // ---- front end: three small GEMMs on 3xTF32 tensor-core tiles (fp32-class accuracy, fp32 LayerNorm / GELU) ----
  cp_filter(Wb, p.w1t, 64, 64, tid);
  asm volatile("cp.async.commit_group;" ::: "memory");
  for (int e = tid; e < kAxRows * 16; e += 256) {  // pooled rows -> X
    const int rr = e >> 4, c4 = e & 15;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (rr < R) v = __ldg(reinterpret_cast<const float4*>(p.pooled + ((int64_t)b0 * n + rr) * 64 + c4 * 4));
    *reinterpret_cast<float4*>(X + rr * kXS + c4 * 4) = v;
  }
  asm volatile("cp.async.wait_group 0;" ::: "memory");
  __syncthreads();
  {
    // t1 = pooled x W1^T: 4 m tiles x 8 column tiles -> warp w: m tile w >> 1, column tiles (w & 1) * 4 .. + 3
    float acc[4][4];
    const int mt = warp >> 1, nt0 = (warp & 1) * 4;
    tile_tf32<4>(X, kXS, Wb, 64 + kWPad, 64, mt, nt0, lane, acc);
    __syncthreads();  // every warp has read X and Wb
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int c = (nt0 + j) * 8 + t * 2;
      *reinterpret_cast<float2*>(Y + (mt * 16 + g) * kYS + c) = make_float2(acc[j][0], acc[j][1]);
      *reinterpret_cast<float2*>(Y + (mt * 16 + g + 8) * kYS + c) = make_float2(acc[j][2], acc[j][3]);
    }
  }
  cp_filter(Wb, p.wf1t, 64, 128, tid);
  asm volatile("cp.async.commit_group;" ::: "memory");
  __syncthreads();
  {  // LayerNorm over the 64 features of each row: 4 lanes per row (fp32, two passes), result -> X
    const int rr = tid >> 2, part = tid & 3;
    const float* yr = Y + rr * kYS + part * 16;
    float v[16];
    float m = 0.f;
#pragma unroll
    for (int k = 0; k < 16; ++k) {
      v[k] = yr[k];
      m += v[k];
    }
    m += __shfl_xor_sync(0xffffffffu, m, 1);
    m += __shfl_xor_sync(0xffffffffu, m, 2);
    m *= (1.f / 64.f);
    float var = 0.f;
#pragma unroll
    for (int k = 0; k < 16; ++k) {
      const float d = v[k] - m;
      var = fmaf(d, d, var);
    }
    var += __shfl_xor_sync(0xffffffffu, var, 1);
    var += __shfl_xor_sync(0xffffffffu, var, 2);
    const float rstd = rsqrtf(var * (1.f / 64.f) + p.ln_eps);
#pragma unroll
    for (int k = 0; k < 16; ++k)
      X[rr * kXS + part * 16 + k] = (v[k] - m) * rstd * __ldg(p.ln_g + part * 16 + k) + __ldg(p.ln_b + part * 16 + k);
  }
  asm volatile("cp.async.wait_group 0;" ::: "memory");
  __syncthreads();
  {
    // hidden = GELU(t x Wf1^T): 4 m tiles x 16 column tiles -> warp w: m tile w >> 1, column tiles (w & 1) * 8 .. + 7
    float acc[8][4];
    const int mt = warp >> 1, nt0 = (warp & 1) * 8;
    tile_tf32<8>(X, kXS, Wb, 128 + kWPad, 64, mt, nt0, lane, acc);
    __syncthreads();
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int c = (nt0 + j) * 8 + t * 2;
      *reinterpret_cast<float2*>(Y + (mt * 16 + g) * kYS + c) =
          make_float2(act_gelu_fast(acc[j][0]), act_gelu_fast(acc[j][1]));
      *reinterpret_cast<float2*>(Y + (mt * 16 + g + 8) * kYS + c) =
          make_float2(act_gelu_fast(acc[j][2]), act_gelu_fast(acc[j][3]));
    }
  }
  cp_filter(Wb, p.wf2t, 128, 64, tid);
  asm volatile("cp.async.commit_group;" ::: "memory");
  asm volatile("cp.async.wait_group 0;" ::: "memory");
  __syncthreads();
  {
    // z = hidden x Wf2^T + b -> 16-bit MMA operand rows
    float acc[4][4];
    const int mt = warp >> 1, nt0 = (warp & 1) * 4;
    tile_tf32<4>(Y, kYS, Wb, 64 + kWPad, 128, mt, nt0, lane, acc);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int c = (nt0 + j) * 8 + t * 2;
      const float bz0 = __ldg(p.bf2 + c), bz1 = __ldg(p.bf2 + c + 1);
      const int r0 = mt * 16 + g, r1 = r0 + 8;
      *reinterpret_cast<uint32_t*>(Hs + r0 * kHS + c) = r0 < R ? pack2_h16<F16>(acc[j][0] + bz0, acc[j][1] + bz1) : 0u;
      *reinterpret_cast<uint32_t*>(Hs + r1 * kHS + c) = r1 < R ? pack2_h16<F16>(acc[j][2] + bz0, acc[j][3] + bz1) : 0u;
    }
  }
  for (int e = tid; e < n * 64; e += 256) cs_s[e] = make_float2(__ldg(p.cos_t + e), __ldg(p.sin_t + e));
  __syncthreads();  // Hs complete; X / Y / Wb dead: the union switches to its back-end view

  // ---- back end (tensor cores): per head  qk = z x Wqk_h^T -> rotary -> K = q k^T ----
  auto load_head = [&](int h) {
    for (int e = tid; e < 256 * 8; e += 256) {
      const int row = e >> 3, ch = e & 7;  // rows 0..127: q_h, 128..255: k_h
      const uint16_t* src = p.wqk + ((int64_t)(row < 128 ? 0 : heads * 128) + h * 128 + (row & 127)) * 64 + ch * 8;
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(s32(Wq + row * kHS + ch * 8)), "l"(src) : "memory");
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };
  load_head(0);
  const int mtiles = (R + 15) >> 4;
  const int n16 = (n + 15) & ~15;
  const uint32_t Hs_a = s32(Hs), Wq_a = s32(Wq), Qs_a = s32(Qs), Ks_a = s32(Ks);
  for (int h = 0; h < heads; ++h) {
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    __syncthreads();  // Wq(h) landed; the previous head's q k^T reads of Qs / Ks are done
    // GEMM 1 + rotary: unit = (m tile, q|k, column pair block): accumulators for columns f and f + 64 sit in one thread.
    // A warp takes a CONTIGUOUS range of units, so the A fragments (and the rows' positions) are fetched once per m tile.
    {
      const int U = mtiles * 16;
      const int u_begin = (warp * U) >> 3, u_end = ((warp + 1) * U) >> 3;
      int cur_mt = -1;
      uint32_t a[4][4];
      int pos[2] = {0, 0};
      for (int unit = u_begin; unit < u_end; ++unit) {
        const int mt = unit >> 4, which = (unit >> 3) & 1, ntp = unit & 7;
        if (mt != cur_mt) {
          cur_mt = mt;
#pragma unroll
          for (int ks = 0; ks < 4; ++ks)
            ldsm4(Hs_a + (uint32_t)(((mt * 16 + (lane & 15)) * kHS + ks * 16 + (lane >> 4) * 8) * 2), a[ks][0], a[ks][1], a[ks][2], a[ks][3]);
          pos[0] = (mt * 16 + g) % n;      // position of the row inside its sample
          pos[1] = (mt * 16 + g + 8) % n;
        }
        float lo[4] = {0.f, 0.f, 0.f, 0.f}, hi[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) {
          uint32_t b0r, b1r;
          ldsm2(Wq_a + (uint32_t)(((which * 128 + ntp * 8 + (lane & 7)) * kHS + ks * 16 + ((lane >> 3) & 1) * 8) * 2), b0r, b1r);
          mma16816<F16>(lo, a[ks], b0r, b1r);
          ldsm2(Wq_a + (uint32_t)(((which * 128 + 64 + ntp * 8 + (lane & 7)) * kHS + ks * 16 + ((lane >> 3) & 1) * 8) * 2), b0r, b1r);
          mma16816<F16>(hi, a[ks], b0r, b1r);
        }
        const int f = ntp * 8 + t * 2;
        uint16_t* dst = which ? Ks : Qs;
#pragma unroll
        for (int half = 0; half < 2; ++half) {
          const int row = mt * 16 + g + half * 8;
          uint32_t o1 = 0u, o2 = 0u;
          if (row < R) {
            const float4 cs = *reinterpret_cast<const float4*>(cs_s + pos[half] * 64 + f);  // (cos f, sin f, cos f+1, sin f+1)
            const float a1x = lo[half * 2], a1y = lo[half * 2 + 1], a2x = hi[half * 2], a2y = hi[half * 2 + 1];
            o1 = pack2_h16<F16>(a1x * cs.x - a2x * cs.y, a1y * cs.z - a2y * cs.w);
            o2 = pack2_h16<F16>(a2x * cs.x + a1x * cs.y, a2y * cs.z + a1y * cs.w);
          }
          *reinterpret_cast<uint32_t*>(dst + row * kQS + f) = o1;
          *reinterpret_cast<uint32_t*>(dst + row * kQS + f + 64) = o2;
        }
      }
    }
    __syncthreads();  // Qs / Ks of this head complete, Wq free
    if (h + 1 < heads) load_head(h + 1);
    // GEMM 2: K[b][h] = q k^T * scaling per sample; unit = (sample, 16-row tile, 8-column tile)
    const int mts = n16 >> 4, nts = n16 >> 3;
    for (int unit = warp; unit < ns * mts * nts; unit += 8) {
      const int s = unit / (mts * nts), rem = unit - s * mts * nts;
      const int mt = rem / nts, nt = rem - mt * nts;
      const int base = s * n;
      float acc[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
      for (int ks = 0; ks < 8; ++ks) {
        uint32_t a[4], b0r, b1r;
        ldsm4(Qs_a + (uint32_t)(((base + mt * 16 + (lane & 15)) * kQS + ks * 16 + (lane >> 4) * 8) * 2), a[0], a[1], a[2], a[3]);
        ldsm2(Ks_a + (uint32_t)(((base + nt * 8 + (lane & 7)) * kQS + ks * 16 + ((lane >> 3) & 1) * 8) * 2), b0r, b1r);
        mma16816<F16>(acc, a, b0r, b1r);
      }
      float* Kg = p.K + (((int64_t)(b0 + s) * heads + h) * n) * n;
      const int i0 = mt * 16 + g, i1 = i0 + 8, j = nt * 8 + t * 2;
      if (i0 < n && j < n) Kg[i0 * n + j] = acc[0] * p.scaling;
      if (i0 < n && j + 1 < n) Kg[i0 * n + j + 1] = acc[1] * p.scaling;
      if (i1 < n && j < n) Kg[i1 * n + j] = acc[2] * p.scaling;
      if (i1 < n && j + 1 < n) Kg[i1 * n + j + 1] = acc[3] * p.scaling;
    }
  }
}

static size_t fa_axis_smem(int n) {
  size_t front = ((size_t)kAxRows * kXS + (size_t)kAxRows * kYS + (size_t)128 * (64 + kWPad)) * sizeof(float);  // 64 x 136 is smaller
  size_t back = ((size_t)256 * kHS + 2 * (size_t)kAxRowsAlloc * kQS) * 2;
  return (size_t)kAxRows * kHS * 2 + (size_t)n * 64 * 8 + (front > back ? front : back);
}

}  // namespace lns

extern "C" {

int lns_fa_axis_kernel_supported(int n, int dim, int hidden, int latent, int heads, int d) {
  return dim == 64 && hidden == 64 && latent == 64 && d == 128 && heads >= 1 && n >= 1 && n <= 64;
}

int lns_fa_axis_kernel(const float* pooled, int dtype16, int B, int n, int heads, const float* w1t, const float* ln_g,
                       const float* ln_b, float ln_eps, const float* wf1t, const float* wf2t, const float* bf2, const void* wqk16,
                       const float* cos_tab, const float* sin_tab, float scaling, float* K, void* stream) {
  LNS_REQUIRE(pooled && w1t && ln_g && ln_b && wf1t && wf2t && bf2 && wqk16 && cos_tab && sin_tab && K && B > 0,
              "lns_fa_axis_kernel: bad arguments");
  LNS_REQUIRE(lns::is_h16_host(dtype16), "lns_fa_axis_kernel: the to_qk filter must be LNS_BF16 or LNS_F16 (got %d)", dtype16);
  LNS_REQUIRE(lns_fa_axis_kernel_supported(n, 64, 64, 64, heads, 128), "lns_fa_axis_kernel: n=%d heads=%d not supported", n, heads);
  LNS_REQUIRE(((reinterpret_cast<uintptr_t>(pooled) | reinterpret_cast<uintptr_t>(w1t) | reinterpret_cast<uintptr_t>(wf1t) |
                reinterpret_cast<uintptr_t>(wf2t) | reinterpret_cast<uintptr_t>(wqk16)) & 15) == 0,
              "lns_fa_axis_kernel: pointers must be 16-byte aligned");
  lns::AxisParams p;
  p.pooled = pooled; p.B = B; p.n = n; p.heads = heads;
  p.S = lns::kAxRows / n;
  p.w1t = w1t; p.ln_g = ln_g; p.ln_b = ln_b; p.ln_eps = ln_eps; p.wf1t = wf1t; p.wf2t = wf2t; p.bf2 = bf2;
  p.wqk = reinterpret_cast<const uint16_t*>(wqk16);
  p.cos_t = cos_tab; p.sin_t = sin_tab; p.scaling = scaling; p.K = K;
  const size_t smem = lns::fa_axis_smem(n);
  {
    LNS_OPT_IN_SMEM((lns::fa_axis_kernel<false>), (int)lns::fa_axis_smem(64), "fa_axis");
    LNS_OPT_IN_SMEM((lns::fa_axis_kernel<true>), (int)lns::fa_axis_smem(64), "fa_axis");
  }
  const int grid = (B + p.S - 1) / p.S;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (dtype16 == LNS_F16) lns::fa_axis_kernel<true><<<grid, 256, smem, st>>>(p);
  else lns::fa_axis_kernel<false><<<grid, 256, smem, st>>>(p);
  return lns::check_launch("fa_axis_kernel");
}

}  // extern "C"
