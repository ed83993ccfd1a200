// SABlock (modules/basics.py:331-404) in ONE kernel:  out = x + proj( softmax(q k^T / sqrt(dh)) v ),  q|k|v = Linear(LN(x) + pe)
//
// Unfused: LayerNorm, one 128 -> 1536 GEMM that writes q|k|v (805 MB per 4096 samples of 64 tokens), the attention kernel that
// reads them back, the 512 -> 128 projection: 1.3 ms per decode chunk, HBM / epilogue bound.  Here a CTA owns the tokens of
// one or two samples (<= 128 rows) and walks the heads: the normalised tokens stay in shared memory, each head's q|k|v filter
// slice (48 KB) and projection slice (16 KB) stream in with cp.async one head ahead, q never leaves registers (the GEMM's
// accumulator fragments are re-packed as the next MMA's A operand), k and v live in shared memory for the flash-style
// softmax(q k^T) v of the warp's 16 query rows, and the head's output feeds the projection GEMM whose [rows x 128] fp32
// accumulator stays in registers across the heads.  Everything on mma.sync.m16n8k16 (bf16 or f16 operands, fp32 accumulate).
// HBM traffic per sample: 16 KB in + 16 KB out (+ the 0.5 MB of filters per CTA, L2 resident).
#include "common.cuh"

namespace lns {
namespace {
constexpr int kSaRows = 128;  // token rows per CTA (one sample of <= 128 tokens, or two of <= 64)
constexpr int kXS = 136;      // 16-bit row stride of the 128-wide operands (LN output, q|k|v filter rows)
constexpr int kHS = 72;       // 16-bit row stride of the 64-wide operands (k, v, projection filter rows)

__device__ __forceinline__ uint32_t sa_s32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void sa_ldsm4(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr));
}
__device__ __forceinline__ void sa_ldsm2(uint32_t addr, uint32_t& r0, uint32_t& r1) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x2.shared.b16 {%0,%1}, [%2];" : "=r"(r0), "=r"(r1) : "r"(addr));
}
__device__ __forceinline__ void sa_ldsm2t(uint32_t addr, uint32_t& r0, uint32_t& r1) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x2.trans.shared.b16 {%0,%1}, [%2];" : "=r"(r0), "=r"(r1) : "r"(addr));
}
template <bool F16>
__device__ __forceinline__ void sa_mma(float* c, const uint32_t* a, uint32_t b0, uint32_t b1) {
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

struct SaParams {
  const uint16_t* x;     // [B][n][128] 16-bit tokens (the block input; also the residual)
  uint16_t* y;           // [B][n][128]
  int B, n, heads, S;    // S samples per CTA (2 when n <= 64, else 1)
  const float* ln_g;     // [128]
  const float* ln_b;
  float ln_eps;
  const float* pe;       // [>= n][128] fp32 or NULL (added AFTER the norm, modules/basics.py:385-386)
  const uint16_t* wqkv;  // [3*heads*64][128] 16-bit: q rows (head, d) | k rows | v rows
  const float* bv;       // [heads*64] to_v bias or NULL
  const uint16_t* wproj; // [128][heads*64] 16-bit
  const float* bproj;    // [128] or NULL
  float scale_log2e;
};
}  // namespace

// grid ceil(B / S), block 256 (8 warps: warp w owns token rows 16w .. 16w+15 of the CTA)
template <bool F16>
__global__ void __launch_bounds__(256, 1) sablock_fused_kernel(const SaParams p) {
  extern __shared__ __align__(16) uint8_t sm[];
  uint16_t* xn_s = reinterpret_cast<uint16_t*>(sm);            // [128][kXS]  LN(x) + pe, later the output staging tile
  uint16_t* wh_s = xn_s + kSaRows * kXS;                       // [2][192][kXS]  q|k|v filter rows of a head
  uint16_t* wp_s = wh_s + 2 * 192 * kXS;                       // [2][128][kHS]  projection filter columns of a head
  uint16_t* k_s = wp_s + 2 * 128 * kHS;                        // [128][kHS]
  uint16_t* v_s = k_s + kSaRows * kHS;                         // [128][kHS]
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int g = lane >> 2, t = lane & 3;
  const int n = p.n, heads = p.heads, hd = heads * 64;
  const int b0 = blockIdx.x * p.S;
  const int ns = min(p.S, p.B - b0);
  const int rps = p.S == 2 ? 64 : 128;   // row slots per sample inside the CTA tile
  // this warp's 16 rows: sample `ws`, tokens tok0 .. tok0+15 of it
  const int ws = (warp * 16) / rps, tok0 = warp * 16 - ws * rps;
  const bool warp_on = ws < ns && tok0 < n;
  const int nk = (n + 63) & ~63;         // keys walked in blocks of 64 (padded rows are zero and masked)

  auto load_head = [&](int h, int buf) {
    uint16_t* wd = wh_s + buf * 192 * kXS;
    for (int e = tid; e < 192 * 16; e += 256) {   // 192 rows x 16 chunks of 16 B
      const int row = e >> 4, ch = e & 15;
      const uint16_t* src = p.wqkv + ((int64_t)(row >> 6) * hd + h * 64 + (row & 63)) * 128 + ch * 8;
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(sa_s32(wd + row * kXS + ch * 8)), "l"(src) : "memory");
    }
    uint16_t* pd = wp_s + buf * 128 * kHS;
    for (int e = tid; e < 128 * 8; e += 256) {    // 128 output rows x 8 chunks (this head's 64 input columns)
      const int row = e >> 3, ch = e & 7;
      const uint16_t* src = p.wproj + (int64_t)row * hd + h * 64 + ch * 8;
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(sa_s32(pd + row * kHS + ch * 8)), "l"(src) : "memory");
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };
  load_head(0, 0);

  // ---- LayerNorm (+ pe) of the CTA's rows -> xn_s (16-bit); warp w normalises rows w, w+8, ...: all 16 row loads are issued
  // before the first reduction (one dependent global load per row cost 16 exposed memory latencies per CTA) ----
  {
    uint2 raw[16];
    bool ok[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      const int r = warp + i * 8, s = r / rps, tok = r - s * rps;
      ok[i] = s < ns && tok < n;
      raw[i] = __ldg(reinterpret_cast<const uint2*>(p.x + ((int64_t)(b0 + (ok[i] ? s : 0)) * n + (ok[i] ? tok : 0)) * 128 + lane * 4));
    }
    const float4 gm = __ldg(reinterpret_cast<const float4*>(p.ln_g + lane * 4));
    const float4 bt = __ldg(reinterpret_cast<const float4*>(p.ln_b + lane * 4));
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      const int r = warp + i * 8, s = r / rps, tok = r - s * rps;
      uint32_t o0 = 0u, o1 = 0u;
      if (ok[i]) {  // warp uniform
        const float2 v01 = unpack2_h16<F16>(raw[i].x), v23 = unpack2_h16<F16>(raw[i].y);
        const float m = warp_sum((v01.x + v01.y) + (v23.x + v23.y)) * (1.f / 128.f);
        const float d0 = v01.x - m, d1 = v01.y - m, d2 = v23.x - m, d3 = v23.y - m;
        const float var = warp_sum(fmaf(d0, d0, fmaf(d1, d1, fmaf(d2, d2, d3 * d3)))) * (1.f / 128.f);
        const float rstd = rsqrtf(var + p.ln_eps);
        float4 pe4 = make_float4(0.f, 0.f, 0.f, 0.f);
        if (p.pe) pe4 = __ldg(reinterpret_cast<const float4*>(p.pe + (int64_t)tok * 128 + lane * 4));
        o0 = pack2_h16<F16>(fmaf(d0 * rstd, gm.x, bt.x) + pe4.x, fmaf(d1 * rstd, gm.y, bt.y) + pe4.y);
        o1 = pack2_h16<F16>(fmaf(d2 * rstd, gm.z, bt.z) + pe4.z, fmaf(d3 * rstd, gm.w, bt.w) + pe4.w);
      }
      *reinterpret_cast<uint2*>(xn_s + r * kXS + lane * 4) = make_uint2(o0, o1);
    }
  }

  float oacc[16][4];  // projection accumulator of this warp's 16 rows x 128 output channels, across the heads
#pragma unroll
  for (int i = 0; i < 16; ++i) oacc[i][0] = oacc[i][1] = oacc[i][2] = oacc[i][3] = 0.f;
  const uint32_t xn_a = sa_s32(xn_s), k_a = sa_s32(k_s), v_a = sa_s32(v_s);
  const int row_l = warp * 16 + (lane & 15);  // ldmatrix row of this lane inside the CTA tile

  for (int h = 0; h < heads; ++h) {
    const int buf = h & 1;
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    __syncthreads();  // head h's filters have landed; every warp is done with the previous head's k_s / v_s and filter buffers
    if (h + 1 < heads) load_head(h + 1, buf ^ 1);
    const uint32_t wh_a = sa_s32(wh_s + buf * 192 * kXS), wp_a = sa_s32(wp_s + buf * 128 * kHS);

    // ---- q | k | v of this warp's rows: [16 x 128] x [128 x 192] ----
    uint32_t qa[4][4];  // q as A fragments of the next GEMM (4 k-steps of 16 over d = 64)
    {
      uint32_t xa[8][4];
#pragma unroll
      for (int ks = 0; ks < 8; ++ks)
        sa_ldsm4(xn_a + (uint32_t)((row_l * kXS + ks * 16 + (lane >> 4) * 8) * 2), xa[ks][0], xa[ks][1], xa[ks][2], xa[ks][3]);
      float bvr[4][4];  // to_v bias of this thread's columns, fetched before the GEMMs that hide its latency
#pragma unroll
      for (int np = 0; np < 4; ++np) {
        const int col = h * 64 + np * 16 + t * 2;
        bvr[np][0] = p.bv ? __ldg(p.bv + col) : 0.f;
        bvr[np][1] = p.bv ? __ldg(p.bv + col + 1) : 0.f;
        bvr[np][2] = p.bv ? __ldg(p.bv + col + 8) : 0.f;
        bvr[np][3] = p.bv ? __ldg(p.bv + col + 9) : 0.f;
      }
#pragma unroll
      for (int which = 0; which < 3; ++which) {
#pragma unroll
        for (int nq = 0; nq < 2; ++nq) {   // two pairs of 8-column tiles at a time: four independent accumulator chains
          float c[4][4];
#pragma unroll
          for (int j = 0; j < 4; ++j) c[j][0] = c[j][1] = c[j][2] = c[j][3] = 0.f;
#pragma unroll
          for (int ks = 0; ks < 8; ++ks) {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              uint32_t b0r, b1r;
              sa_ldsm2(wh_a + (uint32_t)(((which * 64 + nq * 32 + j * 8 + (lane & 7)) * kXS + ks * 16 + ((lane >> 3) & 1) * 8) * 2), b0r, b1r);
              sa_mma<F16>(c[j], xa[ks], b0r, b1r);
            }
          }
#pragma unroll
          for (int hp = 0; hp < 2; ++hp) {  // accumulator fragments of two adjacent column tiles = one A fragment (16 x 16)
            const int np = nq * 2 + hp;
            const float* c0 = c[hp * 2];
            const float* c1 = c[hp * 2 + 1];
            if (which == 0) {
              qa[np][0] = pack2_h16<F16>(c0[0], c0[1]);
              qa[np][1] = pack2_h16<F16>(c0[2], c0[3]);
              qa[np][2] = pack2_h16<F16>(c1[0], c1[1]);
              qa[np][3] = pack2_h16<F16>(c1[2], c1[3]);
            } else {
              uint16_t* dst = which == 1 ? k_s : v_s;
              const int col = np * 16 + t * 2;
              const float bb0 = which == 2 ? bvr[np][0] : 0.f, bb1 = which == 2 ? bvr[np][1] : 0.f;
              const float bb2 = which == 2 ? bvr[np][2] : 0.f, bb3 = which == 2 ? bvr[np][3] : 0.f;
              // rows of tokens that do not exist stay zero (k: masked below; v: multiplied by zero probabilities)
              const bool v0 = warp_on && tok0 + g < n, v1 = warp_on && tok0 + g + 8 < n;
              const int r0 = warp * 16 + g, r1 = r0 + 8;
              *reinterpret_cast<uint32_t*>(dst + r0 * kHS + col) = v0 ? pack2_h16<F16>(c0[0] + bb0, c0[1] + bb1) : 0u;
              *reinterpret_cast<uint32_t*>(dst + r1 * kHS + col) = v1 ? pack2_h16<F16>(c0[2] + bb0, c0[3] + bb1) : 0u;
              *reinterpret_cast<uint32_t*>(dst + r0 * kHS + col + 8) = v0 ? pack2_h16<F16>(c1[0] + bb2, c1[1] + bb3) : 0u;
              *reinterpret_cast<uint32_t*>(dst + r1 * kHS + col + 8) = v1 ? pack2_h16<F16>(c1[2] + bb2, c1[3] + bb3) : 0u;
            }
          }
        }
      }
    }
    __syncthreads();  // k_s / v_s of every row of the CTA are complete

    // ---- softmax(q k^T) v over this sample's keys, 64 at a time (online softmax, exp2, fp32 statistics) ----
    float o[8][4];
#pragma unroll
    for (int i = 0; i < 8; ++i) o[i][0] = o[i][1] = o[i][2] = o[i][3] = 0.f;
    float m0 = -INFINITY, m1 = -INFINITY, l0 = 0.f, l1 = 0.f;
    const int kbase = ws * rps;  // first key row of this warp's sample inside k_s / v_s
    if (warp_on) {
      for (int kb = 0; kb < nk; kb += 64) {
        float sacc[8][4];
#pragma unroll
        for (int nt = 0; nt < 8; ++nt) {
          sacc[nt][0] = sacc[nt][1] = sacc[nt][2] = sacc[nt][3] = 0.f;
#pragma unroll
          for (int ks = 0; ks < 4; ++ks) {
            uint32_t b0r, b1r;
            sa_ldsm2(k_a + (uint32_t)(((kbase + kb + nt * 8 + (lane & 7)) * kHS + ks * 16 + ((lane >> 3) & 1) * 8) * 2), b0r, b1r);
            sa_mma<F16>(sacc[nt], qa[ks], b0r, b1r);
          }
        }
        float bm0 = -INFINITY, bm1 = -INFINITY;
#pragma unroll
        for (int nt = 0; nt < 8; ++nt) {
          const int key = kb + nt * 8 + t * 2;
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            float v = sacc[nt][j] * p.scale_log2e;
            if (key + (j & 1) >= n) v = -INFINITY;
            sacc[nt][j] = v;
          }
          bm0 = fmaxf(bm0, fmaxf(sacc[nt][0], sacc[nt][1]));
          bm1 = fmaxf(bm1, fmaxf(sacc[nt][2], sacc[nt][3]));
        }
        bm0 = fmaxf(bm0, __shfl_xor_sync(0xffffffffu, bm0, 1)); bm0 = fmaxf(bm0, __shfl_xor_sync(0xffffffffu, bm0, 2));
        bm1 = fmaxf(bm1, __shfl_xor_sync(0xffffffffu, bm1, 1)); bm1 = fmaxf(bm1, __shfl_xor_sync(0xffffffffu, bm1, 2));
        const float nm0 = fmaxf(m0, bm0), nm1 = fmaxf(m1, bm1);
        const float c0 = exp2f(m0 - nm0), c1 = exp2f(m1 - nm1);
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
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) {
          uint32_t pa[4];
          pa[0] = pack2_h16<F16>(sacc[2 * kk][0], sacc[2 * kk][1]);
          pa[1] = pack2_h16<F16>(sacc[2 * kk][2], sacc[2 * kk][3]);
          pa[2] = pack2_h16<F16>(sacc[2 * kk + 1][0], sacc[2 * kk + 1][1]);
          pa[3] = pack2_h16<F16>(sacc[2 * kk + 1][2], sacc[2 * kk + 1][3]);
#pragma unroll
          for (int nt = 0; nt < 8; ++nt) {
            uint32_t b0r, b1r;
            sa_ldsm2t(v_a + (uint32_t)(((kbase + kb + kk * 16 + (lane & 15)) * kHS + nt * 8) * 2), b0r, b1r);
            sa_mma<F16>(o[nt], pa, b0r, b1r);
          }
        }
      }
    }
    l0 += __shfl_xor_sync(0xffffffffu, l0, 1); l0 += __shfl_xor_sync(0xffffffffu, l0, 2);
    l1 += __shfl_xor_sync(0xffffffffu, l1, 1); l1 += __shfl_xor_sync(0xffffffffu, l1, 2);
    const float i0 = l0 > 0.f ? 1.f / l0 : 0.f, i1 = l1 > 0.f ? 1.f / l1 : 0.f;

    // ---- projection of this head's output: oacc += O_h [16 x 64] x Wproj[:, h*64 .. +63]^T ----
#pragma unroll
    for (int ks = 0; ks < 4; ++ks) {
      uint32_t oa[4];
      oa[0] = pack2_h16<F16>(o[2 * ks][0] * i0, o[2 * ks][1] * i0);
      oa[1] = pack2_h16<F16>(o[2 * ks][2] * i1, o[2 * ks][3] * i1);
      oa[2] = pack2_h16<F16>(o[2 * ks + 1][0] * i0, o[2 * ks + 1][1] * i0);
      oa[3] = pack2_h16<F16>(o[2 * ks + 1][2] * i1, o[2 * ks + 1][3] * i1);
#pragma unroll
      for (int nt = 0; nt < 16; ++nt) {
        uint32_t b0r, b1r;
        sa_ldsm2(wp_a + (uint32_t)(((nt * 8 + (lane & 7)) * kHS + ks * 16 + ((lane >> 3) & 1) * 8) * 2), b0r, b1r);
        sa_mma<F16>(oacc[nt], oa, b0r, b1r);
      }
    }
  }

  // ---- out = oacc + bias + x (residual, fp32 add) -> staged in xn_s (its last reader was the last head's q|k|v GEMM) ----
  __syncthreads();
  {
    const int r0 = warp * 16 + g, r1 = r0 + 8;
    const bool v0 = warp_on && tok0 + g < n, v1 = warp_on && tok0 + g + 8 < n;
    const uint16_t* x0 = p.x + ((int64_t)(b0 + ws) * n + tok0 + g) * 128;
    const uint16_t* x1 = x0 + 8 * 128;
#pragma unroll
    for (int nt = 0; nt < 16; ++nt) {
      const int col = nt * 8 + t * 2;
      const float bp0 = p.bproj ? __ldg(p.bproj + col) : 0.f, bp1 = p.bproj ? __ldg(p.bproj + col + 1) : 0.f;
      if (v0) {
        const float2 r = unpack2_h16<F16>(__ldg(reinterpret_cast<const uint32_t*>(x0 + col)));
        *reinterpret_cast<uint32_t*>(xn_s + r0 * kXS + col) = pack2_h16<F16>(oacc[nt][0] + bp0 + r.x, oacc[nt][1] + bp1 + r.y);
      }
      if (v1) {
        const float2 r = unpack2_h16<F16>(__ldg(reinterpret_cast<const uint32_t*>(x1 + col)));
        *reinterpret_cast<uint32_t*>(xn_s + r1 * kXS + col) = pack2_h16<F16>(oacc[nt][2] + bp0 + r.x, oacc[nt][3] + bp1 + r.y);
      }
    }
  }
  __syncthreads();
  for (int e = tid; e < kSaRows * 16; e += 256) {  // 16 lanes x 16 B per token row: full 256-byte rows
    const int r = e >> 4, ch = e & 15;
    const int s = r / rps, tok = r - s * rps;
    if (s < ns && tok < n)
      *reinterpret_cast<uint4*>(p.y + ((int64_t)(b0 + s) * n + tok) * 128 + ch * 8) = *reinterpret_cast<const uint4*>(xn_s + r * kXS + ch * 8);
  }
}

static size_t sablock_smem() {
  return ((size_t)kSaRows * kXS + 2 * 192 * (size_t)kXS + 2 * 128 * (size_t)kHS + 2 * (size_t)kSaRows * kHS) * 2;
}

}  // namespace lns

extern "C" {

int lns_sablock_fused_supported(int n, int dim, int heads, int dim_head) {
  return dim == 128 && dim_head == 64 && heads >= 1 && n >= 1 && n <= 128;
}

int lns_sablock_fused(const void* x, int dtype, int B, int n, int heads, const float* ln_g, const float* ln_b, float ln_eps,
                      const float* pe, const void* wqkv16, const float* bv, const void* wproj16, const float* bproj, float scale,
                      void* y, void* stream) {
  LNS_REQUIRE(x && y && ln_g && ln_b && wqkv16 && wproj16 && B > 0, "lns_sablock_fused: bad arguments");
  LNS_REQUIRE(lns::is_h16_host(dtype), "lns_sablock_fused: x/y must be LNS_BF16 or LNS_F16 (got %d)", dtype);
  LNS_REQUIRE(lns_sablock_fused_supported(n, 128, heads, 64), "lns_sablock_fused: n=%d heads=%d not supported", n, heads);
  LNS_REQUIRE(((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(y) | reinterpret_cast<uintptr_t>(wqkv16) |
                reinterpret_cast<uintptr_t>(wproj16) | reinterpret_cast<uintptr_t>(ln_g) | reinterpret_cast<uintptr_t>(ln_b) |
                reinterpret_cast<uintptr_t>(pe)) & 15) == 0,
              "lns_sablock_fused: pointers must be 16-byte aligned");
  lns::SaParams p;
  p.x = reinterpret_cast<const uint16_t*>(x);
  p.y = reinterpret_cast<uint16_t*>(y);
  p.B = B; p.n = n; p.heads = heads;
  p.S = n <= 64 ? 2 : 1;
  p.ln_g = ln_g; p.ln_b = ln_b; p.ln_eps = ln_eps; p.pe = pe;
  p.wqkv = reinterpret_cast<const uint16_t*>(wqkv16);
  p.bv = bv;
  p.wproj = reinterpret_cast<const uint16_t*>(wproj16);
  p.bproj = bproj;
  p.scale_log2e = scale * 1.4426950408889634f;
  const size_t smem = lns::sablock_smem();
  {
    LNS_OPT_IN_SMEM((lns::sablock_fused_kernel<false>), 227 * 1024, "sablock_fused");
    LNS_OPT_IN_SMEM((lns::sablock_fused_kernel<true>), 227 * 1024, "sablock_fused");
  }
  const int grid = (B + p.S - 1) / p.S;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (dtype == LNS_F16) lns::sablock_fused_kernel<true><<<grid, 256, smem, st>>>(p);
  else lns::sablock_fused_kernel<false><<<grid, 256, smem, st>>>(p);
  return lns::check_launch("sablock_fused_kernel");
}

}  // extern "C"
