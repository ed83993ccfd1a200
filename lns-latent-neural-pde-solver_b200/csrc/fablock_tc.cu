// FABlock2D (modules/factorized_attention.py:144-159) with EVERY contraction on tcgen05 -- the successor of fablock_full.cu, whose
// in_proj and axial contractions ran as mma.sync phases at ~4x their tensor floor (33 % of the NS2d rollout at 12 % of the tensor peak).
//
// Same mathematics and the same folding as fablock_full.cu (GroupNorm folded into the in_proj filter, InstanceNorm folded into
// to_out[1], the to_out accumulator of the CTA's pixels resident in tensor memory across the heads), restructured so that each
// phase is one batch of tcgen05.mma instructions followed by a TMEM -> shared-memory drain:
//   A  u_phi_h = u x Ws^T                         A = pixel rows (K-major), B = in_proj slice (K-major)
//   B  contraction over image rows with Kx[h]     A = blockdiag(Kx) [128 x 128] (K-major), B = the pixel rows as an MN-MAJOR operand
//   C  contraction over image columns with Ky[h]  A = blockdiag(Ky), B = pixel rows (MN-major)
//   E  acc += v_h x W1'^T                         A = pixel rows (K-major), B = folded to_out[1] slice
// The pixel-row image -- one 128-byte row of 64 channels per pixel, 16-byte chunks XOR-ed with (row & 7) -- is BOTH a K-major A
// operand (phases A, E) and, read through an MN-major descriptor (instruction-descriptor bit 16; tools/umma_mn_probe.cu: SBO = 1024,
// K step of 16 pixel rows = +2048 B), the B operand of the axial contractions: out[(line, i)][c] = sum_j K[i][j] u[(line, j)][c] is a
// GEMM over PIXELS once the lines of a 128-row tile are stacked and K is block-diagonal.  Phase B wants lines = image columns,
// phase C lines = image rows: the drain of B writes its rows transposed, so no separate transpose pass exists.
//
// Tensor memory: the to_out accumulator needs H*W x 64 fp32 = all 512 columns at 32x32, leaving nothing for the phases.  A 32x32
// sample is therefore split over a CLUSTER OF TWO CTAs: CTA r owns image columns [16r, 16r+16) in phases A-B and image rows
// [16r, 16r+16) in phases C-E; the drain of phase B stores each transposed row into the shared memory of the CTA that owns its
// image row (st.shared::cluster), and the InstanceNorm partial sums are exchanged the same way.  Each CTA holds 256 accumulator +
// 256 scratch columns.  A 16x16 sample is one CTA (128 + 128 columns), no exchange.
#include <cuda_runtime.h>

#include "common.cuh"

namespace lns {

namespace tptx {
__device__ __forceinline__ uint32_t s32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "LNST_WAIT_%=:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra LNST_DONE_%=;\n\t"
      "bra LNST_WAIT_%=;\n\t"
      "LNST_DONE_%=:\n\t"
      "}" ::"r"(bar),
      "r"(parity)
      : "memory");
}
// generic-proxy shared-memory writes (this CTA's, or a peer's that a cluster barrier has made visible) -> async proxy (tcgen05.mma).
// (The unqualified `fence.proxy.async` compiles to MEMBAR.ALL.GPU + the fence: 7 % of the kernel's stall samples in the first profile.)
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tmem_alloc(uint32_t smem_dst, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_dst), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tmem_relinquish() { asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void umma_lohi(uint32_t d_tmem, uint32_t a_lo, uint32_t a_hi, uint32_t b_lo, uint32_t b_hi, uint32_t idesc,
                                          uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      ".reg .b64 da, db;\n\t"
      "setp.ne.b32 p, %6, 0;\n\t"
      "mov.b64 da, {%1, %2};\n\t"
      "mov.b64 db, {%3, %4};\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %5, p;\n\t"
      "}" ::"r"(d_tmem),
      "r"(a_lo), "r"(a_hi), "r"(b_lo), "r"(b_hi), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t* r) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void st_shared_v4(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
// address of `local_addr` (a shared::cta address of THIS CTA's window) in CTA `rank` of the cluster
__device__ __forceinline__ uint32_t map_to_cta(uint32_t local_addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(local_addr), "r"(rank));
  return r;
}
__device__ __forceinline__ void st_cluster_v4(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared::cluster.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
__device__ __forceinline__ void st_cluster_f32(uint32_t addr, float v) {
  asm volatile("st.shared::cluster.f32 [%0], %1;" ::"r"(addr), "f"(v) : "memory");
}
__device__ __forceinline__ void cluster_arrive() { asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory"); }
__device__ __forceinline__ void cluster_wait() { asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory"); }
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
}  // namespace tptx

// -DLNS_TC_TRACE: thread 0 of CTA 0 records clock64() at every stage boundary of the first heads into p.trace (tools only)
#ifdef LNS_TC_TRACE
#define TC_MARK(id) do { if (blockIdx.x == 0 && tid == 0 && p.trace && h < 4) p.trace[h * 16 + (id)] = clock64(); } while (0)
#else
#define TC_MARK(id) do { } while (0)
#endif

namespace {
struct TcParams {
  const uint16_t* u;       // [B][n][n][64] 16-bit
  uint16_t* out;           // [B][n][n][64]
  const float* gn_scale;   // [B][64]
  const float* gn_shift;   // [B][64]
  const float* w_in;       // [heads*64][64]
  const float* Kx;         // [B][heads][n][n]
  const float* Ky;         // [B][heads][n][n]
  const float* w_out1;     // [64][heads*64]
  const float* w_out2;     // [64][64]
  float eps;
  int heads;
  long long* trace;        // LNS_TC_TRACE builds only
};

__device__ __forceinline__ uint32_t row_off(int r, int chunk) { return (uint32_t)r * 128u + (uint32_t)((chunk ^ (r & 7)) << 4); }

// sum over the 32 lanes of v[0..63]; afterwards lane L holds the totals of channels 2L and 2L+1 in v[0], v[1]
__device__ __forceinline__ void warp_reduce_transpose64(float (&v)[64], int lane) {
#pragma unroll
  for (int off = 16, n = 32; off >= 1; off >>= 1, n >>= 1) {
    const bool upper = (lane & off) != 0;
#pragma unroll
    for (int j = 0; j < 32; ++j) {
      if (j < n) {
        const float send = upper ? v[j] : v[j + n];
        const float keep = upper ? v[j + n] : v[j];
        v[j] = keep + __shfl_xor_sync(0xffffffffu, send, off);
      }
    }
  }
}

// N = line length (16: one CTA per sample; 32: cluster of two), CL = CTAs per sample, F16 = IEEE half (else bf16)
template <int N, int CL, bool F16>
__device__ __forceinline__ void fablock_tc_body(const TcParams& p) {
  constexpr int NPX = 16 * N;        // pixels per CTA: 16 lines of N
  constexpr int T = NPX / 128;       // 128-row MMA tiles per CTA
  constexpr int LPT = 128 / N;       // lines per tile
  constexpr int NTHR = 128 * T;      // one warp per (tile, TMEM lane quadrant)
  constexpr int NW = NTHR / 32;
  constexpr uint32_t kBuf = (uint32_t)NPX * 128u;
  // shared memory
  constexpr uint32_t oP = 0, oQ = kBuf, oBD0 = 2 * kBuf, oBD1 = oBD0 + 32768u, oWs = oBD1 + 32768u, oWo = oWs + 8192u,
                     oRed = oWo + 8192u, oBias = oRed + (uint32_t)NW * 128u * 4u, oObias = oBias + 256u, oStat = oObias + 256u,
                     oGn = oStat + 512u, oPart = oGn + 512u, oBar = oPart + 1024u, oSlot = oBar + 64u;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (tptx::s32(smem_raw) + 1023u) & ~1023u;
  uint8_t* gen = smem_raw + (base - tptx::s32(smem_raw));
  float* red_s = reinterpret_cast<float*>(gen + oRed);     // [NW][64][2]
  float* bias_s = reinterpret_cast<float*>(gen + oBias);   // in_proj bias of this head (GroupNorm shift folded)
  float* obias_s = reinterpret_cast<float*>(gen + oObias); // to_out[1] bias accumulated over the heads (InstanceNorm means folded)
  float* stat_s = reinterpret_cast<float*>(gen + oStat);   // [64][2] rstd, -mean*rstd
  float* gn_s = reinterpret_cast<float*>(gen + oGn);       // [128] scale | shift
  float* part_s = reinterpret_cast<float*>(gen + oPart);   // [2][128] own | peer partial (sum, sumsq) of this head
  volatile uint32_t* slot_gen = reinterpret_cast<volatile uint32_t*>(gen + oSlot);
  const uint32_t barA = base + oBar, barB = barA + 8, barC = barA + 16, barE = barA + 24;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int b = blockIdx.x / CL;
  const uint32_t rank = CL > 1 ? tptx::cluster_ctarank() : 0u;
  const int heads = p.heads, C = heads * 64;
  const uint16_t* ub = p.u + (int64_t)b * N * N * 64;
  constexpr uint32_t kTmemCols = 2 * T * 64;  // scratch [0, T*64) | to_out accumulator [T*64, 2*T*64)

  if (tid == 0) {
    // one MMA-issuing thread per 128-row tile (lane 0 of the tile's first warp): the T instruction streams run in parallel, every
    // phase barrier collects T tcgen05.commit arrivals
    tptx::mbar_init(barA, T);
    tptx::mbar_init(barB, T);
    tptx::mbar_init(barC, T);
    tptx::mbar_init(barE, T);
    tptx::fence_mbar_init();
  }
  if (warp == 0) {
    tptx::tmem_alloc(base + oSlot, kTmemCols);
    tptx::tmem_relinquish();
  }
  if (tid < 128) gn_s[tid] = tid < 64 ? __ldg(p.gn_scale + (int64_t)b * 64 + tid) : __ldg(p.gn_shift + (int64_t)b * 64 + tid - 64);
  if (tid < 64) obias_s[tid] = 0.f;

  // ---- helpers -------------------------------------------------------------------------------------------------------------
  // raw input of this CTA's 16 image COLUMNS, line-major: row xl*N + y <- pixel (y, x = 16*rank + xl)
  auto load_raw = [&](uint32_t dstbuf) {
    const int ch = tid & 7;
    for (int r = tid >> 3; r < NPX; r += NTHR / 8) {
      const int xl = r / N, y = r - xl * N;
      const uint16_t* src = ub + ((int64_t)y * N + (int)rank * 16 * (CL > 1) + xl) * 64 + ch * 8;
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dstbuf + row_off(r, ch)), "l"(src) : "memory");
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };
  // in_proj slice of head hh with the GroupNorm scale folded in (B operand [64 n][64 k], K-major) + its bias (shift folded)
  auto build_ws = [&](int hh) {
    const float* wsrc = p.w_in + (int64_t)hh * 64 * 64;
    for (int e = tid; e < 64 * 8; e += NTHR) {
      const int n = e >> 3, kc = e & 7;
      const float4 w0 = __ldg(reinterpret_cast<const float4*>(wsrc + n * 64 + kc * 8));
      const float4 w1 = __ldg(reinterpret_cast<const float4*>(wsrc + n * 64 + kc * 8 + 4));
      const float4 s0 = *reinterpret_cast<const float4*>(gn_s + kc * 8);
      const float4 s1 = *reinterpret_cast<const float4*>(gn_s + kc * 8 + 4);
      const float4 t0 = *reinterpret_cast<const float4*>(gn_s + 64 + kc * 8);
      const float4 t1 = *reinterpret_cast<const float4*>(gn_s + 64 + kc * 8 + 4);
      tptx::st_shared_v4(base + oWs + row_off(n, kc), pack2_h16<F16>(w0.x * s0.x, w0.y * s0.y), pack2_h16<F16>(w0.z * s0.z, w0.w * s0.w),
                         pack2_h16<F16>(w1.x * s1.x, w1.y * s1.y), pack2_h16<F16>(w1.z * s1.z, w1.w * s1.w));
      float bsum = w0.x * t0.x + w0.y * t0.y + w0.z * t0.z + w0.w * t0.w + w1.x * t1.x + w1.y * t1.y + w1.z * t1.z + w1.w * t1.w;
      bsum += __shfl_xor_sync(0xffffffffu, bsum, 1);  // the 8 k-chunks of output channel n sit in 8 consecutive lanes
      bsum += __shfl_xor_sync(0xffffffffu, bsum, 2);
      bsum += __shfl_xor_sync(0xffffffffu, bsum, 4);
      if (kc == 0) bias_s[n] = bsum;
    }
  };
  // block-diagonal [128 x 128] of the N x N kernel Kmat (fp32, row-major [i][j]) as a K-major A operand in two 64-column slabs
  auto build_bd = [&](uint32_t dst, const float* Kmat) {
    constexpr int NIT = 128 * 16 / NTHR;  // 16-byte chunks per thread; all global loads are issued before the first store
    float4 ka[NIT], kb[NIT];
#pragma unroll
    for (int q = 0; q < NIT; ++q) {
      const int e = tid + q * NTHR, r = e >> 4, kc = e & 15;
      const int line = r / N, i = r - line * N;
      const int kline = (kc * 8) / N, j0 = kc * 8 - kline * N;
      ka[q] = kb[q] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (kline == line) {
        ka[q] = __ldg(reinterpret_cast<const float4*>(Kmat + i * N + j0));
        kb[q] = __ldg(reinterpret_cast<const float4*>(Kmat + i * N + j0 + 4));
      }
    }
#pragma unroll
    for (int q = 0; q < NIT; ++q) {
      const int e = tid + q * NTHR, r = e >> 4, kc = e & 15;
      tptx::st_shared_v4(dst + (uint32_t)(kc >> 3) * 16384u + row_off(r, kc & 7), pack2_h16<F16>(ka[q].x, ka[q].y), pack2_h16<F16>(ka[q].z, ka[q].w),
                         pack2_h16<F16>(kb[q].x, kb[q].y), pack2_h16<F16>(kb[q].z, kb[q].w));
    }
  };
  // descriptor halves: K-major SWIZZLE_128B (8-row groups 1024 B apart) -- also the MN-major form (SBO = 1024, LBO unused)
  const uint32_t desc_hi = (1024u >> 4) | (1u << 14) | (2u << 29);
  auto lo16 = [](uint32_t addr) { return (addr & 0x3FFFFu) >> 4; };
  const uint32_t fmt = F16 ? 0u : ((1u << 7) | (1u << 10));
  const uint32_t idesc_kk = (1u << 4) | fmt | ((64u >> 3) << 17) | ((128u >> 4) << 24);  // A, B K-major
  const uint32_t idesc_kmn = idesc_kk | (1u << 16);                                        // B MN-major
  // TMEM -> registers: the 64 fp32 columns `col0 ..` of this thread's lane
  auto ld_row = [&](uint32_t tmem, int col0, int quad, float (&v)[64]) {
    uint32_t raw[32];
#pragma unroll
    for (int half = 0; half < 2; ++half) {
      tptx::tmem_ld32(tmem + (uint32_t)(col0 + half * 32) + ((uint32_t)(quad * 32) << 16), raw);
      tptx::tmem_ld_wait();
#pragma unroll
      for (int j = 0; j < 32; ++j) v[half * 32 + j] = __uint_as_float(raw[j]);
    }
  };

  // ---- prologue: head 0 operands -----------------------------------------------------------------------------------------
  load_raw(base + oP);
  tptx::tc_fence_before();
  __syncthreads();  // gn_s, barriers, TMEM slot
  tptx::tc_fence_after();
  const uint32_t tmem = *slot_gen;
  const uint32_t tmem_acc = tmem + (uint32_t)(T * 64);
  build_ws(0);
  build_bd(base + oBD0, p.Kx + ((int64_t)b * heads + 0) * N * N);
  build_bd(base + oBD1, p.Ky + ((int64_t)b * heads + 0) * N * N);
  if (CL > 1) {  // both CTAs exist and have initialised their barriers / buffers before anybody stores remotely
    tptx::cluster_arrive();
    tptx::cluster_wait();
  }

  const int tile = warp >> 2, quad = warp & 3;
  const int m = quad * 32 + lane;  // accumulator row of this thread in its tile
  const bool issuer = quad == 0 && lane == 0;  // issues this tile's MMAs

  for (int h = 0; h < heads; ++h) {
    const uint32_t X = base + ((h & 1) ? oQ : oP), Y = base + ((h & 1) ? oP : oQ);
    const uint32_t par = (uint32_t)(h & 1);
    TC_MARK(0);
    // ---- S1: raw(h) has landed in X; Ws / bias / BD0 / BD1 of head h are written ----
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    tptx::fence_proxy_async();
    tptx::tc_fence_before();
    __syncthreads();
    TC_MARK(1);
    // ---- phase A: u_phi = raw x Ws^T -> scratch ----
    if (issuer) {
      tptx::tc_fence_after();
      const uint32_t b_lo = lo16(base + oWs);
      const uint32_t a_lo = lo16(X + (uint32_t)tile * 16384u);
#pragma unroll
      for (int k = 0; k < 4; ++k) tptx::umma_lohi(tmem + (uint32_t)(tile * 64), a_lo + 2u * k, desc_hi, b_lo + 2u * k, desc_hi, idesc_kk, k != 0);
      tptx::umma_commit(barA);
    }
    __syncwarp();
    tptx::mbar_wait(barA, par);
    tptx::tc_fence_after();
    TC_MARK(2);
    {  // drain A: + bias -> 16-bit -> the same rows of X (the MMAs that read them have completed)
      float v[64];
      ld_row(tmem, tile * 64, quad, v);
      const int r = tile * 128 + m;
#pragma unroll
      for (int c8 = 0; c8 < 8; ++c8) {
        uint32_t pk[4];
#pragma unroll
        for (int j = 0; j < 4; ++j)
          pk[j] = pack2_h16<F16>(v[c8 * 8 + 2 * j] + bias_s[c8 * 8 + 2 * j], v[c8 * 8 + 2 * j + 1] + bias_s[c8 * 8 + 2 * j + 1]);
        tptx::st_shared_v4(X + row_off(r, c8), pk[0], pk[1], pk[2], pk[3]);
      }
    }
    tptx::fence_proxy_async();
    tptx::tc_fence_before();
    __syncthreads();  // S2
    TC_MARK(3);
    // ---- phase B: contraction over image rows (lines = image columns): scratch = blockdiag(Kx) x X ----
    if (issuer) {
      tptx::tc_fence_after();
#pragma unroll
      for (int ks = 0; ks < 8; ++ks) {
        const uint32_t a_lo = lo16(base + oBD0 + (uint32_t)(ks >> 2) * 16384u) + 2u * (ks & 3);
        const uint32_t b_lo = lo16(X + (uint32_t)tile * 16384u + (uint32_t)ks * 2048u);
        tptx::umma_lohi(tmem + (uint32_t)(tile * 64), a_lo, desc_hi, b_lo, desc_hi, idesc_kmn, ks != 0);
      }
      tptx::umma_commit(barB);
    }
    __syncwarp();
    // while the tensor core works: the next head's in_proj slice (Ws / bias are no longer read: barA has completed, drain A is over)
    if (h + 1 < heads) build_ws(h + 1);
    TC_MARK(4);
    // Y is the buffer phase E of the previous head read (its V rows): free once that GEMM has completed -- here and in the peer
    if (h > 0) {
      tptx::mbar_wait(barE, (uint32_t)((h - 1) & 1));
      tptx::tc_fence_after();
    }
    if (CL > 1) {
      tptx::cluster_arrive();  // #a: "my Y may be written"
      tptx::cluster_wait();
    }
    tptx::mbar_wait(barB, par);
    tptx::tc_fence_after();
    TC_MARK(5);
    {  // drain B: row (line xl, image row i) -> row il*N + x of the buffer Y of the CTA that owns image row i (transposed)
      float v[64];
      ld_row(tmem, tile * 64, quad, v);
      const int xl = tile * LPT + m / N, i = m % N;
      const int x = (CL > 1 ? (int)rank * 16 : 0) + xl;
      const uint32_t owner = CL > 1 ? (uint32_t)(i >> 4) : 0u;
      const int il = CL > 1 ? (i & 15) : i;
      const int r = il * N + x;
      const uint32_t dst_local = Y + (uint32_t)r * 128u;
      const bool remote = CL > 1 && owner != rank;
      const uint32_t dst = remote ? tptx::map_to_cta(dst_local, owner) : dst_local;
#pragma unroll
      for (int c8 = 0; c8 < 8; ++c8) {
        uint32_t pk[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) pk[j] = pack2_h16<F16>(v[c8 * 8 + 2 * j], v[c8 * 8 + 2 * j + 1]);
        const uint32_t a = dst + (uint32_t)((c8 ^ (r & 7)) << 4);
        if (remote) tptx::st_cluster_v4(a, pk[0], pk[1], pk[2], pk[3]);
        else tptx::st_shared_v4(a, pk[0], pk[1], pk[2], pk[3]);
      }
    }
    tptx::fence_proxy_async();
    tptx::tc_fence_before();
    if (CL > 1) {
      tptx::cluster_arrive();  // #b: every transposed row of both CTAs has landed
      tptx::cluster_wait();
    } else {
      __syncthreads();
    }
    tptx::fence_proxy_async();  // (the peer's stores came through the generic proxy)
    TC_MARK(6);
    // ---- phase C: contraction over image columns (lines = image rows): scratch = blockdiag(Ky) x Y ----
    if (issuer) {
      tptx::tc_fence_after();
#pragma unroll
      for (int ks = 0; ks < 8; ++ks) {
        const uint32_t a_lo = lo16(base + oBD1 + (uint32_t)(ks >> 2) * 16384u) + 2u * (ks & 3);
        const uint32_t b_lo = lo16(Y + (uint32_t)tile * 16384u + (uint32_t)ks * 2048u);
        tptx::umma_lohi(tmem + (uint32_t)(tile * 64), a_lo, desc_hi, b_lo, desc_hi, idesc_kmn, ks != 0);
      }
      tptx::umma_commit(barC);
    }
    __syncwarp();
    // while the tensor core works: the next head's row kernel (BD0 is no longer read: barB has completed)
    if (h + 1 < heads) build_bd(base + oBD0, p.Kx + ((int64_t)b * heads + h + 1) * N * N);
    tptx::mbar_wait(barC, par);
    tptx::tc_fence_after();
    TC_MARK(7);
    {  // drain C: 16-bit V rows into X (free: phase B has read it) + per-channel (sum, sum of squares) of the ROUNDED values.
       // Two passes over the thread's TMEM row (a second tcgen05.ld is cheaper than keeping 128 values live at 512 threads).
      float v[64];
      ld_row(tmem, tile * 64, quad, v);
      const int r = tile * 128 + m;
#pragma unroll
      for (int c8 = 0; c8 < 8; ++c8) {
        uint32_t pk[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          pk[j] = pack2_h16<F16>(v[c8 * 8 + 2 * j], v[c8 * 8 + 2 * j + 1]);
          const float2 back = unpack2_h16<F16>(pk[j]);
          v[c8 * 8 + 2 * j] = back.x;
          v[c8 * 8 + 2 * j + 1] = back.y;
        }
        tptx::st_shared_v4(X + row_off(r, c8), pk[0], pk[1], pk[2], pk[3]);
      }
      warp_reduce_transpose64(v, lane);
      red_s[(warp * 64 + 2 * lane) * 2 + 0] = v[0];
      red_s[(warp * 64 + 2 * lane + 1) * 2 + 0] = v[1];
      ld_row(tmem, tile * 64, quad, v);
#pragma unroll
      for (int j = 0; j < 64; j += 2) {
        const float2 back = unpack2_h16<F16>(pack2_h16<F16>(v[j], v[j + 1]));
        v[j] = back.x * back.x;
        v[j + 1] = back.y * back.y;
      }
      warp_reduce_transpose64(v, lane);
      red_s[(warp * 64 + 2 * lane) * 2 + 1] = v[0];
      red_s[(warp * 64 + 2 * lane + 1) * 2 + 1] = v[1];
    }
    tptx::tc_fence_before();
    __syncthreads();  // S4: Y and BD1 are free (phase C has completed), V rows and the partial statistics are written
    TC_MARK(8);
    if (h + 1 < heads) load_raw(Y);  // next head's raw input (Y is next head's X): lands during the statistics / phase E
    // this thread's share of to_out[1]'s slice (folded below): in flight across the statistics exchange
    float4 wo0[(64 * 8 + NTHR - 1) / NTHR], wo1[(64 * 8 + NTHR - 1) / NTHR];
#pragma unroll
    for (int q = 0; q < (64 * 8 + NTHR - 1) / NTHR; ++q) {
      const int e = tid + q * NTHR, n = e >> 3, kc = e & 7;
      wo0[q] = __ldg(reinterpret_cast<const float4*>(p.w_out1 + (int64_t)n * C + h * 64 + kc * 8));
      wo1[q] = __ldg(reinterpret_cast<const float4*>(p.w_out1 + (int64_t)n * C + h * 64 + kc * 8 + 4));
    }
    if (tid < 128) {
      const int c = tid >> 1, which = tid & 1;
      float s = 0.f;
#pragma unroll 4
      for (int w = 0; w < NW; ++w) s += red_s[(w * 64 + c) * 2 + which];
      part_s[tid] = s;
      if (CL > 1) tptx::st_cluster_f32(tptx::map_to_cta(base + oPart + 512u + (uint32_t)tid * 4u, rank ^ 1u), s);
    }
    if (CL > 1) {
      tptx::cluster_arrive();  // #c: partial statistics exchanged
      tptx::cluster_wait();
    } else {
      __syncthreads();
    }
    TC_MARK(9);
    if (tid < 64) {
      double sm = (double)part_s[tid * 2], ss = (double)part_s[tid * 2 + 1];
      if (CL > 1) {
        sm += (double)part_s[128 + tid * 2];
        ss += (double)part_s[128 + tid * 2 + 1];
      }
      const double inv_n = 1.0 / (double)(N * N);
      const double mean = sm * inv_n;
      double var = ss * inv_n - mean * mean;
      if (var < 0.0) var = 0.0;
      const double ve = var + (double)p.eps;
      double rstd = (double)rsqrtf((float)ve);
      rstd = rstd * (1.5 - 0.5 * ve * rstd * rstd);
      rstd = rstd * (1.5 - 0.5 * ve * rstd * rstd);
      stat_s[tid * 2 + 0] = (float)rstd;
      stat_s[tid * 2 + 1] = (float)(-mean * rstd);
    }
    __syncthreads();  // S5
    TC_MARK(10);
    // ---- phase E: fold the normalisation into to_out[1]'s slice, acc += V x W1'^T ----
#pragma unroll
    for (int q = 0; q < (64 * 8 + NTHR - 1) / NTHR; ++q) {
      const int e = tid + q * NTHR, n = e >> 3, kc = e & 7;
      const float4 w0 = wo0[q], w1 = wo1[q];
      const float wv[8] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w};
      float sc[8];
      float bsum = 0.f;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        sc[j] = wv[j] * stat_s[(kc * 8 + j) * 2 + 0];
        bsum = fmaf(wv[j], stat_s[(kc * 8 + j) * 2 + 1], bsum);
      }
      tptx::st_shared_v4(base + oWo + row_off(n, kc), pack2_h16<F16>(sc[0], sc[1]), pack2_h16<F16>(sc[2], sc[3]), pack2_h16<F16>(sc[4], sc[5]),
                         pack2_h16<F16>(sc[6], sc[7]));
      bsum += __shfl_xor_sync(0xffffffffu, bsum, 1);
      bsum += __shfl_xor_sync(0xffffffffu, bsum, 2);
      bsum += __shfl_xor_sync(0xffffffffu, bsum, 4);
      if (kc == 0) obias_s[n] += bsum;
    }
    tptx::fence_proxy_async();
    tptx::tc_fence_before();
    __syncthreads();  // S6
    if (issuer) {
      tptx::tc_fence_after();
      const uint32_t b_lo = lo16(base + oWo);
      const uint32_t a_lo = lo16(X + (uint32_t)tile * 16384u);
#pragma unroll
      for (int k = 0; k < 4; ++k)
        tptx::umma_lohi(tmem_acc + (uint32_t)(tile * 64), a_lo + 2u * k, desc_hi, b_lo + 2u * k, desc_hi, idesc_kk, (h | k) != 0);
      tptx::umma_commit(barE);
    }
    __syncwarp();
    TC_MARK(11);
    // while phase E runs: the next head's column kernel (BD1 is free since barC)
    if (h + 1 < heads) build_bd(base + oBD1, p.Ky + ((int64_t)b * heads + h + 1) * N * N);
  }

  // ================= after the last head: GELU(acc + b1') -> to_out[3] -> + skip -> out =================
  const uint32_t Yl = base + (((heads - 1) & 1) ? oP : oQ);  // free: the last phase C has read it, no next head was loaded
  // to_out[3] filter -> K-major B operand in the (now unused) BD0 region
  for (int e = tid; e < 64 * 8; e += NTHR) {
    const int n = e >> 3, kc = e & 7;
    const float4 w0 = __ldg(reinterpret_cast<const float4*>(p.w_out2 + n * 64 + kc * 8));
    const float4 w1 = __ldg(reinterpret_cast<const float4*>(p.w_out2 + n * 64 + kc * 8 + 4));
    tptx::st_shared_v4(base + oBD0 + row_off(n, kc), pack2_h16<F16>(w0.x, w0.y), pack2_h16<F16>(w0.z, w0.w), pack2_h16<F16>(w1.x, w1.y),
                       pack2_h16<F16>(w1.z, w1.w));
  }
  tptx::mbar_wait(barE, (uint32_t)((heads - 1) & 1));
  tptx::tc_fence_after();
  {
    float v[64];
    ld_row(tmem_acc, tile * 64, quad, v);
    const int r = tile * 128 + m;
#pragma unroll
    for (int c8 = 0; c8 < 8; ++c8) {
      uint32_t pk[4];
#pragma unroll
      for (int j = 0; j < 4; ++j)
        pk[j] = pack2_h16<F16>(act_gelu_fast(v[c8 * 8 + 2 * j] + obias_s[c8 * 8 + 2 * j]), act_gelu_fast(v[c8 * 8 + 2 * j + 1] + obias_s[c8 * 8 + 2 * j + 1]));
      tptx::st_shared_v4(Yl + row_off(r, c8), pk[0], pk[1], pk[2], pk[3]);
    }
  }
  tptx::fence_proxy_async();
  tptx::tc_fence_before();
  __syncthreads();
  if (issuer) {
    tptx::tc_fence_after();
    const uint32_t b_lo = lo16(base + oBD0);
    const uint32_t a_lo = lo16(Yl + (uint32_t)tile * 16384u);
#pragma unroll
    for (int k = 0; k < 4; ++k) tptx::umma_lohi(tmem + (uint32_t)(tile * 64), a_lo + 2u * k, desc_hi, b_lo + 2u * k, desc_hi, idesc_kk, k != 0);
    tptx::umma_commit(barA);
  }
  __syncwarp();
  tptx::mbar_wait(barA, (uint32_t)(heads & 1));
  tptx::tc_fence_after();
  {  // + skip (the block's raw input, fp32 add) -> 16-bit -> global: accumulator row r = (image row il, column x) of this CTA's rows
    float v[64];
    ld_row(tmem, tile * 64, quad, v);
    const int r = tile * 128 + m;
    const int il = r / N, x = r - il * N;
    const int64_t pix = (int64_t)((CL > 1 ? (int)rank * 16 : 0) + il) * N + x;
    const uint4* skip = reinterpret_cast<const uint4*>(ub + pix * 64);
    uint4* dst = reinterpret_cast<uint4*>(p.out + ((int64_t)b * N * N + pix) * 64);
    uint4 sk[8];
#pragma unroll
    for (int c8 = 0; c8 < 8; ++c8) sk[c8] = __ldg(skip + c8);
#pragma unroll
    for (int c8 = 0; c8 < 8; ++c8) {
      const uint32_t sw[4] = {sk[c8].x, sk[c8].y, sk[c8].z, sk[c8].w};
      uint32_t o[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float2 s2 = unpack2_h16<F16>(sw[j]);
        o[j] = pack2_h16<F16>(v[c8 * 8 + 2 * j] + s2.x, v[c8 * 8 + 2 * j + 1] + s2.y);
      }
      dst[c8] = make_uint4(o[0], o[1], o[2], o[3]);
    }
  }
  tptx::tc_fence_before();
  if (CL > 1) {  // no CTA of the cluster may exit while its peer can still address its shared memory
    tptx::cluster_arrive();
    tptx::cluster_wait();
  } else {
    __syncthreads();
  }
  if (warp == 0) {
    tptx::tc_fence_after();
    tptx::tmem_dealloc(tmem, kTmemCols);
  }
}

constexpr size_t tc_smem_bytes(int N) {
  const size_t npx = 16 * (size_t)N, nw = npx / 32;
  return 2 * npx * 128 + 2 * 32768 + 8192 + 8192 + nw * 128 * 4 + 256 + 256 + 512 + 512 + 1024 + 64 + 64 + 1024;
}
}  // namespace

template <bool F16>
__global__ void __launch_bounds__(256, 1) fablock_tc16_kernel(const TcParams p) { fablock_tc_body<16, 1, F16>(p); }
template <bool F16>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(512, 1) fablock_tc32_kernel(const TcParams p) { fablock_tc_body<32, 2, F16>(p); }

}  // namespace lns

extern "C" {

int lns_fablock_tc_supported(int H, int W, int dim, int dim_head, int dim_out) {
  return dim == 64 && dim_head == 64 && dim_out == 64 && H == W && (H == 16 || H == 32);
}

int lns_fablock_tc(const void* u, int dtype, int B, int H, int W, int heads, const float* gn_scale, const float* gn_shift,
                   const float* w_in_proj, const float* Kx, const float* Ky, float eps, const float* w_out1, const float* w_out2,
                   void* out, void* stream) {
  LNS_REQUIRE(u && gn_scale && gn_shift && w_in_proj && Kx && Ky && w_out1 && w_out2 && out && B > 0 && heads > 0,
              "lns_fablock_tc: bad arguments");
  LNS_REQUIRE(lns::is_h16_host(dtype), "lns_fablock_tc: u/out must be LNS_BF16 or LNS_F16 (got dtype %d)", dtype);
  LNS_REQUIRE(lns_fablock_tc_supported(H, W, 64, 64, 64), "lns_fablock_tc: %dx%d is not covered (16x16 and 32x32 are; use lns_fablock_full)", H, W);
  LNS_REQUIRE((reinterpret_cast<uintptr_t>(u) & 15) == 0 && (reinterpret_cast<uintptr_t>(out) & 15) == 0 &&
                  (reinterpret_cast<uintptr_t>(w_in_proj) & 15) == 0 && (reinterpret_cast<uintptr_t>(w_out1) & 15) == 0 &&
                  (reinterpret_cast<uintptr_t>(w_out2) & 15) == 0 && (reinterpret_cast<uintptr_t>(Kx) & 15) == 0 &&
                  (reinterpret_cast<uintptr_t>(Ky) & 15) == 0,
              "lns_fablock_tc: pointers must be 16-byte aligned");
  lns::TcParams p;
  p.u = reinterpret_cast<const uint16_t*>(u);
  p.out = reinterpret_cast<uint16_t*>(out);
  p.gn_scale = gn_scale; p.gn_shift = gn_shift; p.w_in = w_in_proj; p.Kx = Kx; p.Ky = Ky;
  p.w_out1 = w_out1; p.w_out2 = w_out2; p.eps = eps; p.heads = heads;
  p.trace = nullptr;
#ifdef LNS_TC_TRACE
  {
    static long long* dbg = nullptr;
    if (!dbg) cudaMalloc(&dbg, 64 * sizeof(long long));
    p.trace = dbg;
    cudaMemsetAsync(dbg, 0, 64 * sizeof(long long), reinterpret_cast<cudaStream_t>(stream));
  }
#endif
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const bool f16 = dtype == LNS_F16;
  if (H == 16) {
    const int smem = (int)lns::tc_smem_bytes(16);
    if (f16) {
      LNS_OPT_IN_SMEM((lns::fablock_tc16_kernel<true>), smem, "fablock_tc");
      lns::fablock_tc16_kernel<true><<<B, 256, smem, st>>>(p);
    } else {
      LNS_OPT_IN_SMEM((lns::fablock_tc16_kernel<false>), smem, "fablock_tc");
      lns::fablock_tc16_kernel<false><<<B, 256, smem, st>>>(p);
    }
  } else {
    const int smem = (int)lns::tc_smem_bytes(32);
    if (f16) {
      LNS_OPT_IN_SMEM((lns::fablock_tc32_kernel<true>), smem, "fablock_tc");
      lns::fablock_tc32_kernel<true><<<2 * B, 512, smem, st>>>(p);
    } else {
      LNS_OPT_IN_SMEM((lns::fablock_tc32_kernel<false>), smem, "fablock_tc");
      lns::fablock_tc32_kernel<false><<<2 * B, 512, smem, st>>>(p);
    }
  }
#ifdef LNS_TC_TRACE
  {
    long long hbuf[64];
    cudaStreamSynchronize(st);
    cudaMemcpy(hbuf, p.trace, sizeof(hbuf), cudaMemcpyDeviceToHost);
    static int printed = 0;
    if (printed++ < 2)
      for (int h = 0; h < 4; ++h) {
        fprintf(stderr, "fablock_tc trace head %d (%dx%d):", h, H, W);
        for (int i = 1; i < 12; ++i) fprintf(stderr, " s%d->%d %lld", i - 1, i, hbuf[h * 16 + i] - hbuf[h * 16 + i - 1]);
        if (h < 3) fprintf(stderr, " | to next head %lld", hbuf[(h + 1) * 16] - hbuf[h * 16 + 11]);
        fprintf(stderr, "\n");
      }
  }
#endif
  return lns::check_launch("fablock_tc_kernel");
}

}  // extern "C"
