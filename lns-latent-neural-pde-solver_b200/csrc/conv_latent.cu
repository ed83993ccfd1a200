// tcgen05 implicit-GEMM 3x3 convolution for the LATENT GRID of the propagator (LNS_ENGINE_LATENT): 128 -> 128 channels on
// 8x8 circular samples, dilation 1 or 2 (train_stage2_ns2d.py:34-41: the nine 3x3 convs of every propagator step, and the
// coarse 8x8 stage of the NS2d decoder).
//
// Why a third engine: on these layers the gather engine (conv_umma.cu) moves 590 KB from L2 per 128-pixel tile (every input
// pixel once per tap + the 295 KB filter once per tile) against 4.6k tensor cycles -- L2 -> SM bound at 20 % of the tensor
// peak (322 TFLOP/s, bench.py roofline_by_time), and the halo engine (conv_halo.cu) cannot keep a 128-channel filter resident.
// Here one CTA takes a SUPER TILE of 4 samples (2 MMA tiles of 2 samples each):
//   * the four circular halos ((8+2d)^2 pixels x 256 B each, 102 / 147 KB) are loaded ONCE and stay in shared memory for
//     all nine taps: tap (ky,kx) is a descriptor view of the same bytes, as in the halo engine;
//   * a 128-row MMA tile holds TWO samples.  Their halo rows are interleaved -- pixel (s, hy, hx) of the tile sits in row
//     hy*(2*HWd) + s*HWd + hx -- so that the sixteen 8-row groups (y, s) of a tap view are equally spaced (HWd rows): one
//     SWIZZLE_128B descriptor with SBO = HWd*128 B covers both samples.  Accumulator row m = (y = m>>4, s = (m>>3)&1, x = m&7);
//   * the filter streams through a 3-stage ring in (tap, 64-channel slab) blocks of 16 KB (one cp.async.bulk each) and every
//     block feeds BOTH MMA tiles: L2 -> SM traffic per 128-pixel tile 590 KB -> (102 + 295) / 2 = 199 KB.
// Warp roles (320 threads): warps 0-3 halo producers, warp 4 MMA issuer + TMEM owner, warps 5-8 epilogue, warp 9 filter
// ring producer.  Four 128-column accumulators: the epilogue of super tile i overlaps tile i+1.
#include <stdlib.h>

#include "common.cuh"

namespace lns {

namespace lptx {
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "LNSL_WAIT_%=:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra LNSL_DONE_%=;\n\t"
      "bra LNSL_WAIT_%=;\n\t"
      "LNSL_DONE_%=:\n\t"
      "}" ::"r"(bar),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src, uint32_t src_bytes) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async_arrive_noinc(uint32_t bar) {
  asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src),
               "r"(bytes), "r"(bar)
               : "memory");
}
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
__device__ __forceinline__ void umma_f16(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}" ::"r"(d_tmem),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
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
}  // namespace lptx

namespace {
constexpr int kLatThreads = 320;
constexpr int kBMax = 8;     // most filter-ring stages (barrier slots); the launch uses as many as fit, at least 3
constexpr uint32_t kBBytes = 128u * 128u;  // one (tap, slab) filter block: 128 output channels x 64 input channels x 2 B

__device__ __forceinline__ uint64_t desc_sbo(uint32_t smem_addr, uint32_t sbo_bytes) {
  uint64_t d = (uint64_t)((smem_addr & 0x3FFFFu) >> 4);
  d |= (uint64_t)(sbo_bytes >> 4) << 32;
  d |= 1ull << 46;
  d |= 2ull << 61;
  return d;
}

struct LatentParams {
  const uint16_t* x;   // [B][8][8][128] 16-bit, batch stride x_bstride elements
  const uint16_t* w;   // packed [tap 9][slab 2][128][64] swizzled (LNS_W_UMMA_BF16 / _F16)
  const float* bias;   // [128] or NULL
  int act;
  const void* residual;
  int res_dtype;
  int64_t res_bstride;
  void* y;
  int y_dtype;
  int64_t x_bstride, y_bstride;
  int B, dil, nsuper;
  int HWd;             // halo width = height = 8 + 2*dil
  int plane_bytes;     // one (tile, slab) halo plane: HWd * 2 * HWd rows x 128 B (a multiple of 1024)
  int x_f16;
  int bstages;         // filter ring depth (3 .. kBMax)
  float inv_hwd;
};
}  // namespace

// Epilogue of the four epilogue warps, compiled twice: RES16 = a 16-bit residual travels through the staging tile (see below).
// (Two instantiations, not a runtime flag: the eight parked residual vectors of the RES16 path otherwise cost the GELU
// loop of the residual-free layers its registers -- 44 -> 59 us per launch when both shared one body.)
template <bool RES16>
__device__ __forceinline__ void latent_epilogue(const LatentParams& p, uint32_t tmem_acc, uint32_t stage_out, uint32_t acc_full0,
                                                uint32_t acc_empty0, int warp, int lane) {
  auto acc_full = [&](int a) { return acc_full0 + 8u * a; };
  auto acc_empty = [&](int a) { return acc_empty0 + 8u * a; };
  const int quad = warp & 3;  // TMEM lane quadrant this warp may access
  const int m = quad * 32 + lane;
  const int yo = m >> 4, s_of = (m >> 3) & 1, xo = m & 7;
  const bool y16 = is_h16(p.y_dtype);
  const uint32_t my_stage = stage_out + (uint32_t)(warp - 5) * 4096u;
  const uint32_t my_row_st = my_stage + (uint32_t)lane * 128u;
  const int rd_row = lane >> 3, rd_chunk = lane & 7;
  // 16-bit residual: the warp's 32 rows x 128 B of the NEXT (tile, channel group) are fetched one step ahead with
  // coalesced 16-byte loads (8 lanes per pixel row: full 128-byte lines) and parked in registers; they go through the
  // output staging tile (written before the accumulator is read, each thread then reads ITS row and overwrites it with the
  // result).  Reading the residual row by row straight from global memory -- 16 dependent 8-byte loads per thread, 32
  // lines per instruction -- made the `res` layers 45% slower than the same layer without a residual.
  uint4 rq[8];
  auto res_issue = [&](int sup_, int t_, int cg_) {
#pragma unroll
    for (int pass = 0; pass < 8; ++pass) {
      const int mm = quad * 32 + pass * 4 + rd_row;
      const int bb = sup_ * 4 + t_ * 2 + ((mm >> 3) & 1);
      rq[pass] = make_uint4(0u, 0u, 0u, 0u);
      if (bb < p.B)
        rq[pass] = __ldg(reinterpret_cast<const uint4*>(reinterpret_cast<const uint16_t*>(p.residual) + (int64_t)bb * p.res_bstride +
                                                        (int64_t)((mm >> 4) * 8 + (mm & 7)) * 128 + cg_ + rd_chunk * 8));
    }
  };
  if (RES16 && (int)blockIdx.x < p.nsuper) res_issue(blockIdx.x, 0, 0);
  int it = 0;
  for (int sup = blockIdx.x; sup < p.nsuper; sup += gridDim.x, ++it) {
    const int a = it & 1;
    lptx::mbar_wait(acc_full(a), (uint32_t)((it >> 1) & 1));
    lptx::tc_fence_after();
    for (int t = 0; t < 2; ++t) {
      const int b = sup * 4 + t * 2 + s_of;
      const bool row_ok = b < p.B;
      const int64_t pix = yo * 8 + xo;
      const int64_t yrow = (int64_t)b * p.y_bstride + pix * 128;
      const int64_t rrow = (int64_t)b * p.res_bstride + pix * 128;
      const uint32_t t_lane = tmem_acc + (uint32_t)((a * 2 + t) * 128) + ((uint32_t)(quad * 32) << 16);
      for (int cg = 0; cg < 128; cg += 64) {
        if (RES16) {
          // (the staging tile is free: the previous step's read-back ended with a __syncwarp)
#pragma unroll
          for (int pass = 0; pass < 8; ++pass) {
            const int r = pass * 4 + rd_row;
            asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(my_stage + (uint32_t)r * 128u + (uint32_t)((rd_chunk ^ (r & 7)) << 4)),
                         "r"(rq[pass].x), "r"(rq[pass].y), "r"(rq[pass].z), "r"(rq[pass].w) : "memory");
          }
          __syncwarp();
          if (cg == 0) res_issue(sup, t, 64);
          else if (t == 0) res_issue(sup, 1, 0);
          else if (sup + (int)gridDim.x < p.nsuper) res_issue(sup + (int)gridDim.x, 0, 0);
        }
#pragma unroll
        for (int cc = 0; cc < 64; cc += 32) {
          const int c0 = cg + cc;
          uint32_t raw[32];
          __syncwarp();
          lptx::tmem_ld32(t_lane + (uint32_t)c0, raw);
          lptx::tmem_ld_wait();
          float v[32];
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(raw[j]);
          if (p.bias) {
#pragma unroll
            for (int j = 0; j < 32; j += 4) {
              const float4 bb = __ldg(reinterpret_cast<const float4*>(p.bias + c0 + j));
              v[j] += bb.x; v[j + 1] += bb.y; v[j + 2] += bb.z; v[j + 3] += bb.w;
            }
          }
          if (p.act != LNS_ACT_NONE) {
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = apply_act_fast(v[j], p.act);
          }
          if (RES16) {
#pragma unroll
            for (int h4 = 0; h4 < 4; ++h4) {
              uint32_t w0, w1, w2, w3;
              asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(w0), "=r"(w1), "=r"(w2), "=r"(w3)
                           : "r"(my_row_st + (uint32_t)((((cc >> 3) + h4) ^ (lane & 7)) << 4)));
              const uint32_t ww[4] = {w0, w1, w2, w3};
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                const float2 f = (p.res_dtype == LNS_F16) ? unpack2_h16<true>(ww[j]) : unpack2_h16<false>(ww[j]);
                v[h4 * 8 + 2 * j] += f.x;
                v[h4 * 8 + 2 * j + 1] += f.y;
              }
            }
          } else if (row_ok && p.residual) {
#pragma unroll
            for (int j = 0; j < 32; j += 4) {
              const float4 rr = ld4_as_float(p.residual, p.res_dtype, rrow + c0 + j);
              v[j] += rr.x; v[j + 1] += rr.y; v[j + 2] += rr.z; v[j + 3] += rr.w;
            }
          }
          if (y16) {
#pragma unroll
            for (int h4 = 0; h4 < 4; ++h4) {
              uint32_t pk[4];
              if (p.y_dtype == LNS_F16) {
#pragma unroll
                for (int j = 0; j < 4; ++j) pk[j] = pack2_h16<true>(v[h4 * 8 + 2 * j], v[h4 * 8 + 2 * j + 1]);
              } else {
#pragma unroll
                for (int j = 0; j < 4; ++j) pk[j] = pack2_h16<false>(v[h4 * 8 + 2 * j], v[h4 * 8 + 2 * j + 1]);
              }
              const int ch = (cc >> 3) + h4;
              asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(my_row_st + (uint32_t)((ch ^ (lane & 7)) << 4)),
                           "r"(pk[0]), "r"(pk[1]), "r"(pk[2]), "r"(pk[3]) : "memory");
            }
          } else if (row_ok) {
            float4* dst = reinterpret_cast<float4*>(reinterpret_cast<float*>(p.y) + yrow + c0);
            if (p.y_dtype == LNS_TF32) {
#pragma unroll
              for (int j = 0; j < 32; ++j) v[j] = round_tf32(v[j]);
            }
#pragma unroll
            for (int j = 0; j < 8; ++j) dst[j] = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
          }
        }
        if (y16) {
          __syncwarp();
          // staged row r of this warp = accumulator row quad*32 + r: 8 consecutive rows = the 8 pixels of image row
          // (y, s); each pass moves 4 rows x 128 B (one 64-channel group of 4 pixels: full 128-byte lines)
#pragma unroll
          for (int pass = 0; pass < 8; ++pass) {
            const int r = pass * 4 + rd_row;
            const int mm = quad * 32 + r;
            const int bb = sup * 4 + t * 2 + ((mm >> 3) & 1);
            uint32_t w0, w1, w2, w3;
            asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(w0), "=r"(w1), "=r"(w2), "=r"(w3)
                         : "r"(my_stage + (uint32_t)r * 128u + (uint32_t)((rd_chunk ^ (r & 7)) << 4)));
            if (bb < p.B)
              *reinterpret_cast<uint4*>(reinterpret_cast<uint16_t*>(p.y) + (int64_t)bb * p.y_bstride +
                                        (int64_t)((mm >> 4) * 8 + (mm & 7)) * 128 + cg + rd_chunk * 8) = make_uint4(w0, w1, w2, w3);
          }
          __syncwarp();
        }
      }
    }
    lptx::tc_fence_before();
    lptx::mbar_arrive(acc_empty(a));
  }
}

__global__ void __launch_bounds__(kLatThreads, 1) conv_latent_kernel(const LatentParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (lptx::smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* gen = smem_raw + (base - lptx::smem_u32(smem_raw));
  const uint32_t halo_a = base;                                         // [tile 2][slab 2] planes
  const int BS = p.bstages;
  const uint32_t b_ring = halo_a + 4u * (uint32_t)p.plane_bytes;        // BS x 16 KB
  const uint32_t stage_out = b_ring + (uint32_t)BS * kBBytes;             // 4 warps x 4 KB output staging
  const uint32_t bar_base = stage_out + 4u * 4096u;
  const uint32_t halo_full = bar_base, halo_empty = bar_base + 8;
  auto b_full = [&](int s) { return bar_base + 16u + 8u * s; };
  auto b_empty = [&](int s) { return bar_base + 16u + 8u * (kBMax + s); };
  auto acc_full = [&](int a) { return bar_base + 16u + 8u * (2 * kBMax + a); };
  auto acc_empty = [&](int a) { return bar_base + 16u + 8u * (2 * kBMax + 2 + a); };
  const uint32_t tmem_slot = bar_base + 16u + 8u * (2 * kBMax + 4);
  volatile uint32_t* tmem_slot_gen = reinterpret_cast<volatile uint32_t*>(gen + (tmem_slot - base));
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int d = p.dil, HWd = p.HWd;

  if (tid == 0) {
    lptx::mbar_init(halo_full, 128);
    lptx::mbar_init(halo_empty, 1);
    for (int s = 0; s < kBMax; ++s) {
      lptx::mbar_init(b_full(s), 1);
      lptx::mbar_init(b_empty(s), 1);
    }
    for (int a = 0; a < 2; ++a) {
      lptx::mbar_init(acc_full(a), 1);
      lptx::mbar_init(acc_empty(a), 128);
    }
    lptx::fence_mbar_init();
  }
  if (warp == 4) {
    lptx::tmem_alloc(tmem_slot, 512);
    lptx::tmem_relinquish();
  }
  lptx::tc_fence_before();
  __syncthreads();
  lptx::tc_fence_after();
  const uint32_t tmem_acc = *tmem_slot_gen;

  if (warp < 4) {
    // ============================== halo producers (128 threads) ==============================
    const int c16 = tid & 15;            // 16-byte chunk of the pixel's 256 B: slab = c16 >> 3, chunk = c16 & 7
    const int slab = c16 >> 3, chunk = c16 & 7;
    const int npx = HWd * HWd;
    int it = 0;
    for (int sup = blockIdx.x; sup < p.nsuper; sup += gridDim.x, ++it) {
      if (it > 0) lptx::mbar_wait(halo_empty, (uint32_t)((it - 1) & 1));  // the MMAs of the previous super tile have read the halos
      for (int ts = 0; ts < 4; ++ts) {                                     // sample ts = tile (ts >> 1), slot s = ts & 1
        const int b = sup * 4 + ts;
        const bool bok = b < p.B;
        const uint16_t* xb = p.x + (int64_t)(bok ? b : 0) * p.x_bstride + c16 * 8;
        const uint32_t plane = halo_a + (uint32_t)((ts >> 1) * 2 + slab) * (uint32_t)p.plane_bytes;
        for (int q = tid >> 4; q < npx; q += 8) {
          const int hy = __float2int_rd(((float)q + 0.5f) * p.inv_hwd), hx = q - hy * HWd;
          const int ys = (hy - d) & 7, xs = (hx - d) & 7;                  // circular wrap on the 8 x 8 grid
          const int r = hy * (2 * HWd) + (ts & 1) * HWd + hx;
          lptx::cp_async16(plane + (uint32_t)r * 128u + (uint32_t)((chunk ^ (r & 7)) << 4), xb + (ys * 8 + xs) * 128, bok ? 16u : 0u);
        }
      }
      lptx::cp_async_arrive_noinc(halo_full);
    }
  } else if (warp == 4) {
    // ============================== MMA issuer ==============================
    if (lane == 0) {
      const uint32_t idesc = (1u << 4) | (p.x_f16 ? 0u : ((1u << 7) | (1u << 10))) | ((128u >> 3) << 17) | ((128u >> 4) << 24);
      const uint32_t sbo = (uint32_t)HWd * 128u;
      int s = 0, it = 0;
      uint32_t bphase = 0;  // filter ring position: stage s, phase bit flips on every wrap
      for (int sup = blockIdx.x; sup < p.nsuper; sup += gridDim.x, ++it) {
        const int a = it & 1;
        if (it >= 2) lptx::mbar_wait(acc_empty(a), (uint32_t)(((it >> 1) - 1) & 1));
        lptx::mbar_wait(halo_full, (uint32_t)(it & 1));
        lptx::fence_proxy_async();
        lptx::tc_fence_after();
        for (int kb = 0; kb < 18; ++kb) {
          lptx::mbar_wait(b_full(s), bphase);
          lptx::tc_fence_after();
          const int tap = kb >> 1, sl = kb & 1, ky = tap / 3, kx = tap - ky * 3;
          const uint64_t bdesc = desc_sbo(b_ring + (uint32_t)s * kBBytes, 1024u);
          const uint32_t view = (uint32_t)((ky * d) * (2 * HWd) + kx * d) * 128u;
#pragma unroll
          for (int t = 0; t < 2; ++t) {
            const uint64_t adesc = desc_sbo(halo_a + (uint32_t)(t * 2 + sl) * (uint32_t)p.plane_bytes + view, sbo);
            const uint32_t d_tmem = tmem_acc + (uint32_t)((a * 2 + t) * 128);
#pragma unroll
            for (int k = 0; k < 4; ++k)
              lptx::umma_f16(d_tmem, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), idesc, (kb | k) != 0 ? 1u : 0u);
          }
          lptx::umma_commit(b_empty(s));
          if (++s == BS) {
            s = 0;
            bphase ^= 1u;
          }
        }
        lptx::umma_commit(halo_empty);
        lptx::umma_commit(acc_full(a));
      }
    }
    __syncwarp();
  } else if (warp == 9) {
    // ============================== filter ring producer (one thread) ==============================
    // the 18 (tap, slab) blocks of every super tile through the ring; always the same 295 KB: L2 hits.  Runs ahead of the
    // MMAs by the ring depth, independent of the halo producers.
    if (lane == 0) {
      int s = 0;
      uint32_t wrap = 0;  // completed trips around the ring
      for (int sup = blockIdx.x; sup < p.nsuper; sup += gridDim.x) {
        for (int kb = 0; kb < 18; ++kb) {
          if (wrap > 0) lptx::mbar_wait(b_empty(s), (wrap - 1u) & 1u);
          lptx::mbar_expect_tx(b_full(s), kBBytes);
          lptx::bulk_g2s(b_ring + (uint32_t)s * kBBytes, p.w + (int64_t)kb * 128 * 64, kBBytes, b_full(s));
          if (++s == BS) {
            s = 0;
            ++wrap;
          }
        }
      }
    }
    __syncwarp();
  } else {
    // ============================== epilogue (warps 5-8) ==============================
    if (p.residual != nullptr && is_h16(p.res_dtype)) latent_epilogue<true>(p, tmem_acc, stage_out, acc_full(0), acc_empty(0), warp, lane);
    else latent_epilogue<false>(p, tmem_acc, stage_out, acc_full(0), acc_empty(0), warp, lane);
  }

  lptx::tc_fence_before();
  __syncthreads();
  if (warp == 4) {
    lptx::tc_fence_after();
    lptx::tmem_dealloc(tmem_acc, 512);
  }
}

bool conv_latent_supported(const LnsConvDesc* d) {
  return d->KH == 3 && d->KW == 3 && d->stride == 1 && d->Cin == 128 && d->Cout == 128 && d->Hin == 8 && d->Win == 8 &&
         d->Hv == 8 && d->Wv == 8 && d->Hout == 8 && d->Wout == 8 && (d->dil == 1 || d->dil == 2) && d->pad_t == d->dil &&
         d->pad_l == d->dil && d->pad_mode_h == LNS_PAD_CIRCULAR && d->pad_mode_w == LNS_PAD_CIRCULAR &&
         is_h16_host(d->x_dtype) && d->x_layout == LNS_NHWC && d->y_layout == LNS_NHWC && d->pro_scale == nullptr &&
         d->pro_act == LNS_ACT_NONE && d->sample_bias == nullptr && d->pre_add == nullptr &&
         d->w_format == (d->x_dtype == LNS_F16 ? LNS_W_UMMA_F16 : LNS_W_UMMA_BF16);
}

int conv2d_latent(const LnsConvDesc* d, cudaStream_t stream) {
  LNS_REQUIRE(conv_latent_supported(d),
              "lns_conv2d(latent): needs a 3x3 stride-1 circular conv 128 -> 128 on 8x8 samples, dilation 1|2, NHWC bf16/f16 "
              "input, UMMA-packed weights, no prologue / sample bias / pre-add");
  LNS_REQUIRE(d->x_bstride % 8 == 0 && d->y_bstride % 8 == 0, "lns_conv2d(latent): batch strides must be multiples of 8");
  LNS_REQUIRE((reinterpret_cast<uintptr_t>(d->x) & 15) == 0 && (reinterpret_cast<uintptr_t>(d->y) & 15) == 0 &&
                  (reinterpret_cast<uintptr_t>(d->w) & 15) == 0, "lns_conv2d(latent): x, y, w must be 16-byte aligned");
  if (d->residual) LNS_REQUIRE(d->res_bstride % 8 == 0 && (reinterpret_cast<uintptr_t>(d->residual) & 15) == 0, "lns_conv2d(latent): residual alignment");
  if (d->bias) LNS_REQUIRE((reinterpret_cast<uintptr_t>(d->bias) & 15) == 0, "lns_conv2d(latent): bias alignment");
  LatentParams p;
  p.x = reinterpret_cast<const uint16_t*>(d->x);
  p.w = reinterpret_cast<const uint16_t*>(d->w);
  p.bias = d->bias;
  p.act = d->act;
  p.residual = d->residual; p.res_dtype = d->res_dtype; p.res_bstride = d->res_bstride;
  p.y = d->y; p.y_dtype = d->y_dtype;
  p.x_bstride = d->x_bstride; p.y_bstride = d->y_bstride;
  p.B = d->B; p.dil = d->dil;
  p.nsuper = (d->B + 3) / 4;
  p.HWd = 8 + 2 * d->dil;
  p.plane_bytes = p.HWd * 2 * p.HWd * 128;
  LNS_REQUIRE(p.plane_bytes % 1024 == 0, "lns_conv2d(latent): internal: halo plane not 1024-byte aligned");
  p.x_f16 = d->x_dtype == LNS_F16 ? 1 : 0;
  p.inv_hwd = 1.0f / (float)p.HWd;
  // filter ring depth: 3 stages by default.  LNS_LATENT_RING raises it up to what the shared memory left by the halos allows
  // (dil 1: 6): measured on B200 a 6-stage ring is SLOWER (43 -> 52 us per launch at 1184 samples) -- the kernel is not
  // bound by the filter stream's latency, and the extra 48 KB of shared memory shrink the L1 the epilogue's loads hit.
  const int fixed = 4 * p.plane_bytes + 4 * 4096 + 256 + 1024;
  int bst = (227 * 1024 - fixed) / (int)kBBytes;
  {
    static int cap = 0;
    if (!cap) {
      const char* c = getenv("LNS_LATENT_RING");
      cap = c ? atoi(c) : 3;
      if (cap < 3) cap = 3;
      if (cap > kBMax) cap = kBMax;
    }
    if (bst > cap) bst = cap;
  }
  LNS_REQUIRE(bst >= 3, "lns_conv2d(latent): shared memory too small for a 3-stage filter ring at dilation %d", d->dil);
  p.bstages = bst;
  const int smem = fixed + bst * (int)kBBytes;
  LNS_OPT_IN_SMEM(conv_latent_kernel, 227 * 1024, "conv_latent");
  const int sms = device_sm_count();
  const int grid = p.nsuper < sms ? p.nsuper : sms;
  conv_latent_kernel<<<grid, kLatThreads, smem, stream>>>(p);
  return check_launch("conv_latent_kernel");
}

}  // namespace lns
