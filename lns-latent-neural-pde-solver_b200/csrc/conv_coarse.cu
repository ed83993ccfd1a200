// tcgen05 implicit-GEMM 3x3 convolution for the COARSE levels of the autoencoders and the latent grids (LNS_ENGINE_COARSE):
// same-size stride-1 convs with Cin, Cout in {64, 128}, dilation 1-3, any padding mode per axis, optional nearest resize folded
// into the read -- modules/basics.py:249,252,295-299 (ResidualBlock / UpSampleBlock at 8x8 ... 24x48), train_stage2_SW.py:34-41
// and train_stage2_twophase.py:34-41 (the propagators' 3x3 convs on the 12x24 / 7x15 latent grids).
//
// It generalises conv_latent.cu (8x8 circular samples only) and exists for two reasons:
//  1. the gather engine (conv_umma.cu) re-reads every input pixel once per tap and the filter once per tile: L2 -> SM bound at
//     17-38 % of the tensor peak on these layers, and twice as bad on fp32-stored activations;
//  2. the split-operand precision mode ('fp16s') keeps the coarse levels in fp32 storage.  Here the activation is read ONCE
//     per 8x8 output block, converted ONCE into two IEEE-half planes (hi = rn(a), lo = rn(a - hi)) of the shared-memory halo, and
//     every tap is a descriptor view of those bytes: A.W = Ahi.W + Alo.W costs MMAs, not memory traffic.
//
// Tiling: the output grid of a sample is cut into 8x8 BLOCKS (ragged edges are computed and discarded).  A 128-row MMA tile
// holds two blocks whose (8+2d)x(8+2d) halos are interleaved row-wise -- pixel (slot, hy, hx) sits in row hy*(2*HWd) + slot*HWd +
// hx -- so that the sixteen 8-row groups (y, slot) of a tap view are equally spaced (HWd rows) and ONE SWIZZLE_128B descriptor with
// SBO = HWd*128 B covers both (accumulator row m = (y = m>>4, slot = (m>>3)&1, x = m&7)).  A CTA takes a super tile of 1 or 2 MMA
// tiles (what shared memory allows), keeps their halo planes resident for all nine taps and streams the filter through a ring
// in (tap, 64-channel slab) blocks that feed every plane of every tile.
// Warp roles (448 threads): warps 0-7 halo producers, warp 8 MMA issuer + TMEM owner, warps 9-12 epilogue, warp 13 filter ring.
#include "common.cuh"

namespace lns {

namespace cptx {
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
      "LNSC_WAIT_%=:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra LNSC_DONE_%=;\n\t"
      "bra LNSC_WAIT_%=;\n\t"
      "LNSC_DONE_%=:\n\t"
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
__device__ __forceinline__ int4 ldg_nc16(const int4* p) {
  int4 v;
  asm volatile("ld.global.nc.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
  return v;
}
__device__ __forceinline__ void st_shared_v4(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
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
// same with the descriptors given as (lo, hi) 32-bit halves
__device__ __forceinline__ void umma_f16_lohi(uint32_t d_tmem, uint32_t a_lo, uint32_t a_hi, uint32_t b_lo, uint32_t b_hi, uint32_t idesc,
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
}  // namespace cptx

namespace {
constexpr int kCoarseThreads = 448;  // warps 0-7 halo producers, 8 MMA issuer + TMEM owner, 9-12 epilogue, 13 filter ring
constexpr int kProducers = 256;
constexpr int kRingMax = 6;  // filter-ring barrier slots

struct CoarseParams {
  ConvGeom g;
  const void* x;       // NHWC 16-bit (AMODE 0) or fp32 (AMODE 1)
  const uint16_t* w;   // packed [tap 9][slab][Cout][64] swizzled; split filter: the lo image w_plane_elems further
  const float* bias;
  const float* pro_scale;  // fp32 input only: per-(sample, channel) affine + activation applied in the producer, before the
  const float* pro_shift;  // hi | lo split (the GroupNorm apply + Swish of the previous layer never exist as a tensor)
  int pro_act;
  int act;
  const void* residual;
  int res_dtype;
  int64_t res_bstride;
  void* y;
  int y_dtype;
  float* stats;        // [B][nb * 4][Cout][4] per-channel centred partial sums (sum d, sum d^2, pivot, count) of the stored output, or null
  int x_f16;           // operands are IEEE half (else bf16)
  int slabs;           // Cin / 64
  int nbx, nby, nb;    // 8x8 blocks per sample: columns, rows, total
  int nblocks;         // B * nb
  int tiles;           // MMA tiles per super tile (1 | 2)
  int nsuper;
  int HWd;             // halo width = height = 8 + 2*dil
  int plane_bytes;     // one halo plane: HWd * 2 * HWd rows x 128 B (a multiple of 1024)
  int aplanes;         // planes per (tile, slab): 1, or 2 = hi | lo of an fp32 activation
  int wsplit;          // split filter: every ring stage holds the hi block and the lo block
  int64_t w_plane_elems;
  int bstages;         // filter ring depth
  uint32_t bblock;     // bytes of one (tap, slab) filter block: Cout * 128
  int resize;          // 0 none, 1 exact 2x nearest, 2 general nearest
  float inv_hwd, inv_hv, inv_wv;
};

// source pixel offset (in pixels, within the sample) of halo pixel (hy, hx) of block (by, bx); -1 = zero padding
__device__ __forceinline__ int coarse_src(const CoarseParams& p, int by, int bx, int hy, int hx) {
  const ConvGeom& g = p.g;
  int yv = by * 8 - g.dil + hy, xv = bx * 8 - g.dil + hx;
  bool ok = true;
  if (g.circ_h) {
    yv += (yv < 0) ? g.Hv : 0;
    yv -= (yv >= g.Hv) ? g.Hv : 0;
  }
  ok = ok && ((unsigned)yv < (unsigned)g.Hv);  // (circular: only the discarded rows of a ragged block can still be outside)
  if (g.circ_w) {
    xv += (xv < 0) ? g.Wv : 0;
    xv -= (xv >= g.Wv) ? g.Wv : 0;
  }
  ok = ok && ((unsigned)xv < (unsigned)g.Wv);
  if (p.resize == 1) {
    yv >>= 1;
    xv >>= 1;
  } else if (p.resize == 2) {
    yv = __float2int_rd(((float)(yv * g.Hin) + 0.5f) * p.inv_hv);
    xv = __float2int_rd(((float)(xv * g.Win) + 0.5f) * p.inv_wv);
  }
  return ok ? yv * g.Win + xv : -1;
}
}  // namespace

// AMODE 0: 16-bit activations, one plane per (tile, slab), cp.async.  AMODE 1: fp32 activations, hi | lo planes through registers.
template <int AMODE>
__global__ void __launch_bounds__(kCoarseThreads, 1) conv_coarse_kernel(const CoarseParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (cptx::smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* gen = smem_raw + (base - cptx::smem_u32(smem_raw));
  const ConvGeom& g = p.g;
  const int T = p.tiles, SL = p.slabs, AP = p.aplanes;
  const uint32_t halo_a = base;  // planes [tile][slab][hi|lo]
  const int BS = p.bstages;
  const uint32_t bstage = p.bblock * (p.wsplit ? 2u : 1u);
  const uint32_t b_ring = halo_a + (uint32_t)(T * SL * AP) * (uint32_t)p.plane_bytes;
  const uint32_t bar_base = b_ring + (uint32_t)BS * bstage;
  // one (full, empty) barrier pair per 64-channel slab: the slabs of a super tile are filled and consumed one after the other,
  // so the fill of slab 1 (or of the next super tile's slab 0) overlaps the MMAs that read the other slab's planes
  auto halo_full = [&](int sl) { return bar_base + 8u * sl; };
  auto halo_empty = [&](int sl) { return bar_base + 16u + 8u * sl; };
  auto b_full = [&](int s) { return bar_base + 32u + 8u * s; };
  auto b_empty = [&](int s) { return bar_base + 32u + 8u * (kRingMax + s); };
  auto acc_full = [&](int a) { return bar_base + 32u + 8u * (2 * kRingMax + a); };
  auto acc_empty = [&](int a) { return bar_base + 32u + 8u * (2 * kRingMax + 2 + a); };
  const uint32_t tmem_slot = bar_base + 32u + 8u * (2 * kRingMax + 4);
  volatile uint32_t* tmem_slot_gen = reinterpret_cast<volatile uint32_t*>(gen + (tmem_slot - base));
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int d = g.dil, HWd = p.HWd;
  const int nkb = 9 * SL;

  if (tid == 0) {
    for (int sl = 0; sl < 2; ++sl) {
      cptx::mbar_init(halo_full(sl), kProducers);
      cptx::mbar_init(halo_empty(sl), 1);
    }
    for (int s = 0; s < kRingMax; ++s) {
      cptx::mbar_init(b_full(s), 1);
      cptx::mbar_init(b_empty(s), 1);
    }
    for (int a = 0; a < 2; ++a) {
      cptx::mbar_init(acc_full(a), 1);
      cptx::mbar_init(acc_empty(a), 128);
    }
    cptx::fence_mbar_init();
  }
  if (warp == 8) {
    cptx::tmem_alloc(tmem_slot, 512);
    cptx::tmem_relinquish();
  }
  cptx::tc_fence_before();
  __syncthreads();
  cptx::tc_fence_after();
  const uint32_t tmem_acc = *tmem_slot_gen;

  if (warp < 8) {
    // ============================== halo producers (256 threads) ==============================
    // all 256 threads work on ONE slab at a time: 8 chunk tasks per pixel (8 channels each: 16 B of 16-bit data, 32 B of fp32),
    // 32 pixels per pass
    const int chunk = tid & 7;
    constexpr int pstep = kProducers / 8;
    const int p0 = tid >> 3;
    const int npx = HWd * HWd;
    const int esz = AMODE == 1 ? 4 : 2;
    int it = 0;
    for (int sup = blockIdx.x; sup < p.nsuper; sup += gridDim.x, ++it) {
      for (int slab = 0; slab < SL; ++slab) {
        if (it > 0) cptx::mbar_wait(halo_empty(slab), (uint32_t)((it - 1) & 1));  // the previous super tile's MMAs have read this slab
        for (int ts = 0; ts < 2 * T; ++ts) {
          const int gb = sup * 2 * T + ts;
          if (gb >= p.nblocks) break;  // (rows of missing blocks are never stored)
          const int b = gb / p.nb;
          const int rem = gb - b * p.nb;
          const int by = rem / p.nbx, bx = rem - by * p.nbx;
          const uint8_t* xb = reinterpret_cast<const uint8_t*>(p.x) + ((int64_t)b * g.x_bstride + slab * 64 + chunk * 8) * esz;
          const uint32_t plane = halo_a + (uint32_t)((((ts >> 1) * SL + slab) * AP)) * (uint32_t)p.plane_bytes;
          float psc[8], psh[8];
          if (AMODE == 1 && p.pro_scale) {
            const float4* s4 = reinterpret_cast<const float4*>(p.pro_scale + (int64_t)b * g.Cin + slab * 64 + chunk * 8);
            const float4* t4 = reinterpret_cast<const float4*>(p.pro_shift + (int64_t)b * g.Cin + slab * 64 + chunk * 8);
            const float4 sa = __ldg(s4), sb = __ldg(s4 + 1), ta = __ldg(t4), tb = __ldg(t4 + 1);
            psc[0] = sa.x; psc[1] = sa.y; psc[2] = sa.z; psc[3] = sa.w; psc[4] = sb.x; psc[5] = sb.y; psc[6] = sb.z; psc[7] = sb.w;
            psh[0] = ta.x; psh[1] = ta.y; psh[2] = ta.z; psh[3] = ta.w; psh[4] = tb.x; psh[5] = tb.y; psh[6] = tb.z; psh[7] = tb.w;
          }
          if (AMODE == 0) {
            for (int q = p0; q < npx; q += pstep) {
              const int hy = __float2int_rd(((float)q + 0.5f) * p.inv_hwd), hx = q - hy * HWd;
              const int src = coarse_src(p, by, bx, hy, hx);
              const int r = hy * (2 * HWd) + (ts & 1) * HWd + hx;
              cptx::cp_async16(plane + (uint32_t)r * 128u + (uint32_t)((chunk ^ (r & 7)) << 4),
                               xb + (int64_t)(src < 0 ? 0 : src) * g.Cin * 2, src < 0 ? 0u : 16u);
            }
          } else {
            // fp32 -> (hi, lo) halves: kU pixels (2 kU 16-byte loads) in flight per thread (dilation 1: 100 halo pixels / 32 per
            // pass -> one batch of four)
            constexpr int kU = 4;
            for (int q0 = p0; q0 < npx; q0 += kU * pstep) {
              int4 raw[2 * kU];
              uint32_t dst[kU];
              bool live[kU], ok[kU];
#pragma unroll
              for (int u = 0; u < kU; ++u) {
                const int q = q0 + u * pstep;
                live[u] = q < npx;
                const int qq = live[u] ? q : 0;
                const int hy = __float2int_rd(((float)qq + 0.5f) * p.inv_hwd), hx = qq - hy * HWd;
                const int src = coarse_src(p, by, bx, hy, hx);
                ok[u] = src >= 0;
                const int r = hy * (2 * HWd) + (ts & 1) * HWd + hx;
                dst[u] = plane + (uint32_t)r * 128u + (uint32_t)((chunk ^ (r & 7)) << 4);
                const int4* s4 = reinterpret_cast<const int4*>(xb + (int64_t)(ok[u] ? src : 0) * g.Cin * 4);
                raw[2 * u] = cptx::ldg_nc16(s4);
                raw[2 * u + 1] = cptx::ldg_nc16(s4 + 1);
              }
#pragma unroll
              for (int u = 0; u < kU; ++u) {
                float v[8] = {__int_as_float(raw[2 * u].x), __int_as_float(raw[2 * u].y), __int_as_float(raw[2 * u].z),
                              __int_as_float(raw[2 * u].w), __int_as_float(raw[2 * u + 1].x), __int_as_float(raw[2 * u + 1].y),
                              __int_as_float(raw[2 * u + 1].z), __int_as_float(raw[2 * u + 1].w)};
                if (p.pro_scale) {
#pragma unroll
                  for (int j = 0; j < 8; ++j) v[j] = fmaf(v[j], psc[j], psh[j]);
                }
                if (p.pro_act != LNS_ACT_NONE) {
#pragma unroll
                  for (int j = 0; j < 8; ++j) v[j] = apply_act_fast(v[j], p.pro_act);
                }
                uint32_t hi[4], lo[4];
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                  if (p.x_f16) {
                    hi[j] = ok[u] ? pack2_h16<true>(v[2 * j], v[2 * j + 1]) : 0u;
                    const float2 back = unpack2_h16<true>(hi[j]);
                    lo[j] = ok[u] ? pack2_h16<true>(v[2 * j] - back.x, v[2 * j + 1] - back.y) : 0u;
                  } else {
                    hi[j] = ok[u] ? pack2_h16<false>(v[2 * j], v[2 * j + 1]) : 0u;
                    const float2 back = unpack2_h16<false>(hi[j]);
                    lo[j] = ok[u] ? pack2_h16<false>(v[2 * j] - back.x, v[2 * j + 1] - back.y) : 0u;
                  }
                }
                if (live[u]) {
                  cptx::st_shared_v4(dst[u], hi[0], hi[1], hi[2], hi[3]);
                  cptx::st_shared_v4(dst[u] + (uint32_t)p.plane_bytes, lo[0], lo[1], lo[2], lo[3]);
                }
              }
            }
          }
        }
        if (AMODE == 0) {
          cptx::cp_async_arrive_noinc(halo_full(slab));
        } else {
          cptx::fence_proxy_async();  // generic-proxy stores -> visible to tcgen05.mma's async-proxy reads
          cptx::mbar_arrive(halo_full(slab));
        }
      }
    }
  } else if (warp == 8) {
    // ============================== MMA issuer ==============================
    // One thread issues every tcgen05.mma of the CTA: its instruction stream must stay below the tensor pipe's 32 (N = 64) / 64
    // (N = 128) cycles per MMA.  Descriptors are (lo, hi) 32-bit pairs -- hi (SBO | version | swizzle) is loop invariant, lo is
    // the 16-byte-unit start address = plane base + tap view + 2k -- so an MMA costs two integer adds, not a 64-bit rebuild.
    if (lane == 0) {
      const uint32_t idesc = (1u << 4) | (p.x_f16 ? 0u : ((1u << 7) | (1u << 10))) | (((uint32_t)g.Cout >> 3) << 17) | ((128u >> 4) << 24);
      const uint32_t hi_a = (((uint32_t)HWd * 128u) >> 4) | (1u << 14) | (2u << 29);
      const uint32_t hi_b = (1024u >> 4) | (1u << 14) | (2u << 29);
      const uint32_t plane16 = (uint32_t)p.plane_bytes >> 4;
      const uint32_t halo_lo = (halo_a & 0x3FFFFu) >> 4, ring_lo = (b_ring & 0x3FFFFu) >> 4;
      const uint32_t bstage16 = bstage >> 4, blo16 = p.bblock >> 4;
      int s = 0, it = 0;
      uint32_t bphase = 0;  // filter ring position: stage s, phase bit flips on every wrap
      for (int sup = blockIdx.x; sup < p.nsuper; sup += gridDim.x, ++it) {
        const int a = it & 1;
        if (it >= 2) cptx::mbar_wait(acc_empty(a), (uint32_t)(((it >> 1) - 1) & 1));
        for (int sl = 0; sl < SL; ++sl) {
          cptx::mbar_wait(halo_full(sl), (uint32_t)(it & 1));
          cptx::fence_proxy_async();
          cptx::tc_fence_after();
          int ky = 0, kx = 0;
          for (int tap = 0; tap < 9; ++tap) {
            cptx::mbar_wait(b_full(s), bphase);
            cptx::tc_fence_after();
            const uint32_t b_lo = ring_lo + (uint32_t)s * bstage16;
            const uint32_t view16 = (uint32_t)((ky * d) * (2 * HWd) + kx * d) * 8u;  // 128-byte rows in 16-byte units
            for (int t = 0; t < T; ++t) {
              const uint32_t d_tmem = tmem_acc + (uint32_t)((a * T + t) * 128);
              const uint32_t pl = halo_lo + (uint32_t)((t * SL + sl) * AP) * plane16 + view16;
              for (int hl = 0; hl < AP; ++hl) {
                const uint32_t a_lo = pl + (uint32_t)hl * plane16;
#pragma unroll
                for (int k = 0; k < 4; ++k)
                  cptx::umma_f16_lohi(d_tmem, a_lo + 2u * k, hi_a, b_lo + 2u * k, hi_b, idesc, (sl | tap | hl | k) != 0 ? 1u : 0u);
              }
              if (p.wsplit) {  // Ahi . Wlo
#pragma unroll
                for (int k = 0; k < 4; ++k) cptx::umma_f16_lohi(d_tmem, pl + 2u * k, hi_a, b_lo + blo16 + 2u * k, hi_b, idesc, 1u);
              }
            }
            cptx::umma_commit(b_empty(s));
            if (++s == BS) {
              s = 0;
              bphase ^= 1u;
            }
            if (++kx == 3) {
              kx = 0;
              ++ky;
            }
          }
          cptx::umma_commit(halo_empty(sl));  // this slab's planes may be refilled
        }
        cptx::umma_commit(acc_full(a));
      }
    }
    __syncwarp();
  } else if (warp == 13) {
    // ============================== filter ring producer (one thread) ==============================
    if (lane == 0) {
      int s = 0;
      uint32_t wrap = 0;  // completed trips around the ring
      for (int sup = blockIdx.x; sup < p.nsuper; sup += gridDim.x) {
        for (int kb = 0; kb < nkb; ++kb) {  // slab-major, like the MMA loop: block (tap, slab) sits at tap * SL + slab
          if (wrap > 0) cptx::mbar_wait(b_empty(s), (wrap - 1u) & 1u);
          cptx::mbar_expect_tx(b_full(s), bstage);
          const int sl = kb / 9, tap = kb - sl * 9;
          const uint16_t* src = p.w + (int64_t)(tap * SL + sl) * g.Cout * 64;
          cptx::bulk_g2s(b_ring + (uint32_t)s * bstage, src, p.bblock, b_full(s));
          if (p.wsplit) cptx::bulk_g2s(b_ring + (uint32_t)s * bstage + p.bblock, src + p.w_plane_elems, p.bblock, b_full(s));
          if (++s == BS) {
            s = 0;
            ++wrap;
          }
        }
      }
    }
    __syncwarp();
  } else {
    // ============================== epilogue (warps 9-12) ==============================
    // TMEM lane = accumulator row = one output pixel; a thread walks its pixel's channels in 32-column steps: bias, activation,
    // residual, store (a pixel's 32 channels are 64 / 128 contiguous bytes).
    const int quad = warp & 3;
    const int m = quad * 32 + lane;
    const int yl = m >> 4, slot = (m >> 3) & 1, xl = m & 7;
    const bool y16 = is_h16(p.y_dtype);
    int it = 0;
    for (int sup = blockIdx.x; sup < p.nsuper; sup += gridDim.x, ++it) {
      const int a = it & 1;
      cptx::mbar_wait(acc_full(a), (uint32_t)((it >> 1) & 1));
      cptx::tc_fence_after();
      for (int t = 0; t < T; ++t) {
        const int gb = (sup * T + t) * 2 + slot;
        bool row_ok = gb < p.nblocks;
        int b = 0, yo = 0, xo = 0;
        if (row_ok) {
          b = gb / p.nb;
          const int rem = gb - b * p.nb;
          const int by = rem / p.nbx;
          yo = by * 8 + yl;
          xo = (rem - by * p.nbx) * 8 + xl;
          row_ok = yo < g.Hout && xo < g.Wout;
        }
        const int64_t pix = (int64_t)yo * g.Wout + xo;
        const int64_t yrow = (int64_t)b * g.y_bstride + pix * g.Cout;
        const int64_t rrow = (int64_t)b * p.res_bstride + pix * g.Cout;
        const uint32_t t_lane = tmem_acc + (uint32_t)((a * T + t) * 128) + ((uint32_t)(quad * 32) << 16);
        const bool blk_ok = gb < p.nblocks;
        const int blk_in_sample = blk_ok ? gb - b * p.nb : 0;
        for (int c0 = 0; c0 < g.Cout; c0 += 32) {
          uint32_t raw[32];
          __syncwarp();
          cptx::tmem_ld32(t_lane + (uint32_t)c0, raw);
          cptx::tmem_ld_wait();
          if (!row_ok && !p.stats) continue;
          float v[32];
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(raw[j]);
          if (row_ok) {
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
            if (p.residual) {
#pragma unroll
              for (int j = 0; j < 32; j += 4) {
                const float4 rr = ld4_as_float(p.residual, p.res_dtype, rrow + c0 + j);
                v[j] += rr.x; v[j + 1] += rr.y; v[j + 2] += rr.z; v[j + 3] += rr.w;
              }
            }
            if (y16) {
              uint4* dst = reinterpret_cast<uint4*>(reinterpret_cast<uint16_t*>(p.y) + yrow + c0);
#pragma unroll
              for (int h4 = 0; h4 < 4; ++h4) {
                uint32_t pk[4];
#pragma unroll
                for (int j = 0; j < 4; ++j) pk[j] = pack2_rt(p.y_dtype, v[h4 * 8 + 2 * j], v[h4 * 8 + 2 * j + 1]);
                dst[h4] = make_uint4(pk[0], pk[1], pk[2], pk[3]);
                if (p.stats) {  // statistics of the values as stored
#pragma unroll
                  for (int j = 0; j < 4; ++j) {
                    const float2 f = unpack2_rt(p.y_dtype, pk[j]);
                    v[h4 * 8 + 2 * j] = f.x; v[h4 * 8 + 2 * j + 1] = f.y;
                  }
                }
              }
            } else {
              float4* dst = reinterpret_cast<float4*>(reinterpret_cast<float*>(p.y) + yrow + c0);
              if (p.y_dtype == LNS_TF32) {
#pragma unroll
                for (int j = 0; j < 32; ++j) v[j] = round_tf32(v[j]);
              }
#pragma unroll
              for (int j = 0; j < 8; ++j) dst[j] = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
            }
          }
          if (p.stats) {
            // per-channel statistics over the warp's 16 pixels of this block, CENTRED on a pivot (the channel's value at the
            // chunk's first pixel): (sum d, sum d^2, pivot, count) with d = v - pivot.  Squaring the raw values instead loses
            // E[x^2] / var * 2^-24 of the variance -- 7e-4 of rstd on groups whose mean is 50 sigma (measured in the decoder's 16x16
            // level), which showed as a 1.5x larger 20-step drift.  Reduction: a halving butterfly over lane bits 4, 2, 1, 0 (bit
            // 3 = the block slot is not reduced), 30 shuffles per quantity instead of 128; the lane ends up with channels
            // c0 + 16 b4 + 8 b2 + 4 b1 + 2 b0 + {0, 1}.
            float s[32], q[32], pv[32];
            const int src_lane = lane & 8;  // first row of this slot in the warp (always a valid pixel if any row of the chunk is)
            const float cnt = (float)__popc(__ballot_sync(0xffffffffu, row_ok) & (slot ? 0xff00ff00u : 0x00ff00ffu));
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              pv[j] = __shfl_sync(0xffffffffu, v[j], src_lane);
              s[j] = row_ok ? v[j] - pv[j] : 0.f;
              q[j] = s[j] * s[j];
            }
#pragma unroll
            for (int step = 0; step < 4; ++step) {
              const int half = 16 >> step;                     // values kept per lane after this step
              const int bit = step == 0 ? 16 : (8 >> step);   // lane bit exchanged: 16, 4, 2, 1
              const bool up = (lane & bit) != 0;
#pragma unroll
              for (int j = 0; j < half; ++j) {
                const float ks = up ? s[j + half] : s[j], ss = up ? s[j] : s[j + half];
                const float kq = up ? q[j + half] : q[j], sq = up ? q[j] : q[j + half];
                s[j] = ks + __shfl_xor_sync(0xffffffffu, ss, bit);
                q[j] = kq + __shfl_xor_sync(0xffffffffu, sq, bit);
                pv[j] = up ? pv[j + half] : pv[j];               // (the pivot is uniform over the slot's lanes: a select, no shuffle)
              }
            }
            if (blk_ok) {
              const int cb = c0 + ((lane >> 4) & 1) * 16 + ((lane >> 2) & 1) * 8 + ((lane >> 1) & 1) * 4 + (lane & 1) * 2;
              float4* dst = reinterpret_cast<float4*>(p.stats + ((((int64_t)b * p.nb + blk_in_sample) * 4 + quad) * g.Cout + cb) * 4);
              dst[0] = make_float4(s[0], q[0], pv[0], cnt);
              dst[1] = make_float4(s[1], q[1], pv[1], cnt);
            }
          }
        }
      }
      cptx::tc_fence_before();
      cptx::mbar_arrive(acc_empty(a));
    }
  }

  cptx::tc_fence_before();
  __syncthreads();
  if (warp == 8) {
    cptx::tc_fence_after();
    cptx::tmem_dealloc(tmem_acc, 512);
  }
}

bool conv_coarse_supported(const LnsConvDesc* d) {
  if (!(d->KH == 3 && d->KW == 3 && d->stride == 1 && (d->Cin == 64 || d->Cin == 128) && (d->Cout == 64 || d->Cout == 128) &&
        d->dil >= 1 && d->dil <= 3 && d->pad_t == d->dil && d->pad_l == d->dil && d->Hout == d->Hv && d->Wout == d->Wv &&
        d->dil <= d->Hv && d->dil <= d->Wv && d->x_layout == LNS_NHWC && d->y_layout == LNS_NHWC && d->sample_bias == nullptr && d->pre_add == nullptr))
    return false;
  // the gather prologue (per-sample affine + activation) exists in the fp32 producer only (16-bit inputs arrive by cp.async)
  if ((d->pro_scale != nullptr || d->pro_act != LNS_ACT_NONE) && d->x_dtype != LNS_F32) return false;
  if ((d->pro_scale == nullptr) != (d->pro_shift == nullptr)) return false;
  if (d->x_dtype == LNS_F16) return d->w_format == LNS_W_UMMA_F16 || d->w_format == LNS_W_UMMA_F16X2;
  if (d->x_dtype == LNS_BF16) return d->w_format == LNS_W_UMMA_BF16;
  if (d->x_dtype == LNS_F32) return d->w_format == LNS_W_UMMA_F16 || d->w_format == LNS_W_UMMA_F16X2 || d->w_format == LNS_W_UMMA_BF16;
  return false;
}

int conv2d_coarse(const LnsConvDesc* d, cudaStream_t stream) {
  LNS_REQUIRE(conv_coarse_supported(d),
              "lns_conv2d(coarse): needs a same-size 3x3 stride-1 conv, Cin/Cout in {64,128}, pad = dil <= 3, NHWC input (16-bit, or "
              "fp32 = split into hi+lo halves), UMMA-packed weights (plain or split), no sample bias / pre-add (prologue: fp32 input only)");
  const bool f32in = d->x_dtype == LNS_F32;
  const int esz = f32in ? 4 : 2;
  LNS_REQUIRE((d->x_bstride * esz) % 16 == 0 && d->y_bstride % 8 == 0, "lns_conv2d(coarse): batch strides must keep 16-byte alignment");
  LNS_REQUIRE((reinterpret_cast<uintptr_t>(d->x) & 15) == 0 && (reinterpret_cast<uintptr_t>(d->y) & 15) == 0 &&
                  (reinterpret_cast<uintptr_t>(d->w) & 15) == 0, "lns_conv2d(coarse): x, y, w must be 16-byte aligned");
  if (d->residual) LNS_REQUIRE(d->res_bstride % 8 == 0 && (reinterpret_cast<uintptr_t>(d->residual) & 15) == 0, "lns_conv2d(coarse): residual alignment");
  if (d->bias) LNS_REQUIRE((reinterpret_cast<uintptr_t>(d->bias) & 15) == 0, "lns_conv2d(coarse): bias alignment");
  LNS_REQUIRE((int64_t)d->Hv * d->Hin < (1 << 21) && (int64_t)d->Wv * d->Win < (1 << 21), "lns_conv2d(coarse): spatial size too large");
  CoarseParams p;
  p.g = make_geom(d);
  p.x = d->x;
  p.w = reinterpret_cast<const uint16_t*>(d->w);
  p.bias = d->bias;
  p.pro_scale = d->pro_scale; p.pro_shift = d->pro_shift; p.pro_act = d->pro_act;
  if (d->pro_scale) LNS_REQUIRE(((reinterpret_cast<uintptr_t>(d->pro_scale) | reinterpret_cast<uintptr_t>(d->pro_shift)) & 15) == 0,
                                "lns_conv2d(coarse): prologue scale / shift alignment");
  p.act = d->act;
  p.residual = d->residual; p.res_dtype = d->res_dtype; p.res_bstride = d->res_bstride;
  p.y = d->y; p.y_dtype = d->y_dtype;
  p.stats = d->stats;
  if (d->stats) LNS_REQUIRE((reinterpret_cast<uintptr_t>(d->stats) & 15) == 0, "lns_conv2d(coarse): stats alignment");
  p.x_f16 = d->w_format == LNS_W_UMMA_BF16 ? 0 : 1;
  p.slabs = d->Cin / 64;
  p.nbx = cdiv(d->Wout, 8); p.nby = cdiv(d->Hout, 8); p.nb = p.nbx * p.nby;
  const int64_t nblocks = (int64_t)d->B * p.nb;
  LNS_REQUIRE(nblocks < (1ll << 30), "lns_conv2d(coarse): too many 8x8 blocks (%lld)", (long long)nblocks);
  p.nblocks = (int)nblocks;
  p.HWd = 8 + 2 * d->dil;
  p.plane_bytes = p.HWd * 2 * p.HWd * 128;
  LNS_REQUIRE(p.plane_bytes % 1024 == 0, "lns_conv2d(coarse): internal: halo plane not 1024-byte aligned");
  p.aplanes = f32in ? 2 : 1;
  p.wsplit = d->w_format == LNS_W_UMMA_F16X2 ? 1 : 0;
  p.w_plane_elems = (int64_t)d->Cout * d->Cin * 9;
  p.bblock = (uint32_t)d->Cout * 128u;
  p.resize = (d->Hv == d->Hin && d->Wv == d->Win) ? 0 : ((d->Hv == 2 * d->Hin && d->Wv == 2 * d->Win) ? 1 : 2);
  p.inv_hwd = 1.0f / (float)p.HWd;
  p.inv_hv = 1.0f / (float)d->Hv;
  p.inv_wv = 1.0f / (float)d->Wv;
  // shared memory: halo planes of 1 or 2 MMA tiles + a filter ring of >= 3 stages
  const int bstage = (int)p.bblock * (p.wsplit ? 2 : 1);
  const int tile_bytes = p.slabs * p.aplanes * p.plane_bytes;
  const int misc = 512 + 1024;
  int tiles = 2;
  if (2 * tile_bytes + 3 * bstage + misc > 227 * 1024) tiles = 1;
  LNS_REQUIRE(tiles * tile_bytes + 3 * bstage + misc <= 227 * 1024, "lns_conv2d(coarse): shared memory too small (Cin %d, dilation %d, %s input)",
              d->Cin, d->dil, f32in ? "fp32" : "16-bit");
  p.tiles = tiles;
  p.nsuper = cdiv(p.nblocks, 2 * tiles);
  int bst = (227 * 1024 - misc - tiles * tile_bytes) / bstage;
  if (bst > 4) bst = 4;
  p.bstages = bst;
  const int smem = tiles * tile_bytes + bst * bstage + misc;
  const int sms = device_sm_count();
  const int grid = p.nsuper < sms ? p.nsuper : sms;
  if (f32in) {
    LNS_OPT_IN_SMEM(conv_coarse_kernel<1>, 227 * 1024, "conv_coarse");
    conv_coarse_kernel<1><<<grid, kCoarseThreads, smem, stream>>>(p);
  } else {
    LNS_OPT_IN_SMEM(conv_coarse_kernel<0>, 227 * 1024, "conv_coarse");
    conv_coarse_kernel<0><<<grid, kCoarseThreads, smem, stream>>>(p);
  }
  return check_launch("conv_coarse_kernel");
}

}  // namespace lns
