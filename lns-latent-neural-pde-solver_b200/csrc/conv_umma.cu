// tcgen05 / TMEM implicit-GEMM convolution for sm_100a (LNS_ENGINE_UMMA).
//
// GEMM view:  D[m][n] = sum_{tap, c} A[src(m, tap)][c] * W[tap][c][n]
//   M = B*Hout*Wout output pixels (128 per CTA, tiles may straddle samples; rows are decoded individually),
//   N = Cout (tile NT = 64/128/256 accumulator columns in TMEM), K = KH*KW*Cin walked in 64-channel blocks.
// Operands are bf16, accumulation is fp32 in tensor memory.
//
// Shared-memory operand layout: the canonical K-major SWIZZLE_128B layout of tcgen05 -- one 128-byte row per
// M (or N) index holding 64 bf16 of K, the eight 16-byte chunks of a row XOR-ed with (row & 7); eight rows form a
// 1024-byte swizzle atom, atoms are stacked along M/N (stride-byte-offset 1024).
//   A tile (128 x 64): gathered by the four producer warps with 16-byte cp.async (zero-fill for padding taps).  The
//       gather implements every index map of the path in one place: circular / zero / half-periodic padding,
//       dilation, stride 2, nearest up-sampling folded into the read (never materialised), ragged M tails.
//   B tile (NT x 64): the filter was re-laid on the device once (lns_pack_conv_weight, LNS_W_UMMA_BF16) into exactly
//       this swizzled image, so a K block of the filter is a linear copy.
// Pipeline: STAGES-deep ring of (A,B) stages; mbarriers full[] (128 asynchronous producer arrivals,
// cp.async.mbarrier.arrive.noinc, + one fence.proxy.async on the consumer side) and empty[] (tcgen05.commit); one thread of warp 4 issues
// tcgen05.mma (4 x K=16 per stage); the accumulator is handed to the epilogue through a third mbarrier.
// Epilogue: warps 0-3 read their 32 TMEM lanes with tcgen05.ld (32x32b.x16), add bias / per-sample conditioning
// bias / pre-activation addend, apply GELU|SiLU, add the residual, and store 32-byte bf16 (or 64-byte fp32) row
// segments.  Two or three CTAs are resident per SM, so one CTA's epilogue overlaps another's main loop.
#include "common.cuh"

namespace lns {

namespace ptx {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "LNS_WAIT_%=:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra LNS_DONE_%=;\n\t"
      "bra LNS_WAIT_%=;\n\t"
      "LNS_DONE_%=:\n\t"
      "}" ::"r"(bar),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src, uint32_t src_bytes) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ int4 ldg_nc16(const int4* p) {
  int4 v;
  asm volatile("ld.global.nc.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
  return v;
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
// the mbarrier receives one arrival from this thread once all of its previously issued cp.async have landed
// (asynchronous: the thread does not wait); .noinc = the arrival counts against the barrier's initial expected count
__device__ __forceinline__ void cp_async_arrive_noinc(uint32_t bar) {
  asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(bar) : "memory");
}

__device__ __forceinline__ void tmem_alloc(uint32_t smem_dst, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_dst), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tmem_relinquish() {
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// D[tmem] (+)= A[smem desc] * B[smem desc], bf16 x bf16 -> fp32
__device__ __forceinline__ void umma_bf16(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}" ::"r"(d_tmem),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// same for fp32 operands read as TF32 (K = 8 per instruction = the same 32 bytes of a swizzled row)
__device__ __forceinline__ void umma_tf32(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t"
      "}" ::"r"(d_tmem),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// arrive on an mbarrier once all previously issued tcgen05.mma of this thread have completed
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
// 32 lanes x 16 consecutive fp32 columns -> 16 registers per thread
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t* r) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

}  // namespace ptx

// K-major SWIZZLE_128B shared-memory matrix descriptor: start>>4 | LBO(unused)=0 | SBO=1024>>4 | version=1 | layout=2
__device__ __forceinline__ uint64_t make_sw128_desc(uint32_t smem_addr) {
  uint64_t d = (uint64_t)((smem_addr & 0x3FFFFu) >> 4);
  d |= (uint64_t)(1024u >> 4) << 32;
  d |= 1ull << 46;
  d |= 2ull << 61;
  return d;
}
// instruction descriptor, kind::f16: D=f32, A=B=bf16, both K-major, M=128, N=n
__device__ __forceinline__ uint32_t make_idesc_bf16(uint32_t n) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((n >> 3) << 17) | ((128u >> 4) << 24);
}
// same with A=B=f16 (format code 0)
__device__ __forceinline__ uint32_t make_idesc_f16(uint32_t n) { return (1u << 4) | ((n >> 3) << 17) | ((128u >> 4) << 24); }
// kind::tf32: D=f32, A=B=tf32 (format code 2)
__device__ __forceinline__ uint32_t make_idesc_tf32(uint32_t n) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((n >> 3) << 17) | ((128u >> 4) << 24);
}

struct UmmaParams {
  ConvGeom g;
  const void* x;  // bf16 / f16 (ESZ 2) or fp32/tf32 (ESZ 4) NHWC
  const void* w;  // packed [tap][slab][Cout][128 B]
  const float* bias;
  const float* sample_bias;
  int act;
  const void* pre_add;
  int pre_add_dtype;
  int64_t pre_add_bstride;
  const void* residual;
  int res_dtype;
  int64_t res_bstride;
  void* y;
  int y_dtype;
  int x_f16;  // 16-bit operands are IEEE half instead of bf16
  int64_t w_plane_bytes;  // split filter (MODE 1 / 2): distance from the hi plane to the lo plane
  int M;
  int slabs;  // Cin / (128 / ESZ): 128-byte K blocks per pixel
  uint32_t x_bstride8;  // input batch stride in 16-byte units
  int resize;           // 0 none, 1 exact 2x nearest, 2 general nearest
  float inv_wout, inv_hout, inv_hv, inv_wv;
};

constexpr int kUmmaThreads = 160;  // warps 0-3: producers + epilogue; warp 4: TMEM owner + MMA issuer

// MODE 0: one A tile, one B tile per stage (plain operands).
// MODE 1 ("w2"): 16-bit activations, filter split into two IEEE-half planes (hi = rn(w), lo = rn(w - hi)); stage = [A][Bhi][Blo],
//         two MMAs per K step (A.Bhi + A.Blo): the filter rounding error disappears, the activation's stays.
// MODE 2 ("x3"): fp32 activations split IN the producer (hi = rn_f16(a), lo = rn_f16(a - hi)) + the split filter; stage =
//         [Ahi][Alo][Bhi][Blo], three MMAs per K step (Ahi.Bhi + Alo.Bhi + Ahi.Blo): fp32-class result (the dropped
//         Alo.Blo term is 2^-22 relative) on the f16 tensor path.
template <int NT, int STAGES, int MODE = 0>
struct UmmaSmem {
  static constexpr int kATile = 128 * 128;
  static constexpr int kBTile = NT * 128;
  static constexpr int kABytes = (MODE == 2 ? 2 : 1) * kATile;
  static constexpr int kBBytes = (MODE == 0 ? 1 : 2) * kBTile;
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kBarOffset = STAGES * kStageBytes;
  static constexpr int kTotal = kBarOffset + 256 /*barriers + tmem ptr*/ + 1024 /*alignment slack*/;
};

template <int NT, int STAGES, int ESZ, int MODE>
__global__ void __launch_bounds__(kUmmaThreads) conv_umma_kernel(const UmmaParams p) {
  using L = UmmaSmem<NT, STAGES, MODE>;
  static_assert(MODE == 0 || ESZ == 2, "split modes use 16-bit operands");
  constexpr int XSZ = MODE == 2 ? 4 : ESZ;  // bytes per element of the activation in global memory
  extern __shared__ uint8_t smem_raw[];
  // SWIZZLE_128B atoms must be 1024-byte aligned
  const uint32_t smem_base = (ptx::smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_gen = smem_raw + (smem_base - ptx::smem_u32(smem_raw));
  const uint32_t bar_base = smem_base + L::kBarOffset;
  // barriers: full[STAGES], empty[STAGES], accum
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (STAGES + s); };
  const uint32_t accum_bar = bar_base + 8u * (2 * STAGES);
  const uint32_t tmem_slot = bar_base + 8u * (2 * STAGES + 1);
  volatile uint32_t* tmem_slot_gen = reinterpret_cast<volatile uint32_t*>(smem_gen + L::kBarOffset + 8 * (2 * STAGES + 1));

  const ConvGeom& g = p.g;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int m0 = blockIdx.x * 128;
  const int n0 = blockIdx.y * NT;
  const int n_valid = min(NT, g.Cout - n0);  // multiple of 16
  const int taps = g.KH * g.KW;
  const int nkb = taps * p.slabs;
  const int HWo = g.Hout * g.Wout;

  if (tid == 0) {
    for (int s = 0; s < STAGES; ++s) {
      // MODE 2: 128 asynchronous arrivals for the filter copies + 128 plain arrivals after the converted A stores
      ptx::mbar_init(full_bar(s), MODE == 2 ? 256 : 128);
      ptx::mbar_init(empty_bar(s), 1);
    }
    ptx::mbar_init(accum_bar, 1);
    ptx::fence_mbar_init();
  }
  if (warp == 4) {
    ptx::tmem_alloc(tmem_slot, NT);
    ptx::tmem_relinquish();
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_acc = *tmem_slot_gen;

  if (warp < 4) {
    // ============================== producers ==============================
    // Thread t copies 16-byte chunk (t & 7) of rows (t >> 3) + 16*i, i = 0..7: the 8 lanes of a row read one full
    // 128-byte line (coalesced).  The source pixel of a (row, tap) pair is the same for those 8 lanes, so lane
    // (t & 7) computes it for row i = (t & 7) only and the group exchanges the 8 results with shuffles.  No integer
    // division in the loop: rows are decoded incrementally from the tile's first pixel, circular wrap is a
    // conditional add (host guarantees pad <= size), nearest resize is a shift (exact 2x) or an exact float divide.
    const int chunk = tid & 7;
    const int rbase = tid >> 3;
    const uint32_t sw_chunk = (uint32_t)((chunk ^ (rbase & 7)) << 4);
    const uint32_t cin8 = (uint32_t)((g.Cin * XSZ) >> 4);  // 16-byte chunks per pixel
    // my row for the address computation: rbase + 16 * chunk
    int my_yb, my_xb;
    uint32_t my_base;
    bool my_ok;
    {
      int b0 = m0 / HWo;
      int r0 = m0 - b0 * HWo;
      int y0 = r0 / g.Wout;
      int x0 = r0 - y0 * g.Wout;
      int rr = rbase + 16 * chunk;
      // x0 + rr < Wout + 128, y0 + q < Hout + 128: small numbers -> exact float quotient
      int t = x0 + rr;
      int q = __float2int_rd(((float)t + 0.5f) * p.inv_wout);
      int xo = t - q * g.Wout;
      int t2 = y0 + q;
      int q2 = __float2int_rd(((float)t2 + 0.5f) * p.inv_hout);
      int yo = t2 - q2 * g.Hout;
      int b = b0 + q2;
      my_ok = (m0 + rr) < p.M;
      my_yb = yo * g.stride - g.pad_t;
      my_xb = xo * g.stride - g.pad_l;
      my_base = (uint32_t)b * p.x_bstride8;
    }
    const int64_t w_kb_stride = (int64_t)g.Cout * 128;  // BYTES per (tap, slab) filter block
    const int b_chunks = n_valid * 8;                  // 16-byte chunks of the B tile
    const unsigned grp = 0xFFu << (lane & 24);         // the 8 lanes that share my rows
    uint32_t soff[8];                                  // source offset in 16-byte units, 0xFFFFFFFF = zero fill
    int issued = 0;
    for (int tap = 0; tap < taps; ++tap) {
      const int ky = tap / g.KW, kx = tap - ky * g.KW;
      uint32_t mine = 0xFFFFFFFFu;
      {
        int yv = my_yb + ky * g.dil, xv = my_xb + kx * g.dil;
        bool ok = my_ok;
        if (g.circ_h) {
          yv += (yv < 0) ? g.Hv : 0;
          yv -= (yv >= g.Hv) ? g.Hv : 0;
        } else {
          ok = ok && ((unsigned)yv < (unsigned)g.Hv);
        }
        if (g.circ_w) {
          xv += (xv < 0) ? g.Wv : 0;
          xv -= (xv >= g.Wv) ? g.Wv : 0;
        } else {
          ok = ok && ((unsigned)xv < (unsigned)g.Wv);
        }
        if (p.resize == 1) {          // exact 2x nearest
          yv >>= 1;
          xv >>= 1;
        } else if (p.resize == 2) {   // general nearest: floor(v * in / out), exact for these magnitudes
          yv = __float2int_rd(((float)(yv * g.Hin) + 0.5f) * p.inv_hv);
          xv = __float2int_rd(((float)(xv * g.Win) + 0.5f) * p.inv_wv);
        }
        if (ok) mine = my_base + (uint32_t)(yv * g.Win + xv) * cin8;
      }
#pragma unroll
      for (int i = 0; i < 8; ++i) soff[i] = __shfl_sync(0xFFFFFFFFu, mine, (lane & 24) + i);
      (void)grp;
      for (int slab = 0; slab < p.slabs; ++slab, ++issued) {
        const int s = issued % STAGES;
        if (issued >= STAGES) ptx::mbar_wait(empty_bar(s), ((issued / STAGES) & 1) ^ 1);
        const uint32_t a_dst = smem_base + s * L::kStageBytes;
        const uint32_t b_dst = a_dst + L::kABytes;
        const int4* wsrc = reinterpret_cast<const int4*>(reinterpret_cast<const uint8_t*>(p.w) + (int64_t)issued * w_kb_stride + (int64_t)n0 * 128);
        if (MODE == 2) {
          // filter planes first (asynchronous), then the activation through registers: 8 rows x 32 bytes of fp32 per thread,
          // all sixteen loads in flight before the first conversion
          for (int q = tid; q < b_chunks; q += 128) {
            ptx::cp_async16(b_dst + (uint32_t)q * 16u, wsrc + q, 16u);
            ptx::cp_async16(b_dst + L::kBTile + (uint32_t)q * 16u,
                            reinterpret_cast<const int4*>(reinterpret_cast<const uint8_t*>(wsrc) + p.w_plane_bytes) + q, 16u);
          }
          ptx::cp_async_arrive_noinc(full_bar(s));
          const int4* xsrc = reinterpret_cast<const int4*>(p.x) + slab * 16 + chunk * 2;
          int4 raw[16];
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const bool ok = soff[i] != 0xFFFFFFFFu;
            const int4* src = ok ? xsrc + soff[i] : reinterpret_cast<const int4*>(p.x);
            raw[2 * i] = ptx::ldg_nc16(src);  // volatile asm: keeps the sixteen loads ahead of the first conversion
            raw[2 * i + 1] = ptx::ldg_nc16(src + 1);
          }
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const bool ok = soff[i] != 0xFFFFFFFFu;
            const float v[8] = {__int_as_float(raw[2 * i].x), __int_as_float(raw[2 * i].y), __int_as_float(raw[2 * i].z),
                                __int_as_float(raw[2 * i].w), __int_as_float(raw[2 * i + 1].x), __int_as_float(raw[2 * i + 1].y),
                                __int_as_float(raw[2 * i + 1].z), __int_as_float(raw[2 * i + 1].w)};
            uint32_t hi[4], lo[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              hi[j] = ok ? pack2_h16<true>(v[2 * j], v[2 * j + 1]) : 0u;
              const float2 back = unpack2_h16<true>(hi[j]);
              lo[j] = ok ? pack2_h16<true>(v[2 * j] - back.x, v[2 * j + 1] - back.y) : 0u;
            }
            const uint32_t dst = a_dst + (uint32_t)((rbase + 16 * i) * 128) + sw_chunk;
            asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(dst), "r"(hi[0]), "r"(hi[1]), "r"(hi[2]), "r"(hi[3]) : "memory");
            asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(dst + L::kATile), "r"(lo[0]), "r"(lo[1]), "r"(lo[2]), "r"(lo[3]) : "memory");
          }
          ptx::fence_proxy_async();  // generic-proxy stores -> visible to tcgen05.mma's async-proxy reads
          ptx::mbar_arrive(full_bar(s));
          continue;
        }
        const int4* xsrc = reinterpret_cast<const int4*>(p.x) + slab * 8 + chunk;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const bool ok = soff[i] != 0xFFFFFFFFu;
          const void* src = ok ? (const void*)(xsrc + soff[i]) : (const void*)p.x;
          ptx::cp_async16(a_dst + (uint32_t)((rbase + 16 * i) * 128) + sw_chunk, src, ok ? 16u : 0u);
        }
        for (int q = tid; q < b_chunks; q += 128) ptx::cp_async16(b_dst + (uint32_t)q * 16u, wsrc + q, 16u);
        if (MODE == 1) {
          const int4* wlo = reinterpret_cast<const int4*>(reinterpret_cast<const uint8_t*>(wsrc) + p.w_plane_bytes);
          for (int q = tid; q < b_chunks; q += 128) ptx::cp_async16(b_dst + L::kBTile + (uint32_t)q * 16u, wlo + q, 16u);
        }
        // completion is signalled asynchronously: no thread ever blocks on its own copies, so up to STAGES K blocks
        // of loads are in flight per thread
        ptx::cp_async_arrive_noinc(full_bar(s));
      }
    }

    // ============================== epilogue ==============================
    ptx::mbar_wait(accum_bar, 0);
    ptx::tc_fence_after();
    const int m = m0 + warp * 32 + lane;
    const bool row_ok = m < p.M;
    int b = 0, pix = 0;
    if (row_ok) {
      b = m / HWo;
      pix = m - b * HWo;
    }
    const int64_t yrow = (int64_t)b * g.y_bstride + (int64_t)pix * g.Cout + n0;
    const int64_t prow = (int64_t)b * p.pre_add_bstride + (int64_t)pix * g.Cout + n0;
    const int64_t rrow = (int64_t)b * p.res_bstride + (int64_t)pix * g.Cout + n0;
    const uint32_t t_lane = tmem_acc + ((uint32_t)(warp * 32) << 16);
    for (int c0 = 0; c0 < n_valid; c0 += 16) {
      uint32_t raw[16];
      __syncwarp();  // tcgen05.ld is .sync.aligned: the whole warp issues it, only the stores are predicated
      ptx::tmem_ld16(t_lane + (uint32_t)c0, raw);
      ptx::tmem_ld_wait();
      if (row_ok) {
      float v[16];
#pragma unroll
      for (int j = 0; j < 16; ++j) v[j] = __uint_as_float(raw[j]);
      if (p.bias) {
#pragma unroll
        for (int j = 0; j < 16; j += 4) {
          float4 t = __ldg(reinterpret_cast<const float4*>(p.bias + n0 + c0 + j));
          v[j] += t.x; v[j + 1] += t.y; v[j + 2] += t.z; v[j + 3] += t.w;
        }
      }
      if (p.sample_bias) {
#pragma unroll
        for (int j = 0; j < 16; j += 4) {
          float4 t = __ldg(reinterpret_cast<const float4*>(p.sample_bias + (int64_t)b * g.Cout + n0 + c0 + j));
          v[j] += t.x; v[j + 1] += t.y; v[j + 2] += t.z; v[j + 3] += t.w;
        }
      }
      if (p.pre_add) {
#pragma unroll
        for (int j = 0; j < 16; j += 4) {
          float4 t = ld4_as_float(p.pre_add, p.pre_add_dtype, prow + c0 + j);
          v[j] += t.x; v[j + 1] += t.y; v[j + 2] += t.z; v[j + 3] += t.w;
        }
      }
      if (p.act != LNS_ACT_NONE) {
#pragma unroll
        for (int j = 0; j < 16; ++j) v[j] = ESZ == 2 ? apply_act_fast(v[j], p.act) : apply_act(v[j], p.act);
      }
      if (p.residual) {
#pragma unroll
        for (int j = 0; j < 16; j += 4) {
          float4 t = ld4_as_float(p.residual, p.res_dtype, rrow + c0 + j);
          v[j] += t.x; v[j + 1] += t.y; v[j + 2] += t.z; v[j + 3] += t.w;
        }
      }
      if (is_h16(p.y_dtype)) {
        uint32_t pk[8];
        if (p.y_dtype == LNS_F16) {
#pragma unroll
          for (int j = 0; j < 8; ++j) pk[j] = pack2_h16<true>(v[2 * j], v[2 * j + 1]);
        } else {
#pragma unroll
          for (int j = 0; j < 8; ++j) pk[j] = pack2_h16<false>(v[2 * j], v[2 * j + 1]);
        }
        uint4* dst = reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(p.y) + yrow + c0);
        dst[0] = make_uint4(pk[0], pk[1], pk[2], pk[3]);
        dst[1] = make_uint4(pk[4], pk[5], pk[6], pk[7]);
      } else {
        if (p.y_dtype == LNS_TF32) {
#pragma unroll
          for (int j = 0; j < 16; ++j) v[j] = round_tf32(v[j]);
        }
        float4* dst = reinterpret_cast<float4*>(reinterpret_cast<float*>(p.y) + yrow + c0);
#pragma unroll
        for (int j = 0; j < 4; ++j) dst[j] = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
      }
      }  // row_ok
    }
    ptx::tc_fence_before();
  } else {
    // ============================== MMA issuer (warp 4, one thread) ==============================
    if (lane == 0) {
      const uint32_t idesc = ESZ == 2 ? (p.x_f16 ? make_idesc_f16((uint32_t)n_valid) : make_idesc_bf16((uint32_t)n_valid))
                                      : make_idesc_tf32((uint32_t)n_valid);
      for (int kb = 0; kb < nkb; ++kb) {
        const int s = kb % STAGES;
        ptx::mbar_wait(full_bar(s), (kb / STAGES) & 1);
        ptx::fence_proxy_async();  // cp.async wrote the stage through the generic proxy; tcgen05.mma reads it via the async proxy
        ptx::tc_fence_after();
        const uint32_t a_addr = smem_base + s * L::kStageBytes;
        const uint64_t adesc = make_sw128_desc(a_addr);
        const uint64_t bdesc = make_sw128_desc(a_addr + L::kABytes);
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          // advance 16 bf16 = 32 bytes along K inside the swizzle atom: +2 in the (addr >> 4) field
          if (ESZ == 2)
            ptx::umma_bf16(tmem_acc, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), idesc, (kb | k) != 0 ? 1u : 0u);
          else
            ptx::umma_tf32(tmem_acc, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), idesc, (kb | k) != 0 ? 1u : 0u);
          if (MODE == 2)  // Alo . Bhi
            ptx::umma_bf16(tmem_acc, adesc + (uint64_t)(L::kATile >> 4) + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), idesc, 1u);
          if (MODE != 0)  // Ahi . Blo
            ptx::umma_bf16(tmem_acc, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(L::kBTile >> 4) + (uint64_t)(2 * k), idesc, 1u);
        }
        ptx::umma_commit(empty_bar(s));  // frees the stage when these MMAs have read it
      }
      ptx::umma_commit(accum_bar);  // accumulator complete
    }
    __syncwarp();
  }

  __syncthreads();
  if (warp == 4) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc(tmem_acc, NT);
  }
}

template <int NT, int STAGES, int ESZ, int MODE = 0>
static int launch_umma(const UmmaParams& p, int Cout, cudaStream_t stream) {
  using L = UmmaSmem<NT, STAGES, MODE>;
  auto kern = conv_umma_kernel<NT, STAGES, ESZ, MODE>;
  LNS_OPT_IN_SMEM(kern, L::kTotal, "conv_umma");  // per (template instance, device); never repeated under graph capture
  dim3 grid(cdiv(p.M, 128), cdiv(Cout, NT));
  kern<<<grid, kUmmaThreads, L::kTotal, stream>>>(p);
  return check_launch("conv_umma_kernel");
}

int conv2d_umma(const LnsConvDesc* d, cudaStream_t stream) {
  const bool tf32 = d->w_format == LNS_W_UMMA_TF32;
  const bool split = d->w_format == LNS_W_UMMA_F16X2;
  const bool x3 = split && d->x_dtype == LNS_F32;  // fp32 activations split in the producer: 3 MMAs per K step
  const int esz = tf32 ? 4 : 2, kblk = 128 / esz;  // channels per 128-byte K block
  const int xsz = (tf32 || x3) ? 4 : 2;            // bytes per activation element in global memory
  const bool f16 = d->w_format == LNS_W_UMMA_F16 || split;
  LNS_REQUIRE(d->w_format == LNS_W_UMMA_BF16 || f16 || tf32, "lns_conv2d(umma): weights must be packed as LNS_W_UMMA_BF16 / _F16 / _F16X2 / _TF32");
  LNS_REQUIRE(d->x_layout == LNS_NHWC && (tf32 ? !is_h16_host(d->x_dtype) : (x3 || d->x_dtype == (f16 ? LNS_F16 : LNS_BF16))),
              "lns_conv2d(umma): input must be NHWC bf16 / f16 (matching the filter format), NHWC fp32/tf32 (tf32 filter) or NHWC "
              "f16 / fp32 (split f16x2 filter)");
  LNS_REQUIRE(d->y_layout == LNS_NHWC, "lns_conv2d(umma): output must be NHWC");
  LNS_REQUIRE(d->Cin % kblk == 0 && d->Cout % 16 == 0, "lns_conv2d(umma): needs Cin%%%d==0 and Cout%%16==0 (got %d,%d)", kblk,
              d->Cin, d->Cout);
  LNS_REQUIRE(d->pro_scale == nullptr && d->pro_act == LNS_ACT_NONE,
              "lns_conv2d(umma): prologue affine is not fused in this engine; apply lns_affine_act first");
  LNS_REQUIRE(d->x_bstride % 8 == 0 && d->y_bstride % 8 == 0, "lns_conv2d(umma): batch strides must be multiples of 8");
  LNS_REQUIRE((reinterpret_cast<uintptr_t>(d->x) & 15) == 0 && (reinterpret_cast<uintptr_t>(d->y) & 15) == 0 &&
                  (reinterpret_cast<uintptr_t>(d->w) & 15) == 0,
              "lns_conv2d(umma): x, y, w must be 16-byte aligned");
  if (d->pre_add) LNS_REQUIRE(d->pre_add_bstride % 8 == 0 && (reinterpret_cast<uintptr_t>(d->pre_add) & 15) == 0, "lns_conv2d(umma): pre_add alignment");
  if (d->residual) LNS_REQUIRE(d->res_bstride % 8 == 0 && (reinterpret_cast<uintptr_t>(d->residual) & 15) == 0, "lns_conv2d(umma): residual alignment");
  if (d->bias) LNS_REQUIRE((reinterpret_cast<uintptr_t>(d->bias) & 15) == 0, "lns_conv2d(umma): bias must be 16-byte aligned");
  if (d->sample_bias) LNS_REQUIRE((reinterpret_cast<uintptr_t>(d->sample_bias) & 15) == 0, "lns_conv2d(umma): sample_bias alignment");
  int64_t M = (int64_t)d->B * d->Hout * d->Wout;
  LNS_REQUIRE(M < (1ll << 31), "lns_conv2d(umma): too many output pixels");
  int64_t x_elems = (int64_t)(d->B - 1) * d->x_bstride + (int64_t)d->Hin * d->Win * d->Cin;
  LNS_REQUIRE(((x_elems * xsz) >> 4) < 0xFFFFFFFFll, "lns_conv2d(umma): input too large for 32-bit chunk offsets");

  UmmaParams p;
  p.g = make_geom(d);
  p.x = d->x;
  p.w = d->w;
  p.bias = d->bias; p.sample_bias = d->sample_bias;
  p.act = d->act;
  p.pre_add = d->pre_add; p.pre_add_dtype = d->pre_add_dtype; p.pre_add_bstride = d->pre_add_bstride;
  p.residual = d->residual; p.res_dtype = d->res_dtype; p.res_bstride = d->res_bstride;
  p.y = d->y; p.y_dtype = d->y_dtype;
  p.x_f16 = f16 ? 1 : 0;
  p.M = (int)M;
  p.slabs = d->Cin / kblk;
  p.x_bstride8 = (uint32_t)((d->x_bstride * xsz) >> 4);
  p.w_plane_bytes = (int64_t)d->Cout * d->Cin * d->KH * d->KW * 2;
  p.resize = (d->Hv == d->Hin && d->Wv == d->Win) ? 0 : ((d->Hv == 2 * d->Hin && d->Wv == 2 * d->Win) ? 1 : 2);
  p.inv_wout = 1.0f / (float)d->Wout; p.inv_hout = 1.0f / (float)d->Hout;
  p.inv_hv = 1.0f / (float)d->Hv; p.inv_wv = 1.0f / (float)d->Wv;
  // the in-kernel wrap is a single conditional add/subtract and the resize uses exact float quotients of small ints
  LNS_REQUIRE(d->pad_t <= d->Hv && d->pad_l <= d->Wv &&
                  (int64_t)(d->Hout - 1) * d->stride + (int64_t)(d->KH - 1) * d->dil - d->pad_t < 2ll * d->Hv &&
                  (int64_t)(d->Wout - 1) * d->stride + (int64_t)(d->KW - 1) * d->dil - d->pad_l < 2ll * d->Wv,
              "lns_conv2d(umma): padding/dilation larger than the grid is not supported");
  LNS_REQUIRE((int64_t)d->Hv * d->Hin < (1 << 21) && (int64_t)d->Wv * d->Win < (1 << 21) && d->Hout < (1 << 20) &&
                  d->Wout < (1 << 20), "lns_conv2d(umma): spatial size too large");
  if (x3) {  // stage = 2 A tiles + 2 B tiles
    if (d->Cout <= 64) return launch_umma<64, 2, 2, 2>(p, d->Cout, stream);   // 96 KB: two CTAs per SM
    return launch_umma<128, 3, 2, 2>(p, d->Cout, stream);                     // 192 KB: one CTA per SM
  }
  if (split) {  // stage = A tile + 2 B tiles, twice the MMA work of a plain stage
    if (d->Cout <= 64) return launch_umma<64, 3, 2, 1>(p, d->Cout, stream);   // 96 KB
    return launch_umma<128, 2, 2, 1>(p, d->Cout, stream);                     // 96 KB
  }
  if (tf32) {
    if (d->Cout <= 64) return launch_umma<64, 4, 4>(p, d->Cout, stream);
    if (d->Cout <= 128) return launch_umma<128, 3, 4>(p, d->Cout, stream);
    return launch_umma<256, 3, 4>(p, d->Cout, stream);
  }
  if (d->Cout <= 64) return launch_umma<64, 4, 2>(p, d->Cout, stream);
  // short K (1x1 convs / linears): the kernel is epilogue bound and the 256-column tile's 144 KB of stages leave ONE CTA per
  // SM, i.e. no overlap of one tile's epilogue with another's loads; 128-column tiles run 2 CTAs per SM
  const bool short_k = (int64_t)d->KH * d->KW * d->Cin <= 256;
  if (d->Cout <= 128 || short_k) return launch_umma<128, 3, 2>(p, d->Cout, stream);
  return launch_umma<256, 3, 2>(p, d->Cout, stream);
}

}  // namespace lns
