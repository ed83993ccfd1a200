// Propagator FFN in one kernel:  out = x + W2 . GELU( W1 . GN(x) )   (train_stage2_ns2d.py:44-53: GroupNorm(1,C) -> 1x1
// conv -> GELU -> 1x1 conv, both bias-free, residual add; C = hidden = 128).
//
// Unfused this is three launches per block (GroupNorm apply, conv+GELU, conv+residual) that move the 128-channel latent
// activation through L2/HBM six times.  Here a CTA takes 128 pixel rows: it reads x once, applies the per-(sample, channel)
// GroupNorm affine (scale/shift from lns_group_norm_affine) while packing the 16-bit A operand into shared memory
// (K-major SWIZZLE_128B, two 64-channel slabs), runs GEMM 1 on tcgen05 (accumulator in TMEM columns 0-127), applies GELU in
// the TMEM -> register epilogue, writes the hidden activation back over the A operand, runs GEMM 2 (columns 128-255), adds
// the residual and stores through a swizzled staging tile as full 128-byte lines.  Both filters (2 x 32 KB, pre-packed
// LNS_W_UMMA_*) arrive by cp.async.bulk.  97 KB of shared memory: two CTAs per SM overlap each other's phases.
#include "common.cuh"

namespace lns {

namespace qptx {
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
      "LNSQ_WAIT_%=:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra LNSQ_DONE_%=;\n\t"
      "bra LNSQ_WAIT_%=;\n\t"
      "LNSQ_DONE_%=:\n\t"
      "}" ::"r"(bar),
      "r"(parity)
      : "memory");
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
__device__ __forceinline__ void st_shared_v4(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
}  // namespace qptx

namespace {
constexpr int kFfnThreads = 160;  // warps 0-3: operand packing + epilogues, warp 4: TMEM owner + MMA issuer
constexpr uint32_t kSlab = 128u * 128u;  // one 64-channel slab of a 128-row operand: 16 KB

__device__ __forceinline__ uint64_t ffn_desc(uint32_t smem_addr) {
  uint64_t d = (uint64_t)((smem_addr & 0x3FFFFu) >> 4);
  d |= (uint64_t)(1024u >> 4) << 32;
  d |= 1ull << 46;
  d |= 2ull << 61;
  return d;
}

struct FfnParams {
  const uint16_t* x;     // [B][HW][128] 16-bit, batch stride x_bstride (elements)
  const float* scale;    // [B][128] GroupNorm affine (folded gamma / beta / prescale)
  const float* shift;    // [B][128]
  const uint16_t* w1;    // packed LNS_W_UMMA_*: [slab 2][128][64] swizzled
  const uint16_t* w2;
  uint16_t* y;           // [B][HW][128], batch stride y_bstride
  int64_t x_bstride, y_bstride;
  int HW, M;             // pixels per sample, total rows B*HW
  float inv_hw;
};
}  // namespace

template <bool F16>
__global__ void __launch_bounds__(kFfnThreads, 2) ffn_fused_kernel(const FfnParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (qptx::smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* gen = smem_raw + (base - qptx::smem_u32(smem_raw));
  const uint32_t a_s = base;                      // [slab 2][128 rows][128 B]  A operand of GEMM 1, then of GEMM 2
  const uint32_t w1_s = a_s + 2 * kSlab;
  const uint32_t w2_s = w1_s + 2 * kSlab;
  const uint32_t stage = w1_s;                    // 4 warps x 4 KB output staging: W1 is dead once GEMM 1 has completed
  const uint32_t bar_base = w2_s + 2 * kSlab;
  const uint32_t w_bar = bar_base, a_full = bar_base + 8, acc1 = bar_base + 16, a2_full = bar_base + 24, acc2 = bar_base + 32;
  const uint32_t tmem_slot = bar_base + 40;
  volatile uint32_t* tmem_slot_gen = reinterpret_cast<volatile uint32_t*>(gen + (tmem_slot - base));
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int m0 = blockIdx.x * 128;

  if (tid == 0) {
    qptx::mbar_init(w_bar, 1);
    qptx::mbar_init(a_full, 128);
    qptx::mbar_init(acc1, 1);
    qptx::mbar_init(a2_full, 128);
    qptx::mbar_init(acc2, 1);
    qptx::fence_mbar_init();
  }
  if (warp == 4) {
    qptx::tmem_alloc(tmem_slot, 256);
    qptx::tmem_relinquish();
  }
  qptx::tc_fence_before();
  __syncthreads();
  qptx::tc_fence_after();
  const uint32_t tmem_acc = *tmem_slot_gen;
  const uint32_t idesc = (1u << 4) | (F16 ? 0u : ((1u << 7) | (1u << 10))) | ((128u >> 3) << 17) | ((128u >> 4) << 24);

  if (warp == 4) {
    if (lane == 0) {
      qptx::mbar_expect_tx(w_bar, 4 * kSlab);
      qptx::bulk_g2s(w1_s, p.w1, 2 * kSlab, w_bar);
      qptx::bulk_g2s(w2_s, p.w2, 2 * kSlab, w_bar);
      qptx::mbar_wait(w_bar, 0);
      qptx::mbar_wait(a_full, 0);
      qptx::tc_fence_after();
#pragma unroll
      for (int sl = 0; sl < 2; ++sl) {
        const uint64_t ad = ffn_desc(a_s + sl * kSlab), bd = ffn_desc(w1_s + sl * kSlab);
#pragma unroll
        for (int k = 0; k < 4; ++k) qptx::umma_f16(tmem_acc, ad + (uint64_t)(2 * k), bd + (uint64_t)(2 * k), idesc, (sl | k) != 0 ? 1u : 0u);
      }
      qptx::umma_commit(acc1);
      qptx::mbar_wait(a2_full, 0);
      qptx::tc_fence_after();
#pragma unroll
      for (int sl = 0; sl < 2; ++sl) {
        const uint64_t ad = ffn_desc(a_s + sl * kSlab), bd = ffn_desc(w2_s + sl * kSlab);
#pragma unroll
        for (int k = 0; k < 4; ++k)
          qptx::umma_f16(tmem_acc + 128u, ad + (uint64_t)(2 * k), bd + (uint64_t)(2 * k), idesc, (sl | k) != 0 ? 1u : 0u);
      }
      qptx::umma_commit(acc2);
    }
    __syncwarp();
  } else {
    // ---- pack the normalised A operand: item = (row, 16-byte chunk of its 256 B); 16 consecutive threads = one row ----
    {
      const int c16 = tid & 15;
      uint4 raw[16];
      int bsel[16];
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        const int row = (tid >> 4) + i * 8, m = m0 + row;
        const bool ok = m < p.M;
        const int b = ok ? __float2int_rd(((float)m + 0.5f) * p.inv_hw) : 0;
        bsel[i] = ok ? b : -1;
        const int64_t off = (int64_t)b * p.x_bstride + (int64_t)(ok ? m - b * p.HW : 0) * 128 + c16 * 8;
        raw[i] = __ldg(reinterpret_cast<const uint4*>(p.x + off));
      }
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        const int row = (tid >> 4) + i * 8;
        uint32_t o[4] = {0u, 0u, 0u, 0u};
        if (bsel[i] >= 0) {
          const float4 s0 = __ldg(reinterpret_cast<const float4*>(p.scale + (int64_t)bsel[i] * 128 + c16 * 8));
          const float4 s1 = __ldg(reinterpret_cast<const float4*>(p.scale + (int64_t)bsel[i] * 128 + c16 * 8 + 4));
          const float4 t0 = __ldg(reinterpret_cast<const float4*>(p.shift + (int64_t)bsel[i] * 128 + c16 * 8));
          const float4 t1 = __ldg(reinterpret_cast<const float4*>(p.shift + (int64_t)bsel[i] * 128 + c16 * 8 + 4));
          const float2 v0 = unpack2_h16<F16>(raw[i].x), v1 = unpack2_h16<F16>(raw[i].y), v2 = unpack2_h16<F16>(raw[i].z),
                       v3 = unpack2_h16<F16>(raw[i].w);
          o[0] = pack2_h16<F16>(fmaf(v0.x, s0.x, t0.x), fmaf(v0.y, s0.y, t0.y));
          o[1] = pack2_h16<F16>(fmaf(v1.x, s0.z, t0.z), fmaf(v1.y, s0.w, t0.w));
          o[2] = pack2_h16<F16>(fmaf(v2.x, s1.x, t1.x), fmaf(v2.y, s1.y, t1.y));
          o[3] = pack2_h16<F16>(fmaf(v3.x, s1.z, t1.z), fmaf(v3.y, s1.w, t1.w));
        }
        qptx::st_shared_v4(a_s + (uint32_t)(c16 >> 3) * kSlab + (uint32_t)row * 128u + (uint32_t)(((c16 & 7) ^ (row & 7)) << 4), o[0],
                           o[1], o[2], o[3]);
      }
    }
    qptx::fence_proxy_async();
    qptx::mbar_arrive(a_full);

    const int quad = warp & 3;
    const int row = quad * 32 + lane;  // accumulator row of this thread = tile row
    const uint32_t t_lane = tmem_acc + ((uint32_t)(quad * 32) << 16);
    // ---- epilogue 1: hidden = GELU(acc1) -> 16-bit -> back over the A operand (GEMM 1 has finished reading it) ----
    qptx::mbar_wait(acc1, 0);
    qptx::tc_fence_after();
#pragma unroll
    for (int c0 = 0; c0 < 128; c0 += 32) {
      uint32_t raw[32];
      __syncwarp();
      qptx::tmem_ld32(t_lane + (uint32_t)c0, raw);
      qptx::tmem_ld_wait();
#pragma unroll
      for (int c8 = 0; c8 < 4; ++c8) {
        float v[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) v[j] = act_gelu_fast(__uint_as_float(raw[c8 * 8 + j]));
        const int ch = (c0 >> 3) + c8;  // 16-byte chunk 0..15 of the row
        qptx::st_shared_v4(a_s + (uint32_t)(ch >> 3) * kSlab + (uint32_t)row * 128u + (uint32_t)(((ch & 7) ^ (row & 7)) << 4),
                           pack2_h16<F16>(v[0], v[1]), pack2_h16<F16>(v[2], v[3]), pack2_h16<F16>(v[4], v[5]), pack2_h16<F16>(v[6], v[7]));
      }
    }
    qptx::fence_proxy_async();
    qptx::tc_fence_before();
    qptx::mbar_arrive(a2_full);
    // ---- epilogue 2: out = acc2 + x (residual, fp32 add) -> staged -> full-line stores ----
    const int m = m0 + row;
    const bool ok = m < p.M;
    const int b = ok ? __float2int_rd(((float)m + 0.5f) * p.inv_hw) : 0;
    const uint4* xres = reinterpret_cast<const uint4*>(p.x + (int64_t)b * p.x_bstride + (int64_t)(ok ? m - b * p.HW : 0) * 128);
    const uint32_t my_stage = stage + (uint32_t)quad * 4096u;
    const int rd_row = lane >> 3, rd_chunk = lane & 7;
    qptx::mbar_wait(acc2, 0);
    qptx::tc_fence_after();
    for (int cg = 0; cg < 128; cg += 64) {
#pragma unroll
      for (int cc = 0; cc < 64; cc += 32) {
        const int c0 = cg + cc;
        uint32_t raw[32];
        __syncwarp();
        qptx::tmem_ld32(t_lane + 128u + (uint32_t)c0, raw);
        uint4 rs[4];
#pragma unroll
        for (int c8 = 0; c8 < 4; ++c8) rs[c8] = __ldg(xres + (c0 >> 3) + c8);
        qptx::tmem_ld_wait();
#pragma unroll
        for (int c8 = 0; c8 < 4; ++c8) {
          const uint32_t rw[4] = {rs[c8].x, rs[c8].y, rs[c8].z, rs[c8].w};
          uint32_t o[4];
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const float2 r2 = unpack2_h16<F16>(rw[j]);
            o[j] = pack2_h16<F16>(__uint_as_float(raw[c8 * 8 + 2 * j]) + r2.x, __uint_as_float(raw[c8 * 8 + 2 * j + 1]) + r2.y);
          }
          const int ch = (cc >> 3) + c8;
          qptx::st_shared_v4(my_stage + (uint32_t)lane * 128u + (uint32_t)((ch ^ (lane & 7)) << 4), o[0], o[1], o[2], o[3]);
        }
      }
      __syncwarp();
#pragma unroll
      for (int pass = 0; pass < 8; ++pass) {
        const int r = pass * 4 + rd_row;
        const int mm = m0 + quad * 32 + r;
        uint32_t w0, w1, w2, w3;
        asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(w0), "=r"(w1), "=r"(w2), "=r"(w3)
                     : "r"(my_stage + (uint32_t)r * 128u + (uint32_t)((rd_chunk ^ (r & 7)) << 4)));
        if (mm < p.M) {
          const int bb = __float2int_rd(((float)mm + 0.5f) * p.inv_hw);
          *reinterpret_cast<uint4*>(p.y + (int64_t)bb * p.y_bstride + (int64_t)(mm - bb * p.HW) * 128 + cg + rd_chunk * 8) =
              make_uint4(w0, w1, w2, w3);
        }
      }
      __syncwarp();
    }
    qptx::tc_fence_before();
  }

  __syncthreads();
  if (warp == 4) {
    qptx::tc_fence_after();
    qptx::tmem_dealloc(tmem_acc, 256);
  }
}

}  // namespace lns

extern "C" {

int lns_ffn_fused_supported(int C, int hidden) { return C == 128 && hidden == 128; }

int lns_ffn_fused(const void* x, int dtype, int B, int HW, int C, int64_t x_bstride, const float* scale, const float* shift,
                  const void* w1_packed, const void* w2_packed, void* y, int64_t y_bstride, void* stream) {
  LNS_REQUIRE(x && scale && shift && w1_packed && w2_packed && y && B > 0 && HW > 0, "lns_ffn_fused: bad arguments");
  LNS_REQUIRE(lns::is_h16_host(dtype), "lns_ffn_fused: x/y must be LNS_BF16 or LNS_F16 (got %d)", dtype);
  LNS_REQUIRE(lns_ffn_fused_supported(C, 128), "lns_ffn_fused: C=%d (needs C = hidden = 128)", C);
  LNS_REQUIRE(x_bstride % 8 == 0 && y_bstride % 8 == 0, "lns_ffn_fused: batch strides must be multiples of 8");
  LNS_REQUIRE(((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(y) | reinterpret_cast<uintptr_t>(w1_packed) |
                reinterpret_cast<uintptr_t>(w2_packed) | reinterpret_cast<uintptr_t>(scale) | reinterpret_cast<uintptr_t>(shift)) & 15) == 0,
              "lns_ffn_fused: pointers must be 16-byte aligned");
  const int64_t M = (int64_t)B * HW;
  LNS_REQUIRE(M < (1 << 22), "lns_ffn_fused: too many rows (%lld) for the exact float index split", (long long)M);
  lns::FfnParams p;
  p.x = reinterpret_cast<const uint16_t*>(x);
  p.scale = scale; p.shift = shift;
  p.w1 = reinterpret_cast<const uint16_t*>(w1_packed);
  p.w2 = reinterpret_cast<const uint16_t*>(w2_packed);
  p.y = reinterpret_cast<uint16_t*>(y);
  p.x_bstride = x_bstride; p.y_bstride = y_bstride;
  p.HW = HW; p.M = (int)M;
  p.inv_hw = 1.0f / (float)HW;
  const int smem = 6 * (int)lns::kSlab + 64 + 1024;
  {
    LNS_OPT_IN_SMEM((lns::ffn_fused_kernel<false>), 227 * 1024, "ffn_fused");
    LNS_OPT_IN_SMEM((lns::ffn_fused_kernel<true>), 227 * 1024, "ffn_fused");
  }
  const int grid = (int)((M + 127) / 128);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (dtype == LNS_F16) lns::ffn_fused_kernel<true><<<grid, lns::kFfnThreads, smem, st>>>(p);
  else lns::ffn_fused_kernel<false><<<grid, lns::kFfnThreads, smem, st>>>(p);
  return lns::check_launch("ffn_fused_kernel");
}

}  // extern "C"
