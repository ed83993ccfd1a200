// tcgen05 implicit-GEMM 3x3 convolution with a shared-memory input HALO and a RESIDENT filter (LNS_ENGINE_HALO).
//
// Why: the generic gather engine (conv_umma.cu) re-reads every input pixel once per filter tap (9x) and the filter
// once per tile; on the C=64 full-resolution layers that is ~220 KB of L2->SM traffic per 128-pixel tile against
// 1152 tensor-core cycles -- L2/LSU bound at ~17% of the tensor peak (round-1 measurement).  Here
//   * a persistent CTA keeps the whole 3x3 filter (9 x Cout x 64 bf16, 72/144 KB) in shared memory for its lifetime
//     (one cp.async.bulk per tap -> UBLKCP, completion on an mbarrier with expect_tx),
//   * each output tile (16 rows x 8 columns = 128 pixels = the MMA's M) loads its input halo
//     (16+2d) x (8+2d) pixels x 128 B ONCE (cp.async, zero-fill / circular wrap / nearest resize in the index map),
//   * the nine taps are nine *views* of that halo: tcgen05's K-major SWIZZLE_128B operand addressing is a pure function
//     of the absolute shared-memory address (bits [4:6] ^= bits [7:9]; verified on B200 with tools/umma_probe.cu: any
//     128-byte-multiple start address and any 128-byte-multiple stride-byte-offset work with base_offset = 0), so
//     tap (ky,kx) is the SAME bytes addressed with start += ((ky*d)*HW + kx*d)*128 and SBO = HW*128 (HW = halo width).
// L2->SM traffic per tile drops from ~220 KB to 23 KB and the LSU work by 10x; the kernel becomes MMA/epilogue bound.
//
// Warp roles (288 threads): warps 0-3 halo producers, warp 4 TMEM owner + MMA issuer (one thread), warps 5-8 epilogue.
// Pipelines: halo ring (HS stages; full = 128 async cp.async arrivals, empty = tcgen05.commit) and a double-buffered
// TMEM accumulator (full = tcgen05.commit, empty = 128 epilogue arrivals): tile i's epilogue overlaps tile i+1's MMAs
// and tile i+2's halo loads.
#include "common.cuh"

namespace lns {

namespace hptx {
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
      "LNSH_WAIT_%=:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra LNSH_DONE_%=;\n\t"
      "bra LNSH_WAIT_%=;\n\t"
      "LNSH_DONE_%=:\n\t"
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
// 1-D bulk copy global -> shared through the TMA unit, completion counted in bytes on an mbarrier
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src),
               "r"(bytes), "r"(bar)
               : "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
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
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t* r) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
}  // namespace hptx

// K-major SWIZZLE_128B descriptor with an explicit stride-byte-offset (distance between 8-row groups)
__device__ __forceinline__ uint64_t make_desc_sbo(uint32_t smem_addr, uint32_t sbo_bytes) {
  uint64_t d = (uint64_t)((smem_addr & 0x3FFFFu) >> 4);
  d |= (uint64_t)(sbo_bytes >> 4) << 32;
  d |= 1ull << 46;
  d |= 2ull << 61;
  return d;
}

struct HaloParams {
  ConvGeom g;
  const __nv_bfloat16* x;
  const __nv_bfloat16* w;
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
  int tiles_x, tiles_y, ntiles;
  int HH, HW;          // halo height / width in pixels: 16 + 2d, 8 + 2d
  int halo_bytes;      // per stage, multiple of 1024
  int stages;          // halo ring depth
  int resize;          // 0 none, 1 exact 2x nearest, 2 general nearest
  float inv_hw, inv_hv, inv_wv;
};

constexpr int kHaloThreads = 288;
constexpr int kTileH = 16, kTileW = 8;

template <int NT>
__global__ void __launch_bounds__(kHaloThreads, 1) conv_halo_kernel(const HaloParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (hptx::smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_gen = smem_raw + (smem_base - hptx::smem_u32(smem_raw));
  constexpr uint32_t kWBytes = 9u * NT * 128u;
  const uint32_t w_base = smem_base;
  const uint32_t halo_base = smem_base + kWBytes;
  const uint32_t bar_base = halo_base + (uint32_t)p.stages * (uint32_t)p.halo_bytes;
  // barriers: w, halo_full[4], halo_empty[4], acc_full[2], acc_empty[2]; then the TMEM slot
  const uint32_t w_bar = bar_base;
  auto halo_full = [&](int s) { return bar_base + 8u * (1 + s); };
  auto halo_empty = [&](int s) { return bar_base + 8u * (5 + s); };
  auto acc_full = [&](int a) { return bar_base + 8u * (9 + a); };
  auto acc_empty = [&](int a) { return bar_base + 8u * (11 + a); };
  const uint32_t tmem_slot = bar_base + 8u * 13;
  volatile uint32_t* tmem_slot_gen =
      reinterpret_cast<volatile uint32_t*>(smem_gen + kWBytes + (size_t)p.stages * p.halo_bytes + 8 * 13);

  const ConvGeom& g = p.g;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int HS = p.stages;

  if (tid == 0) {
    hptx::mbar_init(w_bar, 1);
    for (int s = 0; s < 4; ++s) {
      hptx::mbar_init(halo_full(s), 128);
      hptx::mbar_init(halo_empty(s), 1);
    }
    for (int a = 0; a < 2; ++a) {
      hptx::mbar_init(acc_full(a), 1);
      hptx::mbar_init(acc_empty(a), 128);
    }
    hptx::fence_mbar_init();
  }
  if (warp == 4) {
    hptx::tmem_alloc(tmem_slot, 2 * NT);
    hptx::tmem_relinquish();
  }
  hptx::tc_fence_before();
  __syncthreads();
  hptx::tc_fence_after();
  const uint32_t tmem_acc = *tmem_slot_gen;

  if (warp < 4) {
    // ============================== halo producers ==============================
    if (tid == 0) {
      // resident filter: one bulk copy per tap (the packed image is already the swizzled shared-memory layout)
      hptx::mbar_expect_tx(w_bar, kWBytes);
      for (int tap = 0; tap < 9; ++tap)
        hptx::bulk_g2s(w_base + (uint32_t)tap * NT * 128u, p.w + (int64_t)tap * g.Cout * 64, (uint32_t)NT * 128u, w_bar);
    }
    const int chunk = tid & 7;
    const int npx = p.HH * p.HW;
    int it = 0;
    for (int tile = blockIdx.x; tile < p.ntiles; tile += gridDim.x, ++it) {
      const int s = it % HS;
      if (it >= HS) hptx::mbar_wait(halo_empty(s), ((it / HS) & 1) ^ 1);
      int tx = tile % p.tiles_x;
      int t2 = tile / p.tiles_x;
      int ty = t2 % p.tiles_y;
      int b = t2 / p.tiles_y;
      const int yv0 = ty * kTileH - g.dil, xv0 = tx * kTileW - g.dil;  // virtual coords of halo pixel (0,0)
      const uint32_t st = halo_base + (uint32_t)s * (uint32_t)p.halo_bytes;
      const __nv_bfloat16* xb = p.x + (int64_t)b * g.x_bstride + chunk * 8;
      for (int q0 = 0; q0 < npx; q0 += 16) {
        const int q = q0 + (tid >> 3);
        if (q < npx) {
          int hy = __float2int_rd(((float)q + 0.5f) * p.inv_hw);
          int hx = q - hy * p.HW;
          int yv = yv0 + hy, xv = xv0 + hx;
          if (g.circ_h) {
            yv += (yv < 0) ? g.Hv : 0;
            yv -= (yv >= g.Hv) ? g.Hv : 0;
          }
          if (g.circ_w) {
            xv += (xv < 0) ? g.Wv : 0;
            xv -= (xv >= g.Wv) ? g.Wv : 0;
          }
          const bool ok = ((unsigned)yv < (unsigned)g.Hv) && ((unsigned)xv < (unsigned)g.Wv);
          if (p.resize == 1) {
            yv >>= 1;
            xv >>= 1;
          } else if (p.resize == 2) {
            yv = __float2int_rd(((float)(yv * g.Hin) + 0.5f) * p.inv_hv);
            xv = __float2int_rd(((float)(xv * g.Win) + 0.5f) * p.inv_wv);
          }
          const void* src = ok ? (const void*)(xb + ((int64_t)yv * g.Win + xv) * 64) : (const void*)p.x;
          hptx::cp_async16(st + (uint32_t)q * 128u + (uint32_t)((chunk ^ (q & 7)) << 4), src, ok ? 16u : 0u);
        }
      }
      hptx::cp_async_arrive_noinc(halo_full(s));
    }
  } else if (warp == 4) {
    // ============================== MMA issuer ==============================
    if (lane == 0) {
      const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | (((uint32_t)g.Cout >> 3) << 17) | ((128u >> 4) << 24);
      const uint32_t sbo = (uint32_t)p.HW * 128u;
      hptx::mbar_wait(w_bar, 0);
      int it = 0;
      for (int tile = blockIdx.x; tile < p.ntiles; tile += gridDim.x, ++it) {
        const int s = it % HS, a = it & 1;
        if (it >= 2) hptx::mbar_wait(acc_empty(a), ((it >> 1) & 1) ^ 1);
        hptx::mbar_wait(halo_full(s), (it / HS) & 1);
        hptx::fence_proxy_async();
        hptx::tc_fence_after();
        const uint32_t st = halo_base + (uint32_t)s * (uint32_t)p.halo_bytes;
        const uint32_t d_tmem = tmem_acc + (uint32_t)(a * NT);
#pragma unroll
        for (int tap = 0; tap < 9; ++tap) {
          const int ky = tap / 3, kx = tap - ky * 3;
          const uint32_t a_addr = st + (uint32_t)((ky * g.dil) * p.HW + kx * g.dil) * 128u;
          const uint32_t b_addr = w_base + (uint32_t)tap * NT * 128u;
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            hptx::umma_bf16(d_tmem, make_desc_sbo(a_addr + k * 32, sbo), make_desc_sbo(b_addr + k * 32, 1024u), idesc,
                            (tap | k) != 0 ? 1u : 0u);
          }
        }
        hptx::umma_commit(halo_empty(s));
        hptx::umma_commit(acc_full(a));
      }
    }
    __syncwarp();
  } else {
    // ============================== epilogue (warps 5..8) ==============================
    const int quad = warp & 3;  // TMEM lane quadrant this warp may access
    const int m = quad * 32 + lane;
    const int ty_l = m >> 3, tx_l = m & 7;
    int it = 0;
    for (int tile = blockIdx.x; tile < p.ntiles; tile += gridDim.x, ++it) {
      const int a = it & 1;
      int tx = tile % p.tiles_x;
      int t2 = tile / p.tiles_x;
      int ty = t2 % p.tiles_y;
      int b = t2 / p.tiles_y;
      const int yo = ty * kTileH + ty_l, xo = tx * kTileW + tx_l;
      const bool row_ok = yo < g.Hout && xo < g.Wout;
      const int64_t pix = (int64_t)yo * g.Wout + xo;
      const int64_t yrow = (int64_t)b * g.y_bstride + pix * g.Cout;
      const int64_t prow = (int64_t)b * p.pre_add_bstride + pix * g.Cout;
      const int64_t rrow = (int64_t)b * p.res_bstride + pix * g.Cout;
      hptx::mbar_wait(acc_full(a), (it >> 1) & 1);
      hptx::tc_fence_after();
      const uint32_t t_lane = tmem_acc + (uint32_t)(a * NT) + ((uint32_t)(quad * 32) << 16);
      for (int c0 = 0; c0 < g.Cout; c0 += 16) {
        uint32_t raw[16];
        __syncwarp();
        hptx::tmem_ld16(t_lane + (uint32_t)c0, raw);
        hptx::tmem_ld_wait();
        if (row_ok) {
          float v[16];
#pragma unroll
          for (int j = 0; j < 16; ++j) v[j] = __uint_as_float(raw[j]);
          if (p.bias) {
#pragma unroll
            for (int j = 0; j < 16; j += 4) {
              float4 t = __ldg(reinterpret_cast<const float4*>(p.bias + c0 + j));
              v[j] += t.x; v[j + 1] += t.y; v[j + 2] += t.z; v[j + 3] += t.w;
            }
          }
          if (p.sample_bias) {
#pragma unroll
            for (int j = 0; j < 16; j += 4) {
              float4 t = __ldg(reinterpret_cast<const float4*>(p.sample_bias + (int64_t)b * g.Cout + c0 + j));
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
            for (int j = 0; j < 16; ++j) v[j] = apply_act(v[j], p.act);
          }
          if (p.residual) {
#pragma unroll
            for (int j = 0; j < 16; j += 4) {
              float4 t = ld4_as_float(p.residual, p.res_dtype, rrow + c0 + j);
              v[j] += t.x; v[j + 1] += t.y; v[j + 2] += t.z; v[j + 3] += t.w;
            }
          }
          if (p.y_dtype == LNS_BF16) {
            uint32_t pk[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              __nv_bfloat162 h2 = __floats2bfloat162_rn(v[2 * j], v[2 * j + 1]);
              pk[j] = *reinterpret_cast<uint32_t*>(&h2);
            }
            uint4* dst = reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(p.y) + yrow + c0);
            dst[0] = make_uint4(pk[0], pk[1], pk[2], pk[3]);
            dst[1] = make_uint4(pk[4], pk[5], pk[6], pk[7]);
          } else {
            float4* dst = reinterpret_cast<float4*>(reinterpret_cast<float*>(p.y) + yrow + c0);
#pragma unroll
            for (int j = 0; j < 4; ++j) dst[j] = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
          }
        }
      }
      hptx::tc_fence_before();
      hptx::mbar_arrive(acc_empty(a));
    }
  }

  hptx::tc_fence_before();
  __syncthreads();
  if (warp == 4) {
    hptx::tc_fence_after();
    hptx::tmem_dealloc(tmem_acc, 2 * NT);
  }
}

bool conv_halo_supported(const LnsConvDesc* d) {
  return d->KH == 3 && d->KW == 3 && d->stride == 1 && d->Cin == 64 && (d->Cout == 64 || d->Cout == 128) &&
         d->pad_t == d->dil && d->pad_l == d->dil && d->Hout == d->Hv && d->Wout == d->Wv && d->dil >= 1 &&
         d->dil <= 3 && d->x_dtype == LNS_BF16 && d->x_layout == LNS_NHWC && d->y_layout == LNS_NHWC &&
         d->pro_scale == nullptr && d->pro_act == LNS_ACT_NONE && d->w_format == LNS_W_UMMA_BF16 &&
         d->dil <= d->Hv && d->dil <= d->Wv;
}

template <int NT>
static int launch_halo(const HaloParams& p, int smem_bytes, int grid, cudaStream_t stream) {
  auto kern = conv_halo_kernel<NT>;
  static bool once = false;
  if (!once) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (e != cudaSuccess) {
      set_error("conv_halo: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
      return LNS_E_CUDA;
    }
    once = true;
  }
  kern<<<grid, kHaloThreads, smem_bytes, stream>>>(p);
  return check_launch("conv_halo_kernel");
}

int conv2d_halo(const LnsConvDesc* d, cudaStream_t stream) {
  LNS_REQUIRE(conv_halo_supported(d),
              "lns_conv2d(halo): needs a same-size 3x3 stride-1 conv, Cin=64, Cout in {64,128}, pad=dil<=3, NHWC bf16 "
              "input, UMMA-packed weights, no fused prologue");
  LNS_REQUIRE(d->x_bstride % 8 == 0 && d->y_bstride % 8 == 0, "lns_conv2d(halo): batch strides must be multiples of 8");
  LNS_REQUIRE((reinterpret_cast<uintptr_t>(d->x) & 15) == 0 && (reinterpret_cast<uintptr_t>(d->y) & 15) == 0 &&
                  (reinterpret_cast<uintptr_t>(d->w) & 15) == 0, "lns_conv2d(halo): x, y, w must be 16-byte aligned");
  if (d->pre_add) LNS_REQUIRE(d->pre_add_bstride % 8 == 0 && (reinterpret_cast<uintptr_t>(d->pre_add) & 15) == 0, "lns_conv2d(halo): pre_add alignment");
  if (d->residual) LNS_REQUIRE(d->res_bstride % 8 == 0 && (reinterpret_cast<uintptr_t>(d->residual) & 15) == 0, "lns_conv2d(halo): residual alignment");
  if (d->bias) LNS_REQUIRE((reinterpret_cast<uintptr_t>(d->bias) & 15) == 0, "lns_conv2d(halo): bias alignment");
  if (d->sample_bias) LNS_REQUIRE((reinterpret_cast<uintptr_t>(d->sample_bias) & 15) == 0, "lns_conv2d(halo): sample_bias alignment");
  LNS_REQUIRE((int64_t)d->Hv * d->Hin < (1 << 21) && (int64_t)d->Wv * d->Win < (1 << 21), "lns_conv2d(halo): spatial size too large");

  HaloParams p;
  p.g = make_geom(d);
  p.x = reinterpret_cast<const __nv_bfloat16*>(d->x);
  p.w = reinterpret_cast<const __nv_bfloat16*>(d->w);
  p.bias = d->bias; p.sample_bias = d->sample_bias;
  p.act = d->act;
  p.pre_add = d->pre_add; p.pre_add_dtype = d->pre_add_dtype; p.pre_add_bstride = d->pre_add_bstride;
  p.residual = d->residual; p.res_dtype = d->res_dtype; p.res_bstride = d->res_bstride;
  p.y = d->y; p.y_dtype = d->y_dtype;
  p.tiles_x = cdiv(d->Wout, kTileW);
  p.tiles_y = cdiv(d->Hout, kTileH);
  int64_t nt = (int64_t)p.tiles_x * p.tiles_y * d->B;
  LNS_REQUIRE(nt < (1ll << 31), "lns_conv2d(halo): too many tiles");
  p.ntiles = (int)nt;
  p.HH = kTileH + 2 * d->dil;
  p.HW = kTileW + 2 * d->dil;
  p.halo_bytes = ((p.HH * p.HW * 128) + 1023) & ~1023;
  p.resize = (d->Hv == d->Hin && d->Wv == d->Win) ? 0 : ((d->Hv == 2 * d->Hin && d->Wv == 2 * d->Win) ? 1 : 2);
  p.inv_hw = 1.0f / (float)p.HW;
  p.inv_hv = 1.0f / (float)d->Hv;
  p.inv_wv = 1.0f / (float)d->Wv;
  const int NT = d->Cout;
  const int fixed = 9 * NT * 128 + 256 + 1024;
  int stages = (227 * 1024 - fixed) / p.halo_bytes;
  if (stages > 4) stages = 4;
  LNS_REQUIRE(stages >= 2, "lns_conv2d(halo): shared memory too small for dilation %d with Cout %d", d->dil, d->Cout);
  p.stages = stages;
  const int smem = fixed + stages * p.halo_bytes;
  int sms = 148;
  {
    static int cached = 0;
    if (!cached) {
      int dev = 0;
      cudaGetDevice(&dev);
      cudaDeviceGetAttribute(&cached, cudaDevAttrMultiProcessorCount, dev);
      if (cached <= 0) cached = 148;
    }
    sms = cached;
  }
  int grid = p.ntiles < sms ? p.ntiles : sms;
  if (NT == 64) return launch_halo<64>(p, smem, grid, stream);
  return launch_halo<128>(p, smem, grid, stream);
}

}  // namespace lns
