// FABlock2D (modules/factorized_attention.py:144-159) in ONE kernel per sample: everything between the block's raw input
// and its output, except the tiny pooled branch that builds the axial kernels Kx / Ky.
//
// fablock_core_kernel (fablock.cu) keeps one head's u_phi slice in shared memory but still WRITES the normalised
// [H][W][heads*64] tensor (1 MB per 32x32 sample) for the two 1x1 convs of `to_out` to read back -- after round-1's other
// fusions that write + the two conv launches were 1/3 of the whole decode.  Here one CTA owns a SAMPLE, walks its heads in
// sequence and never lets the 8x channel-expanded tensor leave the SM:
//   per head h
//     A  u_phi_h = u x (W_in[h] * gn_scale) + W_in[h] . gn_shift          mma.sync, in place in shared memory
//     B  contraction over H with Kx[h] (per image column), C  contraction over W with Ky[h] (per image row)   mma.sync
//     D  InstanceNorm2d statistics of the head's 64 channels (from C's fp32 accumulators)
//     E  the normalisation is FOLDED into to_out[1]:  W1'[o][c] = W1[o][h*64+c] * rstd_c,  b1'[o] += -sum_c W1[o][h*64+c] mean_c rstd_c
//        and   acc[pixel][o] += u_phi_h[pixel][:] . W1'[o][:]   runs on tcgen05 with the fp32 accumulator of ALL pixels
//        (H*W x 64 fp32 = the whole 256 KB of tensor memory at 32x32) resident in TMEM across the heads.
//   after the last head:  GELU(acc + b1') -> 16-bit -> shared memory -> second tcgen05 GEMM with to_out[3] -> + skip -> out.
// HBM traffic per 32x32 sample: 128 KB in (re-read per head from L2) + 64 KB of axial kernels + 128 KB out, versus
// 128 KB + 1 MB + 1 MB + 128 KB + 128 KB + 128 KB for fablock_core + two conv launches.
//
// Shared-memory layout of u_phi_h: one 128-byte row (64 channels) per pixel, rows in tcgen05's K-major SWIZZLE_128B order
// (16-byte chunk index ^= row & 7) so that the SAME bytes are (1) ldmatrix operands of the mma.sync phases -- conflict free
// for 8 consecutive rows -- and (2) the A operand of the tcgen05 GEMM, no copy in between.  Pixel (y, x) lives in row
// p = y*Wp + (x ^ (y & 7)), Wp = W rounded up to 8: the XOR with the image row makes a COLUMN walk (phase B: fixed x,
// consecutive y) hit 8 different swizzle phases too; any pixel order is fine for the pointwise GEMMs.
#include "common.cuh"

namespace lns {

namespace fptx {
__device__ __forceinline__ uint32_t s32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "LNSF_WAIT_%=:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra LNSF_DONE_%=;\n\t"
      "bra LNSF_WAIT_%=;\n\t"
      "LNSF_DONE_%=:\n\t"
      "}" ::"r"(bar),
      "r"(parity)
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
__device__ __forceinline__ void ldsm_x4(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr));
}
__device__ __forceinline__ void ldsm_x2(uint32_t addr, uint32_t& r0, uint32_t& r1) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x2.shared.b16 {%0,%1}, [%2];" : "=r"(r0), "=r"(r1) : "r"(addr));
}
__device__ __forceinline__ void ldsm_x2_trans(uint32_t addr, uint32_t& r0, uint32_t& r1) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x2.trans.shared.b16 {%0,%1}, [%2];" : "=r"(r0), "=r"(r1) : "r"(addr));
}
__device__ __forceinline__ void ldsm_x4_trans(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr));
}
// four 8x8 16-bit matrices: lanes 8k..8k+7 give the row addresses of matrix k, register k of lane L holds row L/4, columns
// 2*(L%4), +1 of matrix k -- exactly a packed mma accumulator fragment
__device__ __forceinline__ void stsm_x4(uint32_t addr, uint32_t r0, uint32_t r1, uint32_t r2, uint32_t r3) {
  asm volatile("stmatrix.sync.aligned.m8n8.x4.shared.b16 [%0], {%1,%2,%3,%4};" ::"r"(addr), "r"(r0), "r"(r1), "r"(r2), "r"(r3) : "memory");
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
__device__ __forceinline__ void st_shared_v4(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
__device__ __forceinline__ void st_shared_b32(uint32_t addr, uint32_t a) {
  asm volatile("st.shared.b32 [%0], %1;" ::"r"(addr), "r"(a) : "memory");
}
}  // namespace fptx

namespace {
constexpr int kWS = 72;  // 16-bit elements per row of the (non-tcgen05) per-sample in_proj filter: 64 + 8 pad

// byte offset of 16-byte chunk `chunk` of pixel row `p` in the swizzled [rows][128 B] image
__device__ __forceinline__ uint32_t sw_off(int p, int chunk) { return (uint32_t)p * 128u + (uint32_t)((chunk ^ (p & 7)) << 4); }

// K-major SWIZZLE_128B descriptor (8-row groups 1024 B apart)
__device__ __forceinline__ uint64_t desc_sw128(uint32_t smem_addr) {
  uint64_t d = (uint64_t)((smem_addr & 0x3FFFFu) >> 4);
  d |= (uint64_t)(1024u >> 4) << 32;
  d |= 1ull << 46;
  d |= 2ull << 61;
  return d;
}

// In-place axial contraction of one line of u_phi_h:  out[i][c] = sum_j K[i][j] * line[j][c].
// AXIS 0: the line is image column `line` (elements j = image rows), AXIS 1: image row `line` (elements = columns).
template <int KT, int AXIS, bool F16>
__device__ __forceinline__ void contract_line_sw(uint32_t U_a, int line, int Wp, int n, uint32_t K_a, int kstride, int lane) {
  // element e of the line -> pixel row p (see the layout note at the top)
  auto slot = [&](int e) -> int { return AXIS == 0 ? e * Wp + (line ^ (e & 7)) : line * Wp + (e ^ (line & 7)); };
  uint32_t bf[KT][8][2];
#pragma unroll
  for (int kt = 0; kt < KT; ++kt) {
    int j = kt * 16 + (lane & 15);
    j = j < n ? j : n - 1;  // padded k rows: multiplied by the zero-padded kernel columns
    const int p = slot(j);
    const uint32_t rowa = U_a + (uint32_t)p * 128u;
    const int ph = p & 7, hi = lane >> 4;  // lanes 16-31 address the next channel chunk: two B fragments per ldmatrix.x4
#pragma unroll
    for (int nt = 0; nt < 8; nt += 2)
      fptx::ldsm_x4_trans(rowa + (uint32_t)(((nt + hi) ^ ph) << 4), bf[kt][nt][0], bf[kt][nt][1], bf[kt][nt + 1][0], bf[kt][nt + 1][1]);
  }
  __syncwarp();
  const int g = lane >> 2, t = lane & 3;
  const bool full = (n & 15) == 0;  // every 16-row tile complete: whole-fragment stmatrix stores (4 matrices per instruction)
#pragma unroll
  for (int mt = 0; mt < KT; ++mt) {
    const int i0 = mt * 16 + g, i1 = i0 + 8;
    const bool v0 = i0 < n, v1 = i1 < n;
    const int p0 = slot(v0 ? i0 : 0), p1 = slot(v1 ? i1 : 0);
    const uint32_t r0 = U_a + (uint32_t)p0 * 128u + (uint32_t)t * 4u, r1 = U_a + (uint32_t)p1 * 128u + (uint32_t)t * 4u;
    const int ph0 = p0 & 7, ph1 = p1 & 7;
    // stmatrix addressing: lane L supplies row (L & 7) of matrix L >> 3 = (channel chunk pair member L >> 4, row half (L >> 3) & 1)
    const int ps = slot(full ? mt * 16 + ((lane >> 3) & 1) * 8 + (lane & 7) : 0);
    const uint32_t rs = U_a + (uint32_t)ps * 128u;
    const int phs = ps & 7, cs = lane >> 4;
    // two halves of the 64 channels: 16 accumulator registers live instead of 32 (the B fragments of the whole line must
    // stay in registers until the last store; with 32 accumulators on top ptxas spilled them at the 128-register cap)
#pragma unroll
    for (int nh = 0; nh < 2; ++nh) {
      float acc[4][4];
#pragma unroll
      for (int nt = 0; nt < 4; ++nt) acc[nt][0] = acc[nt][1] = acc[nt][2] = acc[nt][3] = 0.f;
#pragma unroll
      for (int kt = 0; kt < KT; ++kt) {
        uint32_t a[4];
        fptx::ldsm_x4(K_a + (uint32_t)(((mt * 16 + (lane & 15)) * kstride + kt * 16 + (lane >> 4) * 8) * 2), a[0], a[1], a[2], a[3]);
#pragma unroll
        for (int nt = 0; nt < 4; ++nt) fptx::mma16816<F16>(acc[nt], a, bf[kt][nh * 4 + nt][0], bf[kt][nh * 4 + nt][1]);
      }
      if (full) {
#pragma unroll
        for (int nt = 0; nt < 4; nt += 2)
          fptx::stsm_x4(rs + (uint32_t)(((nh * 4 + nt + cs) ^ phs) << 4), pack2_h16<F16>(acc[nt][0], acc[nt][1]),
                        pack2_h16<F16>(acc[nt][2], acc[nt][3]), pack2_h16<F16>(acc[nt + 1][0], acc[nt + 1][1]),
                        pack2_h16<F16>(acc[nt + 1][2], acc[nt + 1][3]));
      } else {
#pragma unroll
        for (int nt = 0; nt < 4; ++nt) {
          const int c = nh * 4 + nt;
          if (v0) fptx::st_shared_b32(r0 + (uint32_t)((c ^ ph0) << 4), pack2_h16<F16>(acc[nt][0], acc[nt][1]));
          if (v1) fptx::st_shared_b32(r1 + (uint32_t)((c ^ ph1) << 4), pack2_h16<F16>(acc[nt][2], acc[nt][3]));
        }
      }
    }
  }
  __syncwarp();
}

// -DLNS_FULL_TRACE: per-phase cycle counts of one CTA (tools/bench_fablock.py with LNS_B200_LIB pointing at the trace build)
#ifdef LNS_FULL_TRACE
#define LNS_FT(i) do { if (tid == 0 && blockIdx.x == 0 && h < 4) p.trace[h * 16 + (i)] = clock64(); } while (0)
#else
#define LNS_FT(i) do { } while (0)
#endif

struct FullParams {
  const __nv_bfloat16* u;  // [B][H][W][64] 16-bit (opaque)
  int H, W, heads;
  const float* gn_scale;   // [B][64]
  const float* gn_shift;   // [B][64]
  const float* w_in;       // [heads*64][64]
  const float* Kx;         // [B][heads][H][H]
  const float* Ky;         // [B][heads][W][W]
  float eps;
  const float* w_out1;     // [64][heads*64]
  const float* w_out2;     // [64][64]
  __nv_bfloat16* out;      // [B][H][W][64]
  int Wp, T, tmem_cols;
  float inv_wp;
  int k16;      // Kx / Ky slices can be staged with 16-byte copies
  int lgwp;     // log2(Wp) when W == Wp is a power of two and the sample fills its tiles exactly (no pad rows), else -1
  long long* trace;  // -DLNS_FULL_TRACE builds only (tools): clock64() of thread 0 of one CTA at the phase boundaries
};

struct FullSmem {
  uint32_t U, Wo, W2, Ws, Kx, Ky, bias, obias, red, stat, gn, stage, bar, slot, total;
};
__host__ __device__ inline FullSmem full_layout(int H, int W, int T, int nwarp) {
  const int H16 = (H + 15) & ~15, W16 = (W + 15) & ~15;
  FullSmem L;
  uint32_t o = 0;
  L.U = o; o += (uint32_t)T * 16384u;
  L.Wo = o; o += 2u * 8192u;
  L.W2 = o; o += 8192u;
  L.Ws = o; o += 64u * kWS * 2u;
  L.Kx = o; o += (uint32_t)(H16 * (H16 + 8) * 2);
  L.Ky = o; o += (uint32_t)(W16 * (W16 + 8) * 2);
  o = (o + 15u) & ~15u;
  L.bias = o; o += 64u * 4u;
  L.obias = o; o += 64u * 4u;
  L.red = o; o += (uint32_t)nwarp * 64u * 2u * 4u;
  L.stat = o; o += 64u * 2u * 4u;
  L.gn = o; o += 2u * 64u * 4u;                                 // GroupNorm scale | shift of this sample
  L.stage = o; o += (64u * 64u + (uint32_t)(H * H + W * W)) * 4u;  // next head's fp32 in_proj slice | Kx | Ky (cp.async prefetch)
  o = (o + 15u) & ~15u;
  L.bar = o; o += 8u;
  L.slot = o; o += 8u;
  L.total = o;
  return L;
}
}  // namespace

// grid B (one CTA per sample), block NTHR (512 for >256 pixels, else 256)
template <int NTHR, bool F16>
__global__ void __launch_bounds__(NTHR, NTHR == 256 ? 2 : 1) fablock_full_kernel(const FullParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (fptx::s32(smem_raw) + 1023u) & ~1023u;
  uint8_t* gen = smem_raw + (base - fptx::s32(smem_raw));
  constexpr int nwarp = NTHR / 32;
  const int H = p.H, W = p.W, Wp = p.Wp, T = p.T, heads = p.heads;
  const FullSmem L = full_layout(H, W, T, nwarp);
  const int HW = H * W;
  const int H16 = (H + 15) & ~15, W16 = (W + 15) & ~15;
  const int kxs = H16 + 8, kys = W16 + 8;
  const uint32_t U_a = base + L.U, W2_a = base + L.W2, Ws_a = base + L.Ws, Kx_a = base + L.Kx, Ky_a = base + L.Ky;
  const uint32_t bar = base + L.bar;
  uint16_t* Ws_s = reinterpret_cast<uint16_t*>(gen + L.Ws);
  uint16_t* Kx_s = reinterpret_cast<uint16_t*>(gen + L.Kx);
  uint16_t* Ky_s = reinterpret_cast<uint16_t*>(gen + L.Ky);
  float* bias_s = reinterpret_cast<float*>(gen + L.bias);
  float* obias_s = reinterpret_cast<float*>(gen + L.obias);
  float* red_s = reinterpret_cast<float*>(gen + L.red);
  float* stat_s = reinterpret_cast<float*>(gen + L.stat);
  volatile uint32_t* slot_gen = reinterpret_cast<volatile uint32_t*>(gen + L.slot);
  float* gn_s = reinterpret_cast<float*>(gen + L.gn);
  const float* stw_s = reinterpret_cast<const float*>(gen + L.stage);  // [64][64] in_proj slice
  const float* stkx_s = stw_s + 64 * 64;                               // [H][H]
  const float* stky_s = stkx_s + H * H;                                // [W][W]
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int b = blockIdx.x;
  const int C = heads * 64;
  const __nv_bfloat16* ub = p.u + (int64_t)b * HW * 64;
  const int nslots = T * 128;
  // Per-head operands (fp32 in_proj slice 16 KB, Kx, Ky) are PREFETCHED one head ahead with cp.async into a staging area and
  // converted from shared memory: the first version loaded them with ld.global at the top of every head and sat in
  // long-scoreboard stalls for ~11k of its ~49k cycles per head (ncu, profiles/).  One commit group per call.
  auto prefetch_head = [&](int hh) {
    const uint32_t st_a = base + L.stage;
    const float* wsrc = p.w_in + (int64_t)hh * 64 * 64;
    for (int e = tid; e < 64 * 64 / 4; e += NTHR)
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(st_a + (uint32_t)e * 16u), "l"(wsrc + e * 4) : "memory");
    const float* kxsrc = p.Kx + ((int64_t)b * heads + hh) * H * H;
    const float* kysrc = p.Ky + ((int64_t)b * heads + hh) * W * W;
    if (p.k16) {  // H*H and W*W multiples of 4, 16-byte aligned bases: 16-byte copies (a quarter of the instructions)
      for (int e = tid; e < H * H / 4; e += NTHR)
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(st_a + (uint32_t)(64 * 64 + e * 4) * 4u), "l"(kxsrc + e * 4) : "memory");
      for (int e = tid; e < W * W / 4; e += NTHR)
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(st_a + (uint32_t)(64 * 64 + H * H + e * 4) * 4u), "l"(kysrc + e * 4) : "memory");
    } else {
      for (int e = tid; e < H * H; e += NTHR)
        asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(st_a + (uint32_t)(64 * 64 + e) * 4u), "l"(kxsrc + e) : "memory");
      for (int e = tid; e < W * W; e += NTHR)
        asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(st_a + (uint32_t)(64 * 64 + H * H + e) * 4u), "l"(kysrc + e) : "memory");
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };
  prefetch_head(0);
  if (tid < 128) gn_s[tid] = tid < 64 ? __ldg(p.gn_scale + (int64_t)b * 64 + tid) : __ldg(p.gn_shift + (int64_t)b * 64 + tid - 64);

  if (tid == 0) {
    fptx::mbar_init(bar, 1);
    fptx::fence_mbar_init();
  }
  if (warp == 0) {
    fptx::tmem_alloc(base + L.slot, (uint32_t)p.tmem_cols);
    fptx::tmem_relinquish();
  }
  if (tid < 64) obias_s[tid] = 0.f;
  // to_out[3] filter -> 16-bit swizzled K-major B operand [64 n][64 k]
  for (int e = tid; e < 64 * 8; e += NTHR) {
    const int n = e >> 3, kc = e & 7;
    const float4 w0 = __ldg(reinterpret_cast<const float4*>(p.w_out2 + n * 64 + kc * 8));
    const float4 w1 = __ldg(reinterpret_cast<const float4*>(p.w_out2 + n * 64 + kc * 8 + 4));
    fptx::st_shared_v4(W2_a + sw_off(n, kc), pack2_h16<F16>(w0.x, w0.y), pack2_h16<F16>(w0.z, w0.w), pack2_h16<F16>(w1.x, w1.y),
                       pack2_h16<F16>(w1.z, w1.w));
  }
  fptx::tc_fence_before();
  __syncthreads();
  fptx::tc_fence_after();
  const uint32_t tmem_acc = *slot_gen;
  // instruction descriptor, kind::f16: D = f32, A/B = bf16 (1) or f16 (0), K-major, N = 64, M = 128
  const uint32_t idesc = (1u << 4) | (F16 ? 0u : ((1u << 7) | (1u << 10))) | ((64u >> 3) << 17) | ((128u >> 4) << 24);

  const int g = lane >> 2, t = lane & 3;
  const int nblk = nslots >> 4;             // 16-row blocks of phase A (pad rows included: they cost nothing to skip)
  // the raw-input copy and phase A run in NQ pipelined parts of the row range.  Small samples (<= 256 pixels, the 256-thread
  // variant) use two: with four, a 16x16 sample has 4 blocks per part for 8 warps -- half of them idle through four barriers
  constexpr int NQ = NTHR == 256 ? 2 : 4;
  const int blk_per_q = (nblk + NQ - 1) / NQ;

  for (int h = 0; h < heads; ++h) {
    LNS_FT(0);
    // ---- per-(sample, head) setup from the staged fp32 operands: in_proj filter with GroupNorm folded in, its bias, the
    // two kernel matrices (no global-memory latency here; overlaps the previous head's tensor-core GEMM) ----
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    __syncthreads();  // staging of head h has landed for every thread (h == 0: also gn_s)
    for (int e = tid; e < 64 * 8; e += NTHR) {
      const int n = e >> 3, kc = e & 7;
      const float4 w0 = *reinterpret_cast<const float4*>(stw_s + n * 64 + kc * 8);
      const float4 w1 = *reinterpret_cast<const float4*>(stw_s + n * 64 + kc * 8 + 4);
      const float4 s0 = *reinterpret_cast<const float4*>(gn_s + kc * 8);
      const float4 s1 = *reinterpret_cast<const float4*>(gn_s + kc * 8 + 4);
      *reinterpret_cast<uint4*>(Ws_s + n * kWS + kc * 8) =
          make_uint4(pack2_h16<F16>(w0.x * s0.x, w0.y * s0.y), pack2_h16<F16>(w0.z * s0.z, w0.w * s0.w),
                     pack2_h16<F16>(w1.x * s1.x, w1.y * s1.y), pack2_h16<F16>(w1.z * s1.z, w1.w * s1.w));
    }
    {
      const int n = (tid >> 2) & 63, part = tid & 3;  // 4 threads per output channel (threads >= 256 recompute, do not store)
      float a = 0.f;
#pragma unroll
      for (int q4 = 0; q4 < 4; ++q4) {
        const float4 wv = *reinterpret_cast<const float4*>(stw_s + n * 64 + part * 16 + q4 * 4);
        const float4 sv = *reinterpret_cast<const float4*>(gn_s + 64 + part * 16 + q4 * 4);
        a = fmaf(wv.x, sv.x, a); a = fmaf(wv.y, sv.y, a); a = fmaf(wv.z, sv.z, a); a = fmaf(wv.w, sv.w, a);
      }
      a += __shfl_xor_sync(0xffffffffu, a, 1);
      a += __shfl_xor_sync(0xffffffffu, a, 2);
      if (part == 0 && tid < 256) bias_s[n] = a;
    }
    for (int i = warp; i < H16; i += nwarp)
      for (int j = lane; j < H16; j += 32) Kx_s[i * kxs + j] = to_h16<F16>((i < H && j < H) ? stkx_s[i * H + j] : 0.f);
    for (int i = warp; i < W16; i += nwarp)
      for (int j = lane; j < W16; j += 32) Ky_s[i * kys + j] = to_h16<F16>((i < W && j < W) ? stky_s[i * W + j] : 0.f);
    LNS_FT(1);
    // the previous head's tcgen05 GEMM reads U_s and Wo_s[(h-1)&1]: U_s may only be overwritten once it has completed
    if (h > 0) {
      fptx::mbar_wait(bar, (uint32_t)((h - 1) & 1));
      fptx::tc_fence_after();
    }
    __syncthreads();  // the staging area has been consumed by every thread: it may be refilled
    LNS_FT(2);
    // ---- raw input -> pixel rows of U_s (cp.async, four commit groups = four quarters of the row range) ----
    {
      const int ch = tid & 7;
      if (p.lgwp >= 0) {
        // no pad rows, power-of-two width: slot s holds pixel (y, x) = (s >> lg, (s & (Wp - 1)) ^ (y & 7)), i.e. source pixel index
        // s ^ ((s >> lg) & 7); the slot's swizzle phase s & 7 is loop invariant (the stride is a multiple of 8)
        const uint32_t dst0 = U_a + (uint32_t)((ch ^ ((tid >> 3) & 7)) << 4);
        const uint8_t* src0 = reinterpret_cast<const uint8_t*>(ub) + ch * 16;
        for (int q = 0; q < NQ; ++q) {
          const int s_end = min(nslots, (q + 1) * blk_per_q * 16);
          for (int s = q * blk_per_q * 16 + (tid >> 3); s < s_end; s += NTHR / 8)
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst0 + (uint32_t)s * 128u),
                         "l"(src0 + (size_t)(s ^ ((s >> p.lgwp) & 7)) * 128u) : "memory");
          asm volatile("cp.async.commit_group;" ::: "memory");
        }
      } else
      for (int q = 0; q < NQ; ++q) {
        const int s_end = min(nslots, (q + 1) * blk_per_q * 16);
        for (int s = q * blk_per_q * 16 + (tid >> 3); s < s_end; s += NTHR / 8) {
          const int y = __float2int_rd(((float)s + 0.5f) * p.inv_wp);
          const int x = (s - y * Wp) ^ (y & 7);
          const uint32_t dst = U_a + sw_off(s, ch);
          if (y < H && x < W)
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(ub + ((int64_t)y * W + x) * 64 + ch * 8) : "memory");
          else
            fptx::st_shared_v4(dst, 0u, 0u, 0u, 0u);
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
      }
    }
    prefetch_head(h + 1 < heads ? h + 1 : h);  // fifth group in flight; lands during phases A-C (last head: harmless refetch)

    LNS_FT(3);
    // ---- phase A: u_phi_h = u x Ws^T + bias, in place, 16-row blocks per warp (the conv is pointwise: any row order) ----
    uint32_t wf[4][8][2];
    for (int q = 0; q < NQ; ++q) {
      // groups in flight behind part q: the NQ - 1 - q later parts + the next head's prefetch
      if (NQ - q == 4) asm volatile("cp.async.wait_group 4;" ::: "memory");
      else if (NQ - q == 3) asm volatile("cp.async.wait_group 3;" ::: "memory");
      else if (NQ - q == 2) asm volatile("cp.async.wait_group 2;" ::: "memory");
      else asm volatile("cp.async.wait_group 1;" ::: "memory");
      __syncthreads();  // quarter q has landed for every thread (q == 0 also publishes the setup writes)
      if (q == 0) {
#pragma unroll
        for (int ks = 0; ks < 4; ++ks)
#pragma unroll
          for (int nt = 0; nt < 8; ++nt)
            fptx::ldsm_x2(Ws_a + (uint32_t)(((nt * 8 + (lane & 7)) * kWS + ks * 16 + ((lane >> 3) & 1) * 8) * 2), wf[ks][nt][0], wf[ks][nt][1]);
      }
      const int blk_end = min(nblk, (q + 1) * blk_per_q);
      for (int blk = q * blk_per_q + warp; blk < blk_end; blk += nwarp) {
        const int pr = blk * 16 + (lane & 15);
        uint32_t a[4][4];
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) fptx::ldsm_x4(U_a + sw_off(pr, ks * 2 + (lane >> 4)), a[ks][0], a[ks][1], a[ks][2], a[ks][3]);
        float acc[8][4];
#pragma unroll
        for (int nt = 0; nt < 8; ++nt) {  // the bias is the accumulator's initial value (columns 2t, 2t + 1 of both row halves)
          const float2 bb = *reinterpret_cast<const float2*>(bias_s + nt * 8 + t * 2);
          acc[nt][0] = acc[nt][2] = bb.x;
          acc[nt][1] = acc[nt][3] = bb.y;
        }
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) {
#pragma unroll
          for (int nt = 0; nt < 8; ++nt) fptx::mma16816<F16>(acc[nt], a[ks], wf[ks][nt][0], wf[ks][nt][1]);
        }
        __syncwarp();  // every lane's ldmatrix of the raw rows is done before they are overwritten
        // whole-fragment stores: lane L supplies row (L & 7) + 8 * ((L >> 3) & 1) of the block, channel chunk pair member L >> 4
        const int rs_row = blk * 16 + ((lane >> 3) & 1) * 8 + (lane & 7);
        const uint32_t rs = U_a + (uint32_t)rs_row * 128u;
        const int phs = rs_row & 7, cs = lane >> 4;
#pragma unroll
        for (int nt = 0; nt < 8; nt += 2) {
          fptx::stsm_x4(rs + (uint32_t)(((nt + cs) ^ phs) << 4), pack2_h16<F16>(acc[nt][0], acc[nt][1]),
                        pack2_h16<F16>(acc[nt][2], acc[nt][3]), pack2_h16<F16>(acc[nt + 1][0], acc[nt + 1][1]),
                        pack2_h16<F16>(acc[nt + 1][2], acc[nt + 1][3]));
        }
      }
    }
    __syncthreads();
    LNS_FT(4);

    // ---- phase B: contraction over H, one image column per warp at a time ----
    for (int m = warp; m < W; m += nwarp) {
      if (H16 == 16) contract_line_sw<1, 0, F16>(U_a, m, Wp, H, Kx_a, kxs, lane);
      else contract_line_sw<2, 0, F16>(U_a, m, Wp, H, Kx_a, kxs, lane);
    }
    __syncthreads();
    LNS_FT(5);
    // ---- phase C: contraction over W, one image row per warp ----
    for (int i = warp; i < H; i += nwarp) {
      if (W16 == 16) contract_line_sw<1, 1, F16>(U_a, i, Wp, W, Ky_a, kys, lane);
      else contract_line_sw<2, 1, F16>(U_a, i, Wp, W, Ky_a, kys, lane);
    }
    __syncthreads();
    LNS_FT(6);
    // ---- phase D: InstanceNorm statistics of this head's 64 channels, from the 16-bit values the GEMM will read (a separate
    // pass over shared memory: carrying 32 running sums through phase C cost 300 B of register spills per thread) ----
    {
      const int ch = tid & 7;
      float sm8[8], sq8[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) sm8[j] = sq8[j] = 0.f;
      auto accum = [&](uint32_t w0, uint32_t w1, uint32_t w2, uint32_t w3) {
        const uint32_t ww[4] = {w0, w1, w2, w3};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float2 f = unpack2_h16<F16>(ww[j]);
          sm8[2 * j] += f.x; sq8[2 * j] = fmaf(f.x, f.x, sq8[2 * j]);
          sm8[2 * j + 1] += f.y; sq8[2 * j + 1] = fmaf(f.y, f.y, sq8[2 * j + 1]);
        }
      };
      if (Wp == W && nslots == H * Wp && (nslots % (NTHR / 2)) == 0) {
        // no pad rows: four 16-byte loads in flight per thread
        for (int s = tid >> 3; s < nslots; s += NTHR / 2) {
          uint32_t w[4][4];
#pragma unroll
          for (int q4 = 0; q4 < 4; ++q4)
            asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(w[q4][0]), "=r"(w[q4][1]), "=r"(w[q4][2]), "=r"(w[q4][3])
                         : "r"(U_a + sw_off(s + q4 * (NTHR / 8), ch)));
#pragma unroll
          for (int q4 = 0; q4 < 4; ++q4) accum(w[q4][0], w[q4][1], w[q4][2], w[q4][3]);
        }
      } else {
        for (int s = tid >> 3; s < nslots; s += NTHR / 8) {
          const int y = __float2int_rd(((float)s + 0.5f) * p.inv_wp);
          if (y < H && ((s - y * Wp) ^ (y & 7)) < W) {
            uint32_t w0, w1, w2, w3;
            asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(w0), "=r"(w1), "=r"(w2), "=r"(w3) : "r"(U_a + sw_off(s, ch)));
            accum(w0, w1, w2, w3);
          }
        }
      }
#pragma unroll
      for (int j = 0; j < 8; ++j) {  // the 4 pixel lanes of a warp that share this channel chunk
        sm8[j] += __shfl_xor_sync(0xffffffffu, sm8[j], 8);
        sm8[j] += __shfl_xor_sync(0xffffffffu, sm8[j], 16);
        sq8[j] += __shfl_xor_sync(0xffffffffu, sq8[j], 8);
        sq8[j] += __shfl_xor_sync(0xffffffffu, sq8[j], 16);
      }
      if (lane < 8) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          red_s[(warp * 64 + ch * 8 + j) * 2 + 0] = sm8[j];
          red_s[(warp * 64 + ch * 8 + j) * 2 + 1] = sq8[j];
        }
      }
    }
    __syncthreads();
    LNS_FT(7);
    if (tid < 64) {
      double sm = 0.0, ss = 0.0;
      for (int w = 0; w < nwarp; ++w) {
        sm += (double)red_s[(w * 64 + tid) * 2 + 0];
        ss += (double)red_s[(w * 64 + tid) * 2 + 1];
      }
      const double inv_n = 1.0 / (double)HW;
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
    __syncthreads();
    LNS_FT(8);
    // ---- phase E: fold the normalisation into to_out[1]'s filter slice and issue this head's tensor-core GEMM ----
    const uint32_t Wo_a = base + L.Wo + (uint32_t)(h & 1) * 8192u;
    for (int e = tid; e < 64 * 8; e += NTHR) {
      const int n = e >> 3, kc = e & 7;
      const float4 w0 = __ldg(reinterpret_cast<const float4*>(p.w_out1 + (int64_t)n * C + h * 64 + kc * 8));
      const float4 w1 = __ldg(reinterpret_cast<const float4*>(p.w_out1 + (int64_t)n * C + h * 64 + kc * 8 + 4));
      const float wv[8] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w};
      float sc[8];
      float bsum = 0.f;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        sc[j] = wv[j] * stat_s[(kc * 8 + j) * 2 + 0];
        bsum = fmaf(wv[j], stat_s[(kc * 8 + j) * 2 + 1], bsum);
      }
      fptx::st_shared_v4(Wo_a + sw_off(n, kc), pack2_h16<F16>(sc[0], sc[1]), pack2_h16<F16>(sc[2], sc[3]), pack2_h16<F16>(sc[4], sc[5]),
                         pack2_h16<F16>(sc[6], sc[7]));
      // the 8 k-chunks of output channel n sit in 8 consecutive lanes: fixed-order butterfly, one lane adds it to the bias
      bsum += __shfl_xor_sync(0xffffffffu, bsum, 1);
      bsum += __shfl_xor_sync(0xffffffffu, bsum, 2);
      bsum += __shfl_xor_sync(0xffffffffu, bsum, 4);
      if (kc == 0) obias_s[n] += bsum;
    }
    fptx::fence_proxy_async();  // generic-proxy writes of U_s / Wo_s -> visible to the tensor core's async-proxy reads
    fptx::tc_fence_before();
    __syncthreads();
    LNS_FT(9);
    if (tid == 0) {
      fptx::tc_fence_after();
      const uint64_t bdesc = desc_sw128(Wo_a);
      for (int tl = 0; tl < T; ++tl) {
        const uint64_t adesc = desc_sw128(U_a + (uint32_t)tl * 16384u);
#pragma unroll
        for (int k = 0; k < 4; ++k)
          fptx::umma_f16(tmem_acc + (uint32_t)(tl * 64), adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), idesc, (h | k) != 0 ? 1u : 0u);
      }
      fptx::umma_commit(bar);
    }
    LNS_FT(10);
  }

  // ================= after the last head: GELU(acc + b1') -> to_out[3] -> + skip -> out =================
  asm volatile("cp.async.wait_group 0;" ::: "memory");  // (the last head's harmless staging refetch)
  fptx::mbar_wait(bar, (uint32_t)((heads - 1) & 1));
  fptx::tc_fence_after();
  const int quad = warp & 3, sub = warp >> 2, nsub = nwarp >> 2;
  for (int tl = sub; tl < T; tl += nsub) {
    const int pr = tl * 128 + quad * 32 + lane;
    const uint32_t row_a = U_a + (uint32_t)pr * 128u;
    const int ph = pr & 7;
#pragma unroll
    for (int cc = 0; cc < 64; cc += 32) {
      uint32_t raw[32];
      fptx::tmem_ld32(tmem_acc + (uint32_t)(tl * 64 + cc) + ((uint32_t)(quad * 32) << 16), raw);
      fptx::tmem_ld_wait();
#pragma unroll
      for (int c8 = 0; c8 < 4; ++c8) {
        float v[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) v[j] = act_gelu_fast(__uint_as_float(raw[c8 * 8 + j]) + obias_s[cc + c8 * 8 + j]);
        fptx::st_shared_v4(row_a + (uint32_t)((((cc >> 3) + c8) ^ ph) << 4), pack2_h16<F16>(v[0], v[1]), pack2_h16<F16>(v[2], v[3]),
                           pack2_h16<F16>(v[4], v[5]), pack2_h16<F16>(v[6], v[7]));
      }
    }
  }
  fptx::fence_proxy_async();
  fptx::tc_fence_before();
  __syncthreads();
  if (tid == 0) {
    fptx::tc_fence_after();
    const uint64_t bdesc = desc_sw128(W2_a);
    for (int tl = 0; tl < T; ++tl) {
      const uint64_t adesc = desc_sw128(U_a + (uint32_t)tl * 16384u);
#pragma unroll
      for (int k = 0; k < 4; ++k)
        fptx::umma_f16(tmem_acc + (uint32_t)(tl * 64), adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), idesc, k != 0 ? 1u : 0u);
    }
    fptx::umma_commit(bar);
  }
  fptx::mbar_wait(bar, (uint32_t)(heads & 1));
  fptx::tc_fence_after();
  // epilogue 2: + skip (the block's raw input, fp32 add), 16-bit result back into the pixel's own row of U_s
  for (int tl = sub; tl < T; tl += nsub) {
    const int pr = tl * 128 + quad * 32 + lane;
    const int y = __float2int_rd(((float)pr + 0.5f) * p.inv_wp);
    const int x = (pr - y * Wp) ^ (y & 7);
    const bool valid = y < H && x < W;
    const uint4* skip = reinterpret_cast<const uint4*>(ub + ((int64_t)(valid ? y : 0) * W + (valid ? x : 0)) * 64);
    const uint32_t row_a = U_a + (uint32_t)pr * 128u;
    const int ph = pr & 7;
#pragma unroll
    for (int cc = 0; cc < 64; cc += 32) {
      uint32_t raw[32];
      fptx::tmem_ld32(tmem_acc + (uint32_t)(tl * 64 + cc) + ((uint32_t)(quad * 32) << 16), raw);
      uint4 sk[4];
#pragma unroll
      for (int c8 = 0; c8 < 4; ++c8) sk[c8] = __ldg(skip + (cc >> 3) + c8);
      fptx::tmem_ld_wait();
#pragma unroll
      for (int c8 = 0; c8 < 4; ++c8) {
        const uint32_t sw[4] = {sk[c8].x, sk[c8].y, sk[c8].z, sk[c8].w};
        uint32_t o[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float2 s2 = unpack2_h16<F16>(sw[j]);
          o[j] = pack2_h16<F16>(__uint_as_float(raw[c8 * 8 + 2 * j]) + s2.x, __uint_as_float(raw[c8 * 8 + 2 * j + 1]) + s2.y);
        }
        fptx::st_shared_v4(row_a + (uint32_t)((((cc >> 3) + c8) ^ ph) << 4), o[0], o[1], o[2], o[3]);
      }
    }
  }
  fptx::tc_fence_before();
  __syncthreads();
  // copy-out: 8 lanes x 16 B per pixel row -> full 128-byte lines of the NHWC output
  {
    const int ch = tid & 7;
    __nv_bfloat16* ob = p.out + (int64_t)b * HW * 64;
    for (int s = tid >> 3; s < nslots; s += NTHR / 8) {
      const int y = __float2int_rd(((float)s + 0.5f) * p.inv_wp);
      const int x = (s - y * Wp) ^ (y & 7);
      if (y < H && x < W) {
        uint32_t w0, w1, w2, w3;
        asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(w0), "=r"(w1), "=r"(w2), "=r"(w3) : "r"(U_a + sw_off(s, ch)));
        *reinterpret_cast<uint4*>(ob + ((int64_t)y * W + x) * 64 + ch * 8) = make_uint4(w0, w1, w2, w3);
      }
    }
  }
  if (warp == 0) {
    fptx::tc_fence_after();
    fptx::tmem_dealloc(tmem_acc, (uint32_t)p.tmem_cols);
  }
}


// ====================================================================================================================
// fablock_full2_kernel: the same block on PRE-STAGED operands with a producer thread (round 2, third session).
//
// The phase trace of fablock_full_kernel (-DLNS_FULL_TRACE, 32x32: 30.8k cycles per head) showed 8.5k cycles per head in
// which the tensor pipes idle: 3.4k per-(sample, head) operand set-up (GroupNorm folded into the in_proj slice, its bias, Kx /
// Ky conversion), 2.7k issuing 26 cp.async per thread for the raw input and the next head's operands (LSU bound), and 2.2k in
// which 511 threads wait at a block barrier for thread 0 to push 32 tcgen05.mma through the (shared-memory bound) tensor
// queue.  Here
//   * lns_fablock_prepass_staged writes the NORMALISED input once per sample as the byte image of U_s (pixel permutation +
//     128-byte swizzle applied): the in_proj slice is sample independent (16-bit, prepared once per parameter version, no
//     bias), and a head's input is four linear 32 KB bulk copies;
//   * thread 0 is the producer: it issues every bulk copy (input quarters, in_proj slice, Kx | Ky, the fp32 to_out[1] slice)
//     and every tcgen05.mma at two points of the head loop and meets the other warps at mbarriers only -- no warp waits at a
//     block barrier for an instruction issue (a 17th producer warp was tried first: correct at 32x32, but five warps on one
//     scheduler cap the kernel at 96 registers, and two co-resident 9-warp CTAs at 16x16 produced wrong samples);
//   * the head's tcgen05 GEMM is committed per quarter of the pixel rows, so the next head's input quarter q is refilled as soon
//     as the MMAs that read quarter q have completed, and phase A of the next head starts on quarter 0 while the tensor core
//     is still working on quarters 1-3.
// Needs H, W in {16, 32} (power-of-two width, no pad rows, no padded kernel-matrix tiles).
struct Full2Params {
  const uint16_t* us;      // staged input [B][H*W][64] 16-bit: GroupNorm applied, pixel rows in U_s order, chunks swizzled
  const __nv_bfloat16* u;  // the block's RAW input [B][H][W][64] (skip connection)
  int H, W, heads;
  const uint16_t* w_in16;  // [heads][64][kWS] 16-bit in_proj slices (rows padded to kWS)
  const float* Kx;         // [B][heads][H][H]
  const float* Ky;         // [B][heads][W][W]
  float eps;
  const float* w1h;        // [heads][64 out][64 in] fp32: to_out[1] slices, head-major
  const float* w_out2;     // [64][64]
  __nv_bfloat16* out;      // [B][H][W][64]
  int T, tmem_cols, lgw;
  long long* trace;
};

struct Full2Smem {
  uint32_t U, Wo, W2, Ws, Kst, W1st, Kx, Ky, obias, red, stat, bar, slot, total;
};
__host__ __device__ inline Full2Smem full2_layout(int H, int W, int T, int nwarp) {
  Full2Smem L;
  uint32_t o = 0;
  L.U = o; o += (uint32_t)T * 16384u;
  L.Wo = o; o += 2u * 8192u;
  L.W2 = o; o += 8192u;
  L.Ws = o; o += 64u * kWS * 2u;
  L.Kst = o; o += (uint32_t)(H * H + W * W) * 4u;
  L.W1st = o; o += 64u * 64u * 4u;
  L.Kx = o; o += (uint32_t)(H * (H + 8) * 2);
  L.Ky = o; o += (uint32_t)(W * (W + 8) * 2);
  o = (o + 15u) & ~15u;
  L.obias = o; o += 64u * 4u;
  L.red = o; o += (uint32_t)nwarp * 64u * 2u * 4u;
  L.stat = o; o += 64u * 2u * 4u;
  L.bar = o; o += 16u * 8u;
  L.slot = o; o += 8u;
  L.total = o;
  return L;
}

namespace fptx {
__device__ __forceinline__ void mbar_arrive(uint32_t bar) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src),
               "r"(bytes), "r"(bar)
               : "memory");
}
}  // namespace fptx

#ifdef LNS_FULL_TRACE
#define LNS_FT2(i) do { if (tid == 0 && blockIdx.x == 0 && h < 4) p.trace[h * 16 + (i)] = clock64(); } while (0)
#else
#define LNS_FT2(i) do { } while (0)
#endif

// grid B (one CTA per sample), block NTHR threads (512 for > 256 pixels, else 256; two CTAs per SM at 16x16)
template <int NTHR, bool F16>
__global__ void __launch_bounds__(NTHR, NTHR == 256 ? 2 : 1) fablock_full2_kernel(const Full2Params p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (fptx::s32(smem_raw) + 1023u) & ~1023u;
  uint8_t* gen = smem_raw + (base - fptx::s32(smem_raw));
  constexpr int nwarp = NTHR / 32;
  constexpr int NQ = NTHR == 256 ? 2 : 4;  // parts of the pixel-row range: input copies, phase A and the MMA commits
  const int H = p.H, W = p.W, T = p.T, heads = p.heads;
  const Full2Smem L = full2_layout(H, W, T, nwarp);
  const int HW = H * W;  // = pixel rows of U_s (no pad rows)
  const int kxs = H + 8, kys = W + 8;
  const uint32_t U_a = base + L.U, W2_a = base + L.W2, Ws_a = base + L.Ws, Kx_a = base + L.Kx, Ky_a = base + L.Ky;
  const uint32_t Kst_a = base + L.Kst, W1st_a = base + L.W1st;
  const uint32_t bar0 = base + L.bar;
  auto bar_u = [&](int q) { return bar0 + (uint32_t)q * 8u; };          // input quarter q has landed (tx)
  const uint32_t bar_ops = bar0 + 32u;                                   // in_proj slice + Kx | Ky have landed (tx)
  const uint32_t bar_w1 = bar0 + 40u;                                    // to_out[1] slice has landed (tx)
  const uint32_t bar_opfree = bar0 + 48u;                                // every warp is done with Ws / Kst (nwarp arrivals)
  const uint32_t bar_e = bar0 + 56u;                                     // every warp has written U_s / Wo_s of this head
  auto bar_mma = [&](int q) { return bar0 + 64u + (uint32_t)q * 8u; };  // the head's MMAs on quarter q have completed
  const uint32_t bar_fin = bar0 + 96u;
  uint16_t* Kx_s = reinterpret_cast<uint16_t*>(gen + L.Kx);
  uint16_t* Ky_s = reinterpret_cast<uint16_t*>(gen + L.Ky);
  float* obias_s = reinterpret_cast<float*>(gen + L.obias);
  float* red_s = reinterpret_cast<float*>(gen + L.red);
  float* stat_s = reinterpret_cast<float*>(gen + L.stat);
  volatile uint32_t* slot_gen = reinterpret_cast<volatile uint32_t*>(gen + L.slot);
  const float* stkx_s = reinterpret_cast<const float*>(gen + L.Kst);  // [H][H]
  const float* stky_s = stkx_s + H * H;                                // [W][W]
  const float* w1st_s = reinterpret_cast<const float*>(gen + L.W1st);  // [64 out][64 in]
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int b = blockIdx.x;
  const __nv_bfloat16* ub = p.u + (int64_t)b * HW * 64;

  if (tid == 0) {
    for (int q = 0; q < 4; ++q) {
      fptx::mbar_init(bar_u(q), 1);
      fptx::mbar_init(bar_mma(q), 1);
    }
    fptx::mbar_init(bar_ops, 1);
    fptx::mbar_init(bar_w1, 1);
    fptx::mbar_init(bar_opfree, nwarp);
    fptx::mbar_init(bar_e, nwarp);
    fptx::mbar_init(bar_fin, 1);
    fptx::fence_mbar_init();
  }
  if (warp == 0) {
    fptx::tmem_alloc(base + L.slot, (uint32_t)p.tmem_cols);
    fptx::tmem_relinquish();
  }
  if (tid < 64) obias_s[tid] = 0.f;
  // to_out[3] filter -> 16-bit swizzled K-major B operand [64 n][64 k]
  for (int e = tid; e < 64 * 8; e += NTHR) {
    const int n = e >> 3, kc = e & 7;
    const float4 w0 = __ldg(reinterpret_cast<const float4*>(p.w_out2 + n * 64 + kc * 8));
    const float4 w1 = __ldg(reinterpret_cast<const float4*>(p.w_out2 + n * 64 + kc * 8 + 4));
    fptx::st_shared_v4(W2_a + sw_off(n, kc), pack2_h16<F16>(w0.x, w0.y), pack2_h16<F16>(w0.z, w0.w), pack2_h16<F16>(w1.x, w1.y),
                       pack2_h16<F16>(w1.z, w1.w));
  }
  fptx::tc_fence_before();
  __syncthreads();
  fptx::tc_fence_after();
  const uint32_t tmem_acc = *slot_gen;
  // instruction descriptor, kind::f16: D = f32, A/B = bf16 (1) or f16 (0), K-major, N = 64, M = 128
  const uint32_t idesc = (1u << 4) | (F16 ? 0u : ((1u << 7) | (1u << 10))) | ((64u >> 3) << 17) | ((128u >> 4) << 24);
  const uint32_t qbytes = (uint32_t)HW * 128u / NQ;
  const int tl_per_q = T / NQ;

  // ---- producer duties: thread 0 only, at points of the head loop where the other warps are not waiting for it ----
  const uint8_t* usb = reinterpret_cast<const uint8_t*>(p.us) + (size_t)b * HW * 128u;
  const uint32_t kxb = (uint32_t)(H * H) * 4u, kyb = (uint32_t)(W * W) * 4u;
  auto issue_ops = [&](int hh) {
    fptx::mbar_expect_tx(bar_ops, 64u * kWS * 2u + kxb + kyb);
    fptx::bulk_g2s(Ws_a, p.w_in16 + (size_t)hh * 64 * kWS, 64u * kWS * 2u, bar_ops);
    fptx::bulk_g2s(Kst_a, p.Kx + ((int64_t)b * heads + hh) * H * H, kxb, bar_ops);
    fptx::bulk_g2s(Kst_a + kxb, p.Ky + ((int64_t)b * heads + hh) * W * W, kyb, bar_ops);
  };
  auto issue_w1 = [&](int hh) {
    fptx::mbar_expect_tx(bar_w1, 64u * 64u * 4u);
    fptx::bulk_g2s(W1st_a, p.w1h + (size_t)hh * 64 * 64, 64u * 64u * 4u, bar_w1);
  };
  auto issue_u = [&](int q) {
    fptx::mbar_expect_tx(bar_u(q), qbytes);
    fptx::bulk_g2s(U_a + (uint32_t)q * qbytes, usb + (size_t)q * qbytes, qbytes, bar_u(q));
  };
  auto issue_mma = [&](int h, int q) {  // the head's to_out[1] GEMM on the pixel rows of quarter q, committed on its own barrier
    const uint64_t bdesc = desc_sw128(base + L.Wo + (uint32_t)(h & 1) * 8192u);
    for (int tl = q * tl_per_q; tl < (q + 1) * tl_per_q; ++tl) {
      const uint64_t adesc = desc_sw128(U_a + (uint32_t)tl * 16384u);
#pragma unroll
      for (int k = 0; k < 4; ++k)
        fptx::umma_f16(tmem_acc + (uint32_t)(tl * 64), adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), idesc, (h | k) != 0 ? 1u : 0u);
    }
    fptx::umma_commit(bar_mma(q));
  };
  if (tid == 0) {
    issue_ops(0);
    for (int q = 0; q < NQ; ++q) issue_u(q);
    issue_w1(0);
  }
  __syncwarp();

  const int t = lane & 3;
  (void)t;
  const int nblk = HW >> 4;
  const int blk_per_q = nblk / NQ;
  for (int h = 0; h < heads; ++h) {
    const uint32_t par = (uint32_t)(h & 1);
    LNS_FT2(0);
    fptx::mbar_wait(bar_ops, par);
    for (int i = warp; i < H; i += nwarp)
      for (int j = lane; j < H; j += 32) Kx_s[i * kxs + j] = to_h16<F16>(stkx_s[i * H + j]);
    for (int i = warp; i < W; i += nwarp)
      for (int j = lane; j < W; j += 32) Ky_s[i * kys + j] = to_h16<F16>(stky_s[i * W + j]);
    uint32_t wf[4][8][2];
#pragma unroll
    for (int ks = 0; ks < 4; ++ks)
#pragma unroll
      for (int nt = 0; nt < 8; ++nt)
        fptx::ldsm_x2(Ws_a + (uint32_t)(((nt * 8 + (lane & 7)) * kWS + ks * 16 + ((lane >> 3) & 1) * 8) * 2), wf[ks][nt][0], wf[ks][nt][1]);
    __syncwarp();
    if (lane == 0) fptx::mbar_arrive(bar_opfree);
    if (tid == 0 && h + 1 < heads) {
      fptx::mbar_wait(bar_opfree, par);  // every warp holds its in_proj fragments and has converted Kx | Ky: refill for head h + 1
      // generic-proxy READS of Ws / Kst by the other warps (ordered before this point by the mbarrier) -> async-proxy WRITES of the
      // bulk copies: the ISSUING thread needs its own proxy fence.  Without it about 1 sample in 10^4 came out wrong at 16x16
      // (two CTAs per SM; never seen at 32x32) -- tools/dbg_staged3.py counts them over many launches.
      fptx::fence_proxy_async();
      issue_ops(h + 1);
    }
    __syncwarp();
    LNS_FT2(1);
    // ---- phase A: u_phi_h = u_n x W_in[h]^T, in place, 16-row blocks per warp; quarter q as soon as it has landed ----
    for (int q = 0; q < NQ; ++q) {
      fptx::mbar_wait(bar_u(q), par);
      if (q == 0) LNS_FT2(2);
      for (int blk = q * blk_per_q + warp; blk < (q + 1) * blk_per_q; blk += nwarp) {
        const int pr = blk * 16 + (lane & 15);
        uint32_t a[4][4];
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) fptx::ldsm_x4(U_a + sw_off(pr, ks * 2 + (lane >> 4)), a[ks][0], a[ks][1], a[ks][2], a[ks][3]);
        float acc[8][4];
#pragma unroll
        for (int nt = 0; nt < 8; ++nt) acc[nt][0] = acc[nt][1] = acc[nt][2] = acc[nt][3] = 0.f;
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) {
#pragma unroll
          for (int nt = 0; nt < 8; ++nt) fptx::mma16816<F16>(acc[nt], a[ks], wf[ks][nt][0], wf[ks][nt][1]);
        }
        __syncwarp();  // every lane's ldmatrix of the raw rows is done before they are overwritten
        const int rs_row = blk * 16 + ((lane >> 3) & 1) * 8 + (lane & 7);
        const uint32_t rs = U_a + (uint32_t)rs_row * 128u;
        const int phs = rs_row & 7, cs = lane >> 4;
#pragma unroll
        for (int nt = 0; nt < 8; nt += 2) {
          fptx::stsm_x4(rs + (uint32_t)(((nt + cs) ^ phs) << 4), pack2_h16<F16>(acc[nt][0], acc[nt][1]),
                        pack2_h16<F16>(acc[nt][2], acc[nt][3]), pack2_h16<F16>(acc[nt + 1][0], acc[nt + 1][1]),
                        pack2_h16<F16>(acc[nt + 1][2], acc[nt + 1][3]));
        }
      }
    }
    __syncthreads();  // (also publishes Kx_s / Ky_s)
    LNS_FT2(3);
    // ---- phase B: contraction over H, one image column per warp at a time ----
    for (int m = warp; m < W; m += nwarp) {
      if (H == 16) contract_line_sw<1, 0, F16>(U_a, m, W, H, Kx_a, kxs, lane);
      else contract_line_sw<2, 0, F16>(U_a, m, W, H, Kx_a, kxs, lane);
    }
    __syncthreads();
    LNS_FT2(4);
    // ---- phase C: contraction over W, one image row per warp ----
    for (int i = warp; i < H; i += nwarp) {
      if (W == 16) contract_line_sw<1, 1, F16>(U_a, i, W, W, Ky_a, kys, lane);
      else contract_line_sw<2, 1, F16>(U_a, i, W, W, Ky_a, kys, lane);
    }
    __syncthreads();
    LNS_FT2(5);
    // ---- phase D: InstanceNorm statistics of this head's 64 channels, from the 16-bit values the GEMM will read ----
    {
      const int ch = tid & 7;
      float sm8[8], sq8[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) sm8[j] = sq8[j] = 0.f;
      for (int s = tid >> 3; s < HW; s += NTHR / 2) {  // four 16-byte loads in flight per thread
        uint32_t w[4][4];
#pragma unroll
        for (int q4 = 0; q4 < 4; ++q4)
          asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(w[q4][0]), "=r"(w[q4][1]), "=r"(w[q4][2]), "=r"(w[q4][3])
                       : "r"(U_a + sw_off(s + q4 * (NTHR / 8), ch)));
#pragma unroll
        for (int q4 = 0; q4 < 4; ++q4)
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const float2 f = unpack2_h16<F16>(w[q4][j]);
            sm8[2 * j] += f.x; sq8[2 * j] = fmaf(f.x, f.x, sq8[2 * j]);
            sm8[2 * j + 1] += f.y; sq8[2 * j + 1] = fmaf(f.y, f.y, sq8[2 * j + 1]);
          }
      }
#pragma unroll
      for (int j = 0; j < 8; ++j) {  // the 4 pixel lanes of a warp that share this channel chunk
        sm8[j] += __shfl_xor_sync(0xffffffffu, sm8[j], 8);
        sm8[j] += __shfl_xor_sync(0xffffffffu, sm8[j], 16);
        sq8[j] += __shfl_xor_sync(0xffffffffu, sq8[j], 8);
        sq8[j] += __shfl_xor_sync(0xffffffffu, sq8[j], 16);
      }
      if (lane < 8) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          red_s[(warp * 64 + ch * 8 + j) * 2 + 0] = sm8[j];
          red_s[(warp * 64 + ch * 8 + j) * 2 + 1] = sq8[j];
        }
      }
    }
    __syncthreads();
    LNS_FT2(6);
    if (tid < 64) {
      double sm = 0.0, ss = 0.0;
      for (int w = 0; w < nwarp; ++w) {
        sm += (double)red_s[(w * 64 + tid) * 2 + 0];
        ss += (double)red_s[(w * 64 + tid) * 2 + 1];
      }
      const double inv_n = 1.0 / (double)HW;
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
    fptx::mbar_wait(bar_w1, par);
    __syncthreads();
    LNS_FT2(7);
    // ---- phase E: fold the normalisation into to_out[1]'s slice (staged in shared memory); the producer issues the GEMM ----
    const uint32_t Wo_a = base + L.Wo + (uint32_t)(h & 1) * 8192u;
    for (int e = tid; e < 64 * 8; e += NTHR) {
      const int n = e >> 3, kc = e & 7;
      const float4 w0 = *reinterpret_cast<const float4*>(w1st_s + n * 64 + kc * 8);
      const float4 w1 = *reinterpret_cast<const float4*>(w1st_s + n * 64 + kc * 8 + 4);
      const float wv[8] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w};
      float sc[8];
      float bsum = 0.f;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        sc[j] = wv[j] * stat_s[(kc * 8 + j) * 2 + 0];
        bsum = fmaf(wv[j], stat_s[(kc * 8 + j) * 2 + 1], bsum);
      }
      fptx::st_shared_v4(Wo_a + sw_off(n, kc), pack2_h16<F16>(sc[0], sc[1]), pack2_h16<F16>(sc[2], sc[3]), pack2_h16<F16>(sc[4], sc[5]),
                         pack2_h16<F16>(sc[6], sc[7]));
      bsum += __shfl_xor_sync(0xffffffffu, bsum, 1);
      bsum += __shfl_xor_sync(0xffffffffu, bsum, 2);
      bsum += __shfl_xor_sync(0xffffffffu, bsum, 4);
      if (kc == 0) obias_s[n] += bsum;
    }
    fptx::fence_proxy_async();  // generic-proxy writes of U_s / Wo_s -> visible to the tensor core's async-proxy reads
    fptx::tc_fence_before();
    __syncwarp();
    if (lane == 0) fptx::mbar_arrive(bar_e);
    if (tid == 0) {
      // The other warps go on to the next head's operand set-up and wait for its input; nobody waits for this thread except
      // through the data it moves.  MMAs of quarter q + 1 are queued before quarter q's completion is awaited, so the tensor
      // core stays fed while the input of the next head streams in behind it.
      fptx::mbar_wait(bar_e, par);
      fptx::fence_proxy_async();  // (as above: the other warps' reads of W1st, their writes of U_s / Wo_s -> bulk copies / MMAs issued here)
      fptx::tc_fence_after();
      const bool more = h + 1 < heads;
      if (more) issue_w1(h + 1);
      issue_mma(h, 0);
      for (int q = 1; q < NQ; ++q) {
        issue_mma(h, q);
        if (more) {
          fptx::mbar_wait(bar_mma(q - 1), par);
          issue_u(q - 1);
        }
      }
      if (more) {
        fptx::mbar_wait(bar_mma(NQ - 1), par);
        issue_u(NQ - 1);
      }
    }
    __syncwarp();
    LNS_FT2(8);
  }

  // ================= after the last head: GELU(acc + b1') -> to_out[3] -> + skip -> out =================
  fptx::mbar_wait(bar_mma(NQ - 1), (uint32_t)((heads - 1) & 1));
  fptx::tc_fence_after();
  __syncthreads();  // obias_s of the last head is complete
  const int quad = warp & 3, sub = warp >> 2, nsub = nwarp >> 2;
  for (int tl = sub; tl < T; tl += nsub) {
    const int pr = tl * 128 + quad * 32 + lane;
    const uint32_t row_a = U_a + (uint32_t)pr * 128u;
    const int ph = pr & 7;
#pragma unroll
    for (int cc = 0; cc < 64; cc += 32) {
      uint32_t raw[32];
      fptx::tmem_ld32(tmem_acc + (uint32_t)(tl * 64 + cc) + ((uint32_t)(quad * 32) << 16), raw);
      fptx::tmem_ld_wait();
#pragma unroll
      for (int c8 = 0; c8 < 4; ++c8) {
        float v[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) v[j] = act_gelu_fast(__uint_as_float(raw[c8 * 8 + j]) + obias_s[cc + c8 * 8 + j]);
        fptx::st_shared_v4(row_a + (uint32_t)((((cc >> 3) + c8) ^ ph) << 4), pack2_h16<F16>(v[0], v[1]), pack2_h16<F16>(v[2], v[3]),
                           pack2_h16<F16>(v[4], v[5]), pack2_h16<F16>(v[6], v[7]));
      }
    }
  }
  fptx::fence_proxy_async();
  fptx::tc_fence_before();
  __syncthreads();
  if (tid == 0) {
    fptx::tc_fence_after();
    const uint64_t bdesc = desc_sw128(W2_a);
    for (int tl = 0; tl < T; ++tl) {
      const uint64_t adesc = desc_sw128(U_a + (uint32_t)tl * 16384u);
#pragma unroll
      for (int k = 0; k < 4; ++k)
        fptx::umma_f16(tmem_acc + (uint32_t)(tl * 64), adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), idesc, k != 0 ? 1u : 0u);
    }
    fptx::umma_commit(bar_fin);
  }
  fptx::mbar_wait(bar_fin, 0u);
  fptx::tc_fence_after();
  // epilogue 2: + skip (the block's raw input, fp32 add), 16-bit result back into the pixel's own row of U_s
  for (int tl = sub; tl < T; tl += nsub) {
    const int pr = tl * 128 + quad * 32 + lane;
    const int src = pr ^ ((pr >> p.lgw) & 7);  // the pixel that lives in row pr
    const uint4* skip = reinterpret_cast<const uint4*>(ub + (int64_t)src * 64);
    const uint32_t row_a = U_a + (uint32_t)pr * 128u;
    const int ph = pr & 7;
#pragma unroll
    for (int cc = 0; cc < 64; cc += 32) {
      uint32_t raw[32];
      fptx::tmem_ld32(tmem_acc + (uint32_t)(tl * 64 + cc) + ((uint32_t)(quad * 32) << 16), raw);
      uint4 sk[4];
#pragma unroll
      for (int c8 = 0; c8 < 4; ++c8) sk[c8] = __ldg(skip + (cc >> 3) + c8);
      fptx::tmem_ld_wait();
#pragma unroll
      for (int c8 = 0; c8 < 4; ++c8) {
        const uint32_t sw[4] = {sk[c8].x, sk[c8].y, sk[c8].z, sk[c8].w};
        uint32_t o[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float2 s2 = unpack2_h16<F16>(sw[j]);
          o[j] = pack2_h16<F16>(__uint_as_float(raw[c8 * 8 + 2 * j]) + s2.x, __uint_as_float(raw[c8 * 8 + 2 * j + 1]) + s2.y);
        }
        fptx::st_shared_v4(row_a + (uint32_t)((((cc >> 3) + c8) ^ ph) << 4), o[0], o[1], o[2], o[3]);
      }
    }
  }
  fptx::tc_fence_before();
  __syncthreads();
  // copy-out: 8 lanes x 16 B per pixel row -> full 128-byte lines of the NHWC output
  {
    const int ch = tid & 7;
    uint8_t* ob = reinterpret_cast<uint8_t*>(p.out + (int64_t)b * HW * 64) + ch * 16;
    for (int s = tid >> 3; s < HW; s += NTHR / 8) {
      uint32_t w0, w1, w2, w3;
      asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(w0), "=r"(w1), "=r"(w2), "=r"(w3) : "r"(U_a + sw_off(s, ch)));
      *reinterpret_cast<uint4*>(ob + (size_t)(s ^ ((s >> p.lgw) & 7)) * 128u) = make_uint4(w0, w1, w2, w3);
    }
  }
  if (warp == 0) {
    fptx::tc_fence_after();
    fptx::tmem_dealloc(tmem_acc, (uint32_t)p.tmem_cols);
  }
}

}  // namespace lns

extern "C" {

int lns_fablock_full_supported(int H, int W, int dim, int dim_head, int dim_out) {
  if (dim != 64 || dim_head != 64 || dim_out != 64 || H < 1 || W < 1 || H > 32 || W > 32) return 0;
  const int Wp = (W + 7) & ~7;
  const int T = (H * Wp + 127) / 128;
  if (T > 8) return 0;
  return lns::full_layout(H, W, T, 16).total + 1024 <= 227 * 1024;
}

int lns_fablock_full(const void* u, int dtype, int B, int H, int W, int heads, const float* gn_scale, const float* gn_shift,
                     const float* w_in_proj, const float* Kx, const float* Ky, float eps, const float* w_out1, const float* w_out2,
                     void* out, void* stream) {
  LNS_REQUIRE(u && gn_scale && gn_shift && w_in_proj && Kx && Ky && w_out1 && w_out2 && out && B > 0 && heads > 0,
              "lns_fablock_full: bad arguments");
  LNS_REQUIRE(lns::is_h16_host(dtype), "lns_fablock_full: u/out must be LNS_BF16 or LNS_F16 (got dtype %d)", dtype);
  LNS_REQUIRE(lns_fablock_full_supported(H, W, 64, 64, 64), "lns_fablock_full: %dx%d does not fit (use lns_fablock_core + lns_conv2d)", H, W);
  LNS_REQUIRE((reinterpret_cast<uintptr_t>(u) & 15) == 0 && (reinterpret_cast<uintptr_t>(out) & 15) == 0 &&
                  (reinterpret_cast<uintptr_t>(w_in_proj) & 15) == 0 && (reinterpret_cast<uintptr_t>(w_out1) & 15) == 0 &&
                  (reinterpret_cast<uintptr_t>(w_out2) & 15) == 0 && (reinterpret_cast<uintptr_t>(gn_scale) & 15) == 0 &&
                  (reinterpret_cast<uintptr_t>(gn_shift) & 15) == 0,
              "lns_fablock_full: pointers must be 16-byte aligned");
  lns::FullParams p;
  p.u = reinterpret_cast<const __nv_bfloat16*>(u);
  p.H = H; p.W = W; p.heads = heads;
  p.gn_scale = gn_scale; p.gn_shift = gn_shift; p.w_in = w_in_proj; p.Kx = Kx; p.Ky = Ky; p.eps = eps;
  p.w_out1 = w_out1; p.w_out2 = w_out2;
  p.out = reinterpret_cast<__nv_bfloat16*>(out);
  p.Wp = (W + 7) & ~7;
  p.T = (H * p.Wp + 127) / 128;
  p.inv_wp = 1.0f / (float)p.Wp;
  p.k16 = ((H * H) % 4 == 0 && (W * W) % 4 == 0 && (reinterpret_cast<uintptr_t>(Kx) & 15) == 0 && (reinterpret_cast<uintptr_t>(Ky) & 15) == 0) ? 1 : 0;
  p.lgwp = -1;
  if (p.Wp == W && (W & (W - 1)) == 0 && H * W == p.T * 128) {
    int lg = 0;
    while ((1 << lg) < W) ++lg;
    p.lgwp = lg;
  }
  p.trace = nullptr;
#ifdef LNS_FULL_TRACE
  {
    static long long* dbg = nullptr;
    if (!dbg) cudaMalloc(&dbg, 64 * sizeof(long long));
    p.trace = dbg;
    cudaMemsetAsync(dbg, 0, 64 * sizeof(long long), reinterpret_cast<cudaStream_t>(stream));
  }
#endif
  int cols = 32;
  while (cols < p.T * 64) cols <<= 1;
  p.tmem_cols = cols;
  const bool big = H * W > 256;
  const size_t smem = lns::full_layout(H, W, p.T, big ? 16 : 8).total + 1024;
  {
    LNS_OPT_IN_SMEM((lns::fablock_full_kernel<512, false>), 227 * 1024, "fablock_full");
    LNS_OPT_IN_SMEM((lns::fablock_full_kernel<256, false>), 227 * 1024, "fablock_full");
    LNS_OPT_IN_SMEM((lns::fablock_full_kernel<512, true>), 227 * 1024, "fablock_full");
    LNS_OPT_IN_SMEM((lns::fablock_full_kernel<256, true>), 227 * 1024, "fablock_full");
  }
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const bool f16 = dtype == LNS_F16;
  if (big) {
    auto kern = f16 ? lns::fablock_full_kernel<512, true> : lns::fablock_full_kernel<512, false>;
    kern<<<B, 512, smem, st>>>(p);
  } else {
    auto kern = f16 ? lns::fablock_full_kernel<256, true> : lns::fablock_full_kernel<256, false>;
    kern<<<B, 256, smem, st>>>(p);
  }
#ifdef LNS_FULL_TRACE
  {
    long long hbuf[64];
    cudaStreamSynchronize(st);
    cudaMemcpy(hbuf, p.trace, sizeof(hbuf), cudaMemcpyDeviceToHost);
    static int printed = 0;
    if (printed++ < 2)
      for (int h = 1; h < 4; ++h) {
        fprintf(stderr, "fablock_full trace head %d (%dx%d):", h, H, W);
        for (int i = 1; i < 11; ++i) fprintf(stderr, " s%d->%d %lld", i - 1, i, hbuf[h * 16 + i] - hbuf[h * 16 + i - 1]);
        fprintf(stderr, " | head total %lld\n", hbuf[h * 16 + 10] - hbuf[(h - 1) * 16 + 10]);
      }
  }
#endif
  return lns::check_launch("fablock_full_kernel");
}

int lns_fablock_full_staged_supported(int H, int W, int dim, int dim_head, int dim_out) {
  if (dim != 64 || dim_head != 64 || dim_out != 64) return 0;
  if ((H != 16 && H != 32) || (W != 16 && W != 32)) return 0;
  return 1;
}

int lns_fablock_full_staged(const void* u_staged, const void* u, int dtype, int B, int H, int W, int heads, const void* w_in16,
                            const float* Kx, const float* Ky, float eps, const float* w1h, const float* w_out2, void* out,
                            void* stream) {
  LNS_REQUIRE(u_staged && u && w_in16 && Kx && Ky && w1h && w_out2 && out && B > 0 && heads > 0, "lns_fablock_full_staged: bad arguments");
  LNS_REQUIRE(lns::is_h16_host(dtype), "lns_fablock_full_staged: u/out must be LNS_BF16 or LNS_F16 (got dtype %d)", dtype);
  LNS_REQUIRE(lns_fablock_full_staged_supported(H, W, 64, 64, 64), "lns_fablock_full_staged: %dx%d not covered (use lns_fablock_full)", H, W);
  LNS_REQUIRE((reinterpret_cast<uintptr_t>(u_staged) & 15) == 0 && (reinterpret_cast<uintptr_t>(u) & 15) == 0 &&
                  (reinterpret_cast<uintptr_t>(out) & 15) == 0 && (reinterpret_cast<uintptr_t>(w_in16) & 15) == 0 &&
                  (reinterpret_cast<uintptr_t>(w1h) & 15) == 0 && (reinterpret_cast<uintptr_t>(w_out2) & 15) == 0 &&
                  (reinterpret_cast<uintptr_t>(Kx) & 15) == 0 && (reinterpret_cast<uintptr_t>(Ky) & 15) == 0,
              "lns_fablock_full_staged: pointers must be 16-byte aligned");
  lns::Full2Params p;
  p.us = reinterpret_cast<const uint16_t*>(u_staged);
  p.u = reinterpret_cast<const __nv_bfloat16*>(u);
  p.H = H; p.W = W; p.heads = heads;
  p.w_in16 = reinterpret_cast<const uint16_t*>(w_in16);
  p.Kx = Kx; p.Ky = Ky; p.eps = eps; p.w1h = w1h; p.w_out2 = w_out2;
  p.out = reinterpret_cast<__nv_bfloat16*>(out);
  p.T = H * W / 128;
  p.lgw = W == 32 ? 5 : 4;
  p.trace = nullptr;
#ifdef LNS_FULL_TRACE
  {
    static long long* dbg = nullptr;
    if (!dbg) cudaMalloc(&dbg, 64 * sizeof(long long));
    p.trace = dbg;
    cudaMemsetAsync(dbg, 0, 64 * sizeof(long long), reinterpret_cast<cudaStream_t>(stream));
  }
#endif
  int cols = 32;
  while (cols < p.T * 64) cols <<= 1;
  p.tmem_cols = cols;
  const bool big = H * W > 256;
  const size_t smem = lns::full2_layout(H, W, p.T, big ? 16 : 8).total + 1024;
  {
    LNS_OPT_IN_SMEM((lns::fablock_full2_kernel<512, false>), 227 * 1024, "fablock_full_staged");
    LNS_OPT_IN_SMEM((lns::fablock_full2_kernel<256, false>), 227 * 1024, "fablock_full_staged");
    LNS_OPT_IN_SMEM((lns::fablock_full2_kernel<512, true>), 227 * 1024, "fablock_full_staged");
    LNS_OPT_IN_SMEM((lns::fablock_full2_kernel<256, true>), 227 * 1024, "fablock_full_staged");
  }
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const bool f16 = dtype == LNS_F16;
  if (big) {
    auto kern = f16 ? lns::fablock_full2_kernel<512, true> : lns::fablock_full2_kernel<512, false>;
    kern<<<B, 512, smem, st>>>(p);
  } else {
    auto kern = f16 ? lns::fablock_full2_kernel<256, true> : lns::fablock_full2_kernel<256, false>;
    kern<<<B, 256, smem, st>>>(p);
  }
#ifdef LNS_FULL_TRACE
  {
    long long hbuf[64];
    cudaStreamSynchronize(st);
    cudaMemcpy(hbuf, p.trace, sizeof(hbuf), cudaMemcpyDeviceToHost);
    static int printed = 0;
    if (printed++ < 2)
      for (int h = 1; h < 4; ++h) {
        fprintf(stderr, "fablock_full2 trace head %d (%dx%d):", h, H, W);
        for (int i = 1; i < 9; ++i) fprintf(stderr, " s%d->%d %lld", i - 1, i, hbuf[h * 16 + i] - hbuf[h * 16 + i - 1]);
        fprintf(stderr, " | head total %lld\n", hbuf[h * 16 + 8] - hbuf[(h - 1) * 16 + 8]);
      }
  }
#endif
  return lns::check_launch("fablock_full2_kernel");
}

}  // extern "C"
