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
// Warp roles (320 threads): warps 0-3 halo producers, warps 4-5 MMA issuers (warp 4 owns TMEM), warps 6-9 epilogue.
// Pipelines: halo ring (HS stages; full = 128 async cp.async arrivals, empty = tcgen05.commit) and four TMEM
// accumulators (full = tcgen05.commit, empty = 128 epilogue arrivals): epilogue, MMA issue, tensor pipe and halo loads
// of different tiles all overlap.
#include <cuda.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

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
// 16-byte copy from a 64-bit global ADDRESS; `zero` = ignore the source and write zeros (the address must still be valid)
__device__ __forceinline__ void cp_async16_zf(uint32_t dst, uint64_t src, bool zero) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %2, 0;\n\t"
      "cp.async.cg.shared.global [%0], [%1], 16, p;\n\t"
      "}" ::"r"(dst),
      "l"(src), "r"((uint32_t)zero)
      : "memory");
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
// TMA tensor store shared -> global (4-D tiled map, box = one staged tile), tracked by the issuing thread's bulk groups
__device__ __forceinline__ void tma_store_4d(const CUtensorMap* tm, uint32_t src, int c0, int c1, int c2, int c3) {
  asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%1, %2, %3, %4}], [%5];" ::"l"(
                   reinterpret_cast<uint64_t>(tm)),
               "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(src)
               : "memory");
}
// TMA tensor load global -> shared (4-D tiled map), completion counted in bytes on an mbarrier
__device__ __forceinline__ void tma_load_4d(const CUtensorMap* tm, uint32_t dst, uint32_t bar, int c0, int c1, int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];" ::"r"(dst),
      "l"(reinterpret_cast<uint64_t>(tm)), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(bar)
      : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read1() { asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
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
// Leader-predicated forms: the WHOLE warp runs the issue loop (warp-uniform control flow, so descriptor arithmetic can
// live in the uniform datapath) and only the elected lane's instruction takes effect.
__device__ __forceinline__ void umma_bf16_if(uint32_t leader, uint32_t d_tmem, uint32_t a_lo, uint32_t a_hi, uint32_t b_lo,
                                             uint32_t b_hi, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p, q;\n\t"
      ".reg .b64 da, db;\n\t"
      "setp.ne.b32 p, %6, 0;\n\t"
      "setp.ne.b32 q, %7, 0;\n\t"
      "mov.b64 da, {%1, %2};\n\t"
      "mov.b64 db, {%3, %4};\n\t"
      "@q tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %5, p;\n\t"
      "}" ::"r"(d_tmem),
      "r"(a_lo), "r"(a_hi), "r"(b_lo), "r"(b_hi), "r"(idesc), "r"(accumulate), "r"(leader)
      : "memory");
}
__device__ __forceinline__ void umma_commit_if(uint32_t leader, uint32_t bar) {
  asm volatile(
      "{\n\t"
      ".reg .pred q;\n\t"
      "setp.ne.b32 q, %1, 0;\n\t"
      "@q tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n\t"
      "}" ::"r"(bar),
      "r"(leader)
      : "memory");
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
// 32 lanes x 32 consecutive fp32 columns -> 32 registers per thread
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
  int x_f16;  // operands (input, filter) are IEEE half instead of bf16
  int tiles_x, tiles_y, ntiles;
  int HH, HW;          // halo height / width in pixels: 16 + 2d, 8 + 2d
  int halo_bytes;      // per stage, multiple of 1024
  int stages;          // halo ring depth
  int resize;          // 0 none, 1 exact 2x nearest, 2 general nearest
  int nbuf;            // output staging buffers per epilogue warp (4 KB each): 2 when they fit, else 1
  int tma_store;       // 16-bit outputs leave through one TMA tensor store per staged tile (tmap_y is valid)
  int tma_res;         // the 16-bit residual tile arrives through one TMA tensor load per epilogue warp (tmap_r is valid)
  float inv_hw, inv_hv, inv_wv;
  int debug;  // LNS_HALO_DEBUG bits (timing experiments only): 1 skip halo copies, 2 skip epilogue, 4 skip MMAs
};

// kIssuers MMA issue warps (tile it -> issuer it % kIssuers), kAccs TMEM accumulator buffers (tile it -> it % kAccs)
constexpr int kTileH = 16, kTileW = 8;

// One producer warp fills one halo stage for output tile (tx, ty) of sample b.
// Copy schedule: a halo row is HW pixels x 8 sixteen-byte chunks = HW*8 (80 / 96 / 112) slots; lane l owns slots l, l+32,
// l+64(, l+96) of EVERY row, i.e. fixed halo columns hx = (l + 32*sub) >> 3 and the fixed chunk l & 7.  Per tile the lane
// evaluates its <= 4 column terms (as 64-bit column base addresses) and row term `lane` once; per row it needs one shuffle
// (the row term) and per copy a 64-bit add and a predicate.  Everything about the destination is a compile-time constant
// plus a per-lane term.  (ncu source view of the previous version -- two shuffles, a 64-bit pointer select and a wrapping
// (hy, hx) walk per copy: 38 SASS instructions per pass x 45 passes per tile issued by ONE warp -- showed the producers
// 62% of their time in the copy loop with BOTH the MMA issuers and the epilogue waiting on them: ~8000 cycles of producer
// issue per tile against 1184 tensor-pipe cycles.)
template <int DIL>
__device__ __forceinline__ void halo_fill(const HaloParams& p, int tx, int ty, int b, int lane, uint32_t st, uint32_t empty_bar,
                                          bool wait_empty, uint32_t empty_parity, bool docopy) {
  constexpr int HW = kTileW + 2 * DIL, HH = kTileH + 2 * DIL;
  constexpr int NSLOT = HW * 8, NSUB = (NSLOT + 31) / 32;
  const ConvGeom& g = p.g;
  const int chunk = lane & 7;
  // source terms (BYTES): row term of halo row `lane` (HH <= 22); -1 marks "outside -> zero fill"
  int rowterm;
  {
    int yv = ty * kTileH - DIL + lane;
    if (g.circ_h) {
      yv += (yv < 0) ? g.Hv : 0;
      yv -= (yv >= g.Hv) ? g.Hv : 0;
    }
    const bool yok = (unsigned)yv < (unsigned)g.Hv;
    if (p.resize == 1) yv >>= 1;
    else if (p.resize == 2) yv = __float2int_rd(((float)(yv * g.Hin) + 0.5f) * p.inv_hv);
    rowterm = yok ? yv * g.Win * 128 : -1;
  }
  // sample base + this lane's chunk (+ column term) as 64-bit addresses; the plain base is also the valid dummy source of
  // a zero-filled copy
  const uint64_t xb64 = reinterpret_cast<uint64_t>(p.x + (int64_t)b * g.x_bstride + chunk * 8);
  uint64_t cbase[NSUB];
  bool cok[NSUB];
#pragma unroll
  for (int sub = 0; sub < NSUB; ++sub) {
    int xv = tx * kTileW - DIL + ((lane + 32 * sub) >> 3);
    if (g.circ_w) {
      xv += (xv < 0) ? g.Wv : 0;
      xv -= (xv >= g.Wv) ? g.Wv : 0;
    }
    cok[sub] = (unsigned)xv < (unsigned)g.Wv;
    if (p.resize == 1) xv >>= 1;
    else if (p.resize == 2) xv = __float2int_rd(((float)(xv * g.Win) + 0.5f) * p.inv_wv);
    cbase[sub] = xb64 + (cok[sub] ? (uint64_t)(uint32_t)(xv * 128) : 0ull);
    asm volatile("" : "+l"(cbase[sub]));  // opaque: one 64-bit add per copy, no re-derivation from the parts
  }
  if (wait_empty) hptx::mbar_wait(empty_bar, empty_parity);
  if (!docopy) return;
#pragma unroll
  for (int hy = 0; hy < HH; ++hy) {
    const int rt = __shfl_sync(0xFFFFFFFFu, rowterm, hy);
    const bool rok = rt >= 0;
    const uint64_t rtc = rok ? (uint64_t)(uint32_t)rt : 0ull;
#pragma unroll
    for (int sub = 0; sub < NSUB; ++sub) {
      const int sl = lane + 32 * sub;
      const int q = hy * HW + (sl >> 3);  // halo pixel = shared-memory row
      const uint32_t dst = st + (uint32_t)q * 128u + (uint32_t)((chunk ^ (q & 7)) << 4);
      if ((sub + 1) * 32 <= NSLOT || sl < NSLOT) hptx::cp_async16_zf(dst, cbase[sub] + rtc, !(rok && cok[sub]));
    }
  }
}

// The epilogue warps' loop, compiled twice per kernel: TRES = the 16-bit residual arrives by TMA (see below).  Two
// instantiations rather than a runtime flag: with both paths in one body the residual-free layers lost 20 % (1076 -> 854
// TFLOP/s at 64 -> 64 @ 64x64) to the register pressure of the residual path.
template <int NT, int kIssuers, int kAccs, bool TRES, bool FAST>
__device__ __forceinline__ void halo_epilogue(const HaloParams& p, const CUtensorMap& tmap_y, const CUtensorMap& tmap_r, uint32_t tmem_acc,
                                              uint32_t stage_out, uint32_t res_stage0, uint32_t acc_full0, uint32_t acc_empty0,
                                              uint32_t res_full0, int warp, int lane) {
  const ConvGeom& g = p.g;
  auto acc_full = [&](int a) { return acc_full0 + 8u * a; };
  auto acc_empty = [&](int a) { return acc_empty0 + 8u * a; };
  auto res_full = [&](int w) { return res_full0 + 8u * w; };
  // TMEM lane = tile pixel (row m = 8*tile_row + tile_col), so a thread owns one pixel's channels.  Writing them
  // straight to global memory makes every store instruction touch 32 different 128-byte lines.  bf16 outputs are
  // therefore staged per 64-channel group through a per-warp 32 x 128 B tile (16-byte chunks XOR-swizzled with the
  // row so both the row-wise writes and the line-wise reads are conflict free) and leave as full-line stores:
  // 8 consecutive rows = 8 consecutive pixels = 1 KB contiguous in NHWC.
  const int quad = warp & 3;  // TMEM lane quadrant this warp may access
  // FAST: 16-bit output through TMA, no per-sample bias / pre-activation addend / directly loaded residual -- the paths
  // that do not exist are not compiled (the epilogue warps of the generic body lost 36 % of their samples to instruction
  // fetch, ncu `no_inst`)
  const bool y16 = FAST ? true : is_h16(p.y_dtype);
  const int m = quad * 32 + lane;
  const int ty_l = m >> 3, tx_l = m & 7;
  const uint32_t my_stage0 = stage_out + (uint32_t)(warp - (4 + kIssuers)) * (uint32_t)p.nbuf * 4096u;
  const bool tma = FAST ? true : (p.tma_store != 0);
  uint32_t nstore = 0;  // staged tiles written by this warp so far (selects the staging buffer)
  // residual through TMA (Cout = 64 only: one channel group per tile): the warp's 32 pixel rows x 128 B land in its own
  // 4 KB tile, in the same row order and swizzle as the output staging, while the accumulator is still being computed.
  // (Reading the residual straight from global memory cost 16 dependent 8-byte loads per thread, each touching 32 lines:
  // the `res` layers ran 2.3x slower than the same layer without a residual.)
  constexpr bool tres = TRES;
  const int ew = warp - (4 + kIssuers);
  const uint32_t my_res = res_stage0 + (uint32_t)ew * 4096u;
  const uint32_t my_res_row = my_res + (uint32_t)lane * 128u;
  uint32_t res_phase = 0;
  const int rd_row = lane >> 3, rd_chunk = lane & 7;  // read-back: 4 rows per pass, 8 lanes x 16 B per row
  int it = 0;
  int tx, ty, b;
  {
    const int t0 = blockIdx.x;
    tx = t0 % p.tiles_x;
    const int t2 = t0 / p.tiles_x;
    ty = t2 % p.tiles_y;
    b = t2 / p.tiles_y;
  }
  const int eadv = (int)gridDim.x;
  const int eadv_x = eadv % p.tiles_x, eadv_t2 = eadv / p.tiles_x;
  const int eadv_y = eadv_t2 % p.tiles_y, eadv_b = eadv_t2 / p.tiles_y;
  for (int tile = blockIdx.x; tile < p.ntiles; tile += gridDim.x, ++it) {
    const int a = it % kAccs;
    const int yo = ty * kTileH + ty_l, xo = tx * kTileW + tx_l;
    const bool row_ok = yo < g.Hout && xo < g.Wout;
    const int64_t pix = (int64_t)yo * g.Wout + xo;
    const int64_t yrow = (int64_t)b * g.y_bstride + pix * g.Cout;
    const int64_t prow = (int64_t)b * p.pre_add_bstride + pix * g.Cout;
    const int64_t rrow = (int64_t)b * p.res_bstride + pix * g.Cout;
    const bool res_box = tres && (ty * kTileH + quad * 4) < g.Hout && (tx * kTileW) < g.Wout;  // warp uniform
    if (res_box) {
      __syncwarp();  // every lane has finished reading the previous tile's residual
      if (lane == 0) {
        hptx::mbar_expect_tx(res_full(ew), 4096u);
        hptx::tma_load_4d(&tmap_r, my_res, res_full(ew), 0, tx * kTileW, ty * kTileH + quad * 4, b);
      }
    }
    hptx::mbar_wait(acc_full(a), (it / kAccs) & 1);
    hptx::tc_fence_after();
    const uint32_t t_lane = tmem_acc + (uint32_t)(a * NT) + ((uint32_t)(quad * 32) << 16);
    for (int cg = 0; cg < ((p.debug & 2) ? 0 : g.Cout); cg += 64) {
      // staging buffer of this (tile, channel group): with the TMA store the buffer is still being READ by the store issued
      // nbuf groups ago -- lane 0 (the issuer, whose bulk groups track it) waits for that read before anyone overwrites it
      const uint32_t my_stage = my_stage0 + ((p.nbuf == 2) ? (nstore & 1u) * 4096u : 0u);
      const uint32_t my_row_st = my_stage + (uint32_t)lane * 128u;
      ++nstore;
      if (tma && y16) {
        if (lane == 0) {
          if (p.nbuf == 2) hptx::bulk_wait_read1();
          else hptx::bulk_wait_read0();
        }
        __syncwarp();
      }
#pragma unroll
      for (int cc = 0; cc < 64; cc += 32) {
        const int c0 = cg + cc;
        uint32_t raw[32];
        __syncwarp();
        hptx::tmem_ld32(t_lane + (uint32_t)c0, raw);
        hptx::tmem_ld_wait();
        float v[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(raw[j]);
        // (bias: already in the accumulator -- folded into the GEMM by the issue warps)
        if (row_ok) {
          if (!FAST && p.sample_bias) {
#pragma unroll
            for (int j = 0; j < 32; j += 4) {
              float4 t = __ldg(reinterpret_cast<const float4*>(p.sample_bias + (int64_t)b * g.Cout + c0 + j));
              v[j] += t.x; v[j + 1] += t.y; v[j + 2] += t.z; v[j + 3] += t.w;
            }
          }
          if (!FAST && p.pre_add) {
#pragma unroll
            for (int j = 0; j < 32; j += 4) {
              float4 t = ld4_as_float(p.pre_add, p.pre_add_dtype, prow + c0 + j);
              v[j] += t.x; v[j + 1] += t.y; v[j + 2] += t.z; v[j + 3] += t.w;
            }
          }
        }
        if (p.act != LNS_ACT_NONE) {
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = apply_act_fast(v[j], p.act);
        }
        if (tres) {
          if (res_box) {
            if (cc == 0) {
              hptx::mbar_wait(res_full(ew), res_phase);
              res_phase ^= 1u;
            }
#pragma unroll
            for (int h4 = 0; h4 < 4; ++h4) {
              uint32_t w0, w1, w2, w3;
              asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(w0), "=r"(w1), "=r"(w2), "=r"(w3)
                           : "r"(my_res_row + (uint32_t)((((cc >> 3) + h4) ^ (lane & 7)) << 4)));
              const uint32_t ww[4] = {w0, w1, w2, w3};
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                const float2 f = (p.res_dtype == LNS_F16) ? unpack2_h16<true>(ww[j]) : unpack2_h16<false>(ww[j]);
                v[h4 * 8 + 2 * j] += f.x;
                v[h4 * 8 + 2 * j + 1] += f.y;
              }
            }
          }
        } else if (!FAST && row_ok && p.residual) {
#pragma unroll
          for (int j = 0; j < 32; j += 4) {
            float4 t = ld4_as_float(p.residual, p.res_dtype, rrow + c0 + j);
            v[j] += t.x; v[j + 1] += t.y; v[j + 2] += t.z; v[j + 3] += t.w;
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
            const int ch = (cc >> 3) + h4;  // 16-byte chunk index inside the 128-byte staged row
            asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(my_row_st + (uint32_t)((ch ^ (lane & 7)) << 4)),
                         "r"(pk[0]), "r"(pk[1]), "r"(pk[2]), "r"(pk[3]) : "memory");
          }
        } else if (row_ok) {
          float4* dst = reinterpret_cast<float4*>(reinterpret_cast<float*>(p.y) + yrow + c0);
#pragma unroll
          for (int j = 0; j < 8; ++j) dst[j] = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
        }
      }
      if (y16 && tma) {
        // one TMA tensor store per staged tile: box = 64 channels x 8 pixels x 4 image rows, SWIZZLE_128B (the staged rows
        // are in TMEM-lane order = (image row, pixel) order, chunk-swizzled by row & 7 = the map's shared-memory layout);
        // partial tiles are clipped by the TMA unit
        hptx::fence_proxy_async();
        __syncwarp();
        const int yq = ty * kTileH + quad * 4, xq = tx * kTileW;
        if (lane == 0) {
          if (yq < g.Hout && xq < g.Wout) hptx::tma_store_4d(&tmap_y, my_stage, cg, xq, yq, b);
          hptx::bulk_commit();  // one group per staging, empty or not: wait_group.read N then counts stagings
        }
      } else if (y16) {
        __syncwarp();
        // staged row r = 4*pass + rd_row of this warp's quadrant = tile row quad*4 + pass/2, tile col 4*(pass&1) + rd_row
        const int yq = ty * kTileH + quad * 4, xq = tx * kTileW + rd_row;
        __nv_bfloat16* qbase = reinterpret_cast<__nv_bfloat16*>(p.y) + (int64_t)b * g.y_bstride +
                               ((int64_t)yq * g.Wout + xq) * g.Cout + cg + rd_chunk * 8;
        const int64_t row_stride = (int64_t)g.Wout * g.Cout;
        const bool interior = (ty + 1) * kTileH <= g.Hout && (tx + 1) * kTileW <= g.Wout;  // warp uniform
#pragma unroll
        for (int pass = 0; pass < 8; ++pass) {
          const int r = pass * 4 + rd_row;
          uint32_t w0, w1, w2, w3;
          asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(w0), "=r"(w1), "=r"(w2), "=r"(w3)
                       : "r"(my_stage + (uint32_t)r * 128u + (uint32_t)((rd_chunk ^ (r & 7)) << 4)));
          if (interior || (yq + (pass >> 1) < g.Hout && xq + 4 * (pass & 1) < g.Wout))
            *reinterpret_cast<uint4*>(qbase + (pass >> 1) * row_stride + (pass & 1) * 4 * g.Cout) = make_uint4(w0, w1, w2, w3);
        }
        __syncwarp();
      }
    }
    hptx::tc_fence_before();
    hptx::mbar_arrive(acc_empty(a));
    tx += eadv_x;
    if (tx >= p.tiles_x) { tx -= p.tiles_x; ++ty; }
    ty += eadv_y;
    if (ty >= p.tiles_y) { ty -= p.tiles_y; ++b; }
    b += eadv_b;
  }
  if (tma && lane == 0) hptx::bulk_wait_all();  // the staging buffers must outlive the stores that read them
}

// WS = 1: SPLIT filter (LNS_W_UMMA_F16X2: hi plane + lo plane, both resident) and two MMAs per (tap, k) -- A.Whi + A.Wlo --
// so the filter's 16-bit rounding error disappears from the layer (Cout = 64 only: 2 x 72 KB of filter = the Cout = 128 budget).
template <int NT, int kIssuers, int kAccs, int WS>
__global__ void __launch_bounds__(32 * (4 + kIssuers + 4), 1)
    conv_halo_kernel(const HaloParams p, const __grid_constant__ CUtensorMap tmap_y, const __grid_constant__ CUtensorMap tmap_r) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (hptx::smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_gen = smem_raw + (smem_base - hptx::smem_u32(smem_raw));
  constexpr uint32_t kWPlane = 9u * NT * 128u;
  constexpr uint32_t kWBytes = kWPlane * (1u + WS);
  const uint32_t w_base = smem_base;
  const uint32_t halo_base = smem_base + kWBytes;
  const uint32_t stage_out = halo_base + (uint32_t)p.stages * (uint32_t)p.halo_bytes;  // 4 warps x nbuf x 4 KB output staging
  const uint32_t stage_bytes = 4u * (uint32_t)p.nbuf * 4096u;
  const uint32_t res_stage0 = stage_out + stage_bytes;     // 4 warps x 4 KB residual tiles (only with tma_res)
  const uint32_t res_bytes = p.tma_res ? 4u * 4096u : 0u;
  // bias folded into the GEMM: one extra K=16 MMA per tile, A = a "ones" tile (8 rows, every row group aliases it through
  // SBO = 0) with 1.0 in k = 0, 1; B = [Cout][k] with bias split as bf16 hi (k = 0) + lo (k = 1): hi + lo is exact to 2^-17.
  const uint32_t bias_b = stage_out + stage_bytes + res_bytes;         // NT x 128 B, swizzled K-major like the filter
  const uint32_t ones_a = bias_b + (uint32_t)NT * 128u;    // 8 x 128 B
  const uint32_t bar_base = ones_a + 1024u;
  // barriers: w, halo_full[4], halo_empty[4], acc_full[8], acc_empty[8]; then the TMEM slot; then res_full[4 epilogue warps]
  const uint32_t w_bar = bar_base;
  auto halo_full = [&](int s) { return bar_base + 8u * (1 + s); };
  auto halo_empty = [&](int s) { return bar_base + 8u * (5 + s); };
  auto acc_full = [&](int a) { return bar_base + 8u * (9 + a); };
  auto acc_empty = [&](int a) { return bar_base + 8u * (17 + a); };  // up to 8 accumulators
  const uint32_t tmem_slot = bar_base + 8u * 25;
  auto res_full = [&](int w) { return bar_base + 8u * (26 + w); };
  volatile uint32_t* tmem_slot_gen =
      reinterpret_cast<volatile uint32_t*>(smem_gen + kWBytes + (size_t)p.stages * p.halo_bytes + stage_bytes + res_bytes + NT * 128 + 1024 + 8 * 25);

  const ConvGeom& g = p.g;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int HS = p.stages;

  if (tid == 0) {
    hptx::mbar_init(w_bar, 1);
    for (int s = 0; s < 4; ++s) {
      hptx::mbar_init(halo_full(s), 32);
      hptx::mbar_init(halo_empty(s), 1);
    }
    for (int a = 0; a < kAccs; ++a) {
      hptx::mbar_init(acc_full(a), 1);
      hptx::mbar_init(acc_empty(a), 128);
    }
    for (int w = 0; w < 4; ++w) hptx::mbar_init(res_full(w), 1);
    hptx::fence_mbar_init();
  }
  if (warp == 4) {
    hptx::tmem_alloc(tmem_slot, kAccs * NT);
    hptx::tmem_relinquish();
  }
  if (p.bias) {
    uint8_t* bb = smem_gen + kWBytes + (size_t)p.stages * p.halo_bytes + stage_bytes + res_bytes;
    for (int e = tid; e < NT * 8; e += blockDim.x) {  // 16-byte chunks of the bias B tile
      const int n = e >> 3, ch = e & 7;
      uint4 v = make_uint4(0, 0, 0, 0);
      if (ch == 0 && n < g.Cout) {
        const float bv = __ldg(p.bias + n);
        if (p.x_f16) {
          const uint16_t hi = to_h16<true>(bv), lo = to_h16<true>(bv - from_h16<true>(hi));
          v.x = (uint32_t)hi | ((uint32_t)lo << 16);
        } else {
          const uint16_t hi = to_h16<false>(bv), lo = to_h16<false>(bv - from_h16<false>(hi));
          v.x = (uint32_t)hi | ((uint32_t)lo << 16);
        }
      }
      *reinterpret_cast<uint4*>(bb + n * 128 + ((ch ^ (n & 7)) << 4)) = v;
    }
    for (int e = tid; e < 8 * 8; e += blockDim.x) {   // the ones tile
      const int r = e >> 3, ch = e & 7;
      uint4 v = make_uint4(0, 0, 0, 0);
      if (ch == 0) v.x = p.x_f16 ? 0x3C003C00u : 0x3F803F80u;  // (1.0, 1.0) in f16 / bf16
      *reinterpret_cast<uint4*>(bb + NT * 128 + r * 128 + ((ch ^ r) << 4)) = v;
    }
    hptx::fence_proxy_async();  // generic-proxy writes -> visible to the tensor core's async-proxy reads
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
      for (int tap = 0; tap < 9; ++tap) {
        hptx::bulk_g2s(w_base + (uint32_t)tap * NT * 128u, p.w + (int64_t)tap * g.Cout * 64, (uint32_t)NT * 128u, w_bar);
        if (WS)  // lo plane: 9 * Cout * 64 elements behind the hi plane
          hptx::bulk_g2s(w_base + kWPlane + (uint32_t)tap * NT * 128u, p.w + (int64_t)(9 + tap) * g.Cout * 64, (uint32_t)NT * 128u, w_bar);
      }
    }
    // The halo is a rectangle and the source index map is separable: offset(hy, hx) = rowterm(hy) + colterm(hx) (wrap / zero
    // padding / nearest resize; "outside" -> zero fill).
    // One producer WARP per tile: warp w (< HS) takes tiles it = w, w+HS, ... and therefore always fills ring stage w.
    // (One warp per STAGE matters: mbarrier parity waits are only race free when the waits on a barrier are issued in phase
    // order by one agent.)  The per-tile fixed work (tile decode, separable terms, barrier round trip) is paid once per tile
    // by one warp instead of by all four, and HS tiles are in flight independently.
    // Copy schedule (halo_fill<DIL> above): per-lane fixed halo columns, one shuffle per halo row, ~5 instructions per copy.
    const bool docopy = !(p.debug & 1);
    // incremental tile decode: tile = (b * tiles_y + ty) * tiles_x + tx advances by HS * gridDim.x per visit of this warp
    int tile = (warp < HS) ? blockIdx.x + warp * (int)gridDim.x : p.ntiles;  // warps >= HS have no stage: idle
    int tx = tile % p.tiles_x, t2 = tile / p.tiles_x;
    int ty = t2 % p.tiles_y, b = t2 / p.tiles_y;
    const int adv = HS * (int)gridDim.x;
    const int adv_x = adv % p.tiles_x, adv_t2 = adv / p.tiles_x;
    const int adv_y = adv_t2 % p.tiles_y, adv_b = adv_t2 / p.tiles_y;
    for (int it = warp; tile < p.ntiles; it += HS, tile += adv) {
      const int s = warp;  // == it % HS
      const uint32_t st = halo_base + (uint32_t)s * (uint32_t)p.halo_bytes;
      if (g.dil == 1) halo_fill<1>(p, tx, ty, b, lane, st, halo_empty(s), it >= HS, ((it / HS) & 1) ^ 1, docopy);
      else if (g.dil == 2) halo_fill<2>(p, tx, ty, b, lane, st, halo_empty(s), it >= HS, ((it / HS) & 1) ^ 1, docopy);
      else halo_fill<3>(p, tx, ty, b, lane, st, halo_empty(s), it >= HS, ((it / HS) & 1) ^ 1, docopy);
      hptx::cp_async_arrive_noinc(halo_full(s));
      // next tile of this warp
      tx += adv_x;
      if (tx >= p.tiles_x) { tx -= p.tiles_x; ++ty; }
      ty += adv_y;
      if (ty >= p.tiles_y) { ty -= p.tiles_y; ++b; }
      b += adv_b;
    }
  } else if (warp < 4 + kIssuers) {
    // ============================== MMA issuers ==============================
    // Two issue warps: warp 4 takes even tiles, warp 5 odd tiles (a single issuer's instruction stream -- ~10 SASS
    // instructions per 32-cycle MMA at N=64 -- was slower than the tensor pipe).  tcgen05.commit tracks the issuing
    // thread's own MMAs, and the two warps never share an accumulator, so no ordering between them is needed.
    // All 32 lanes run this loop; lane 0's tcgen05.mma / tcgen05.commit are the ones that execute.  Descriptors are
    // (hi, lo) 32-bit pairs: hi (SBO | version | swizzle mode) is loop invariant, lo is the 16-byte-unit start
    // address = stage base + a per-(tap, k) constant -> one add per operand per MMA instead of rebuilding 64-bit
    // descriptors in a single thread (which made the issue thread the bottleneck: 25% tensor-pipe in the first profile).
    {
      const uint32_t leader = (lane == 0) ? 1u : 0u;
      // kind::f16 instruction descriptor: D = f32, A/B format 1 = bf16 or 0 = f16, K-major, N, M = 128
      const uint32_t fmt_ab = p.x_f16 ? 0u : ((1u << 7) | (1u << 10));
      const uint32_t idesc = (1u << 4) | fmt_ab | (((uint32_t)g.Cout >> 3) << 17) | ((128u >> 4) << 24);
      const uint32_t desc_hi_a = (((uint32_t)p.HW * 128u) >> 4) | (1u << 14) | (2u << 29);
      const uint32_t desc_hi_b = (1024u >> 4) | (1u << 14) | (2u << 29);
      const uint32_t w_lo = (w_base & 0x3FFFFu) >> 4;
      const uint32_t row_lo = ((uint32_t)(g.dil * p.HW) * 128u) >> 4;  // one dilated halo row
      const uint32_t col_lo = ((uint32_t)g.dil * 128u) >> 4;           // one dilated halo column
      hptx::mbar_wait(w_bar, 0);
      const int me = warp - 4;
      int it = 0;
      for (int tile = blockIdx.x; tile < p.ntiles; tile += gridDim.x, ++it) {
        if ((it % kIssuers) != me) continue;
        const int s = it % HS, a = it % kAccs;
        if (it >= kAccs) hptx::mbar_wait(acc_empty(a), ((it / kAccs) & 1) ^ 1);
        hptx::mbar_wait(halo_full(s), (it / HS) & 1);
        hptx::fence_proxy_async();
        hptx::tc_fence_after();
        const uint32_t st_lo = ((halo_base + (uint32_t)s * (uint32_t)p.halo_bytes) & 0x3FFFFu) >> 4;
        const uint32_t d_tmem = tmem_acc + (uint32_t)(a * NT);
#pragma unroll
        for (int tap = 0; tap < 9; ++tap) {
          const int ky = tap / 3, kx = tap - ky * 3;
          const uint32_t a_lo = st_lo + (uint32_t)ky * row_lo + (uint32_t)kx * col_lo;
          const uint32_t b_lo = w_lo + (uint32_t)tap * ((uint32_t)NT * 128u >> 4);
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            hptx::umma_bf16_if((p.debug & 4) ? 0u : leader, d_tmem, a_lo + 2u * k, desc_hi_a, b_lo + 2u * k, desc_hi_b, idesc,
                               (tap | k) != 0 ? 1u : 0u);
            if (WS)
              hptx::umma_bf16_if((p.debug & 4) ? 0u : leader, d_tmem, a_lo + 2u * k, desc_hi_a, b_lo + (kWPlane >> 4) + 2u * k,
                                 desc_hi_b, idesc, 1u);
          }
        }
        if (p.bias)
          hptx::umma_bf16_if(leader, d_tmem, (ones_a & 0x3FFFFu) >> 4, (1u << 14) | (2u << 29) /* SBO = 0 */,
                             (bias_b & 0x3FFFFu) >> 4, desc_hi_b, idesc, 1u);
        hptx::umma_commit_if(leader, halo_empty(s));
        hptx::umma_commit_if(leader, acc_full(a));
      }
    }
    __syncwarp();
  } else {
    // ============================== epilogue (last 4 warps) ==============================
    const bool fast = p.tma_store && is_h16(p.y_dtype) && !p.sample_bias && !p.pre_add && (!p.residual || p.tma_res);
#define LNS_HALO_EPI(TR, FA) \
  halo_epilogue<NT, kIssuers, kAccs, TR, FA>(p, tmap_y, tmap_r, tmem_acc, stage_out, res_stage0, acc_full(0), acc_empty(0), res_full(0), warp, lane)
    if (p.tma_res) {
      if (fast) LNS_HALO_EPI(true, true);
      else LNS_HALO_EPI(true, false);
    } else {
      if (fast) LNS_HALO_EPI(false, true);
      else LNS_HALO_EPI(false, false);
    }
#undef LNS_HALO_EPI
  }

  hptx::tc_fence_before();
  __syncthreads();
  if (warp == 4) {
    hptx::tc_fence_after();
    hptx::tmem_dealloc(tmem_acc, kAccs * NT);
  }
}

bool conv_halo_supported(const LnsConvDesc* d) {
  return d->KH == 3 && d->KW == 3 && d->stride == 1 && d->Cin == 64 && (d->Cout == 64 || d->Cout == 128) &&
         d->pad_t == d->dil && d->pad_l == d->dil && d->Hout == d->Hv && d->Wout == d->Wv && d->dil >= 1 &&
         d->dil <= 3 && is_h16_host(d->x_dtype) && d->x_layout == LNS_NHWC && d->y_layout == LNS_NHWC &&
         d->pro_scale == nullptr && d->pro_act == LNS_ACT_NONE &&
         (d->w_format == (d->x_dtype == LNS_F16 ? LNS_W_UMMA_F16 : LNS_W_UMMA_BF16) ||
          (d->w_format == LNS_W_UMMA_F16X2 && d->x_dtype == LNS_F16 && d->Cout == 64)) &&
         d->dil <= d->Hv && d->dil <= d->Wv;
}

// cuTensorMapEncodeTiled through the runtime's driver entry point (the library does not link libcuda)
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn encode_tiled_fn() {
  static EncodeTiledFn fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void* f = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(f);
  }
  return fn;
}

// 4-D map of a 16-bit NHWC output [B][Hout][Wout][Cout] (batch stride in elements) whose box is one epilogue warp's staged
// tile: 64 channels x kTileW pixels x 4 image rows, SWIZZLE_128B
static bool make_y_tmap(CUtensorMap* tm, const LnsConvDesc* d, const void* base, int dtype, int64_t bstride) {
  EncodeTiledFn enc = encode_tiled_fn();
  if (!enc) return false;
  const cuuint64_t dims[4] = {(cuuint64_t)d->Cout, (cuuint64_t)d->Wout, (cuuint64_t)d->Hout, (cuuint64_t)d->B};
  const cuuint64_t strides[3] = {(cuuint64_t)d->Cout * 2ull, (cuuint64_t)d->Wout * d->Cout * 2ull, (cuuint64_t)bstride * 2ull};
  const cuuint32_t box[4] = {64u, (cuuint32_t)kTileW, 4u, 1u};
  const cuuint32_t estr[4] = {1u, 1u, 1u, 1u};
  const CUresult r = enc(tm, dtype == LNS_F16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(base), dims,
                         strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                         CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS;
}

template <int NT, int KI, int KA, int WS = 0>
static int launch_halo(const HaloParams& p, const CUtensorMap& tmap_y, const CUtensorMap& tmap_r, int smem_bytes, int grid,
                       cudaStream_t stream) {
  auto kern = conv_halo_kernel<NT, KI, KA, WS>;
  LNS_OPT_IN_SMEM(kern, 227 * 1024, "conv_halo");  // per (template instance, device)
  kern<<<grid, 32 * (4 + KI + 4), smem_bytes, stream>>>(p, tmap_y, tmap_r);
  return check_launch("conv_halo_kernel");
}

int conv2d_halo(const LnsConvDesc* d, cudaStream_t stream) {
  LNS_REQUIRE(conv_halo_supported(d),
              "lns_conv2d(halo): needs a same-size 3x3 stride-1 conv, Cin=64, Cout in {64,128}, pad=dil<=3, NHWC bf16/f16 "
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
  p.x_f16 = d->x_dtype == LNS_F16 ? 1 : 0;
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
  p.debug = 0;
#ifdef LNS_HALO_DEBUG_BUILD  // timing ablations (skip copies / epilogue / MMAs: WRONG results) exist only in a -DLNS_HALO_DEBUG_BUILD build
  {
    const char* dbg = getenv("LNS_HALO_DEBUG");
    p.debug = dbg ? atoi(dbg) : 0;
  }
#endif
  const int NT = d->Cout;
  const int wsplit = d->w_format == LNS_W_UMMA_F16X2 ? 1 : 0;
  // output staging: two 4 KB buffers per epilogue warp when four halo stages still fit next to them, else one
  p.nbuf = 2;
  int fixed = (1 + wsplit) * 9 * NT * 128 + 4 * p.nbuf * 4096 /*output staging*/ + NT * 128 + 1024 /*bias + ones tiles*/ + 256 + 1024;
  if ((227 * 1024 - fixed) / p.halo_bytes < 4) {
    p.nbuf = 1;
    fixed -= 4 * 4096;
  }
  int stages = (227 * 1024 - fixed) / p.halo_bytes;
  if (stages > 4) stages = 4;
  if (stages == 3) stages = 2;  // the issuer count must divide the ring depth (see below)
  LNS_REQUIRE(stages >= 2, "lns_conv2d(halo): shared memory too small for dilation %d with Cout %d", d->dil, d->Cout);
  p.stages = stages;
  const int sms = device_sm_count();
  int grid = p.ntiles < sms ? p.ntiles : sms;
  // 16-bit outputs leave through TMA tensor stores (LNS_HALO_TMA=0: the staged read-back + st.global path)
  CUtensorMap tmap_y;
  memset(&tmap_y, 0, sizeof(tmap_y));
  p.tma_store = 0;
  {
    const char* c = getenv("LNS_HALO_TMA");
    const bool want = c ? atoi(c) != 0 : true;
    if (want && is_h16_host(d->y_dtype) && d->Cout % 64 == 0 && ((int64_t)d->y_bstride * 2) % 16 == 0 &&
        make_y_tmap(&tmap_y, d, d->y, d->y_dtype, d->y_bstride))
      p.tma_store = 1;
  }
  // a 16-bit residual arrives through TMA tensor loads when its 16 KB of tiles fit next to the 4-stage ring (Cout = 64)
  CUtensorMap tmap_r;
  memset(&tmap_r, 0, sizeof(tmap_r));
  p.tma_res = 0;
  int smem = fixed + stages * p.halo_bytes;
  {
    const char* c = getenv("LNS_HALO_TMA_RES");
    const bool want = c ? atoi(c) != 0 : true;
    if (want && d->residual && is_h16_host(d->res_dtype) && d->Cout == 64 && smem + 4 * 4096 <= 227 * 1024 &&
        ((int64_t)d->res_bstride * 2) % 16 == 0 && make_y_tmap(&tmap_r, d, d->residual, d->res_dtype, d->res_bstride)) {
      p.tma_res = 1;
      smem += 4 * 4096;
    }
  }
  if (getenv("LNS_HALO_VERBOSE"))
    fprintf(stderr,
            "conv_halo: B=%d %dx%d->%dx%d Cout=%d dil=%d resize=%d tma_store=%d nbuf=%d stages=%d grid=%d x=%p xbs=%lld y=%p ybs=%lld "
            "act=%d bias=%d sbias=%d pre=%d res=%d tma_res=%d xdt=%d ydt=%d circ=%d%d\n",
            d->B, d->Hin, d->Win, d->Hout, d->Wout, d->Cout, d->dil, p.resize, p.tma_store, p.nbuf, stages, grid, d->x,
            (long long)d->x_bstride, d->y, (long long)d->y_bstride, d->act, d->bias != nullptr, d->sample_bias != nullptr,
            d->pre_add != nullptr, d->residual != nullptr, p.tma_res, d->x_dtype, d->y_dtype, p.g.circ_h, p.g.circ_w);
  // MMA issue warps.  Tile `it` uses ring stage it % stages and accumulator it % 4 and is issued by warp it % kIssuers: every
  // mbarrier must be waited on by ONE agent in phase order (a parity wait issued a whole phase early passes immediately), so
  // the issuer count has to divide both the ring depth and the accumulator count.  Measured on B200 at 64->64 @ 64x64 (4
  // stages): 1 issuer 0.64 ms, 2 issuers 1053 TFLOP/s, 4 issuers 1101 -- the instruction stream of one issuing warp (~100
  // cycles per tcgen05.mma with its descriptor arithmetic) limits the tensor pipe at N = 64.  (3 issuers measured 1137-1148 but
  // break the one-agent rule: with a 2-stage ring that configuration dead-locked.)  LNS_HALO_ISSUERS=1|2|4 overrides.
  LNS_REQUIRE(p.stages == 2 || p.stages == 4, "lns_conv2d(halo): internal: ring depth %d", p.stages);
  int issuers = p.stages == 4 ? 4 : 2;
  {
    const char* c = getenv("LNS_HALO_ISSUERS");
    const int want = c ? atoi(c) : 0;
    if ((want == 1 || want == 2 || want == 4) && p.stages % want == 0) issuers = want;
  }
  if (wsplit) {
    if (issuers == 4) return launch_halo<64, 4, 4, 1>(p, tmap_y, tmap_r, smem, grid, stream);
    if (issuers == 2) return launch_halo<64, 2, 4, 1>(p, tmap_y, tmap_r, smem, grid, stream);
    return launch_halo<64, 1, 4, 1>(p, tmap_y, tmap_r, smem, grid, stream);
  }
  if (NT == 64) {
    if (issuers == 4) return launch_halo<64, 4, 4>(p, tmap_y, tmap_r, smem, grid, stream);
    if (issuers == 2) return launch_halo<64, 2, 4>(p, tmap_y, tmap_r, smem, grid, stream);
    return launch_halo<64, 1, 4>(p, tmap_y, tmap_r, smem, grid, stream);
  }
  if (issuers == 4) return launch_halo<128, 4, 4>(p, tmap_y, tmap_r, smem, grid, stream);
  if (issuers == 2) return launch_halo<128, 2, 4>(p, tmap_y, tmap_r, smem, grid, stream);
  return launch_halo<128, 1, 4>(p, tmap_y, tmap_r, smem, grid, stream);
}

}  // namespace lns
