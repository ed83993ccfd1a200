// Fused core of FABlock2D (modules/factorized_attention.py:144-159) for the bf16 path.
//
// Unfused, the 8x channel-expanded tensor u_phi [B,H,W,512] makes five HBM round trips (in_proj write, two axial
// contractions read+write, InstanceNorm statistics read, normalise read+write): ~9 MB per 32x32 sample, half of the
// whole rollout's time in the round-1 timeline.  Here ONE CTA per (sample, head) keeps that head's slice
// u_phi_h [H*W][64] (bf16, <= 166 KB) in shared memory from the in_proj GEMM to the normalised output:
//   A  u_phi_h = u (raw, bf16) x (W_in_proj[h] * gn_scale[b]) + W_in_proj[h] . gn_shift[b]      (GroupNorm(1) folded
//      into a per-sample filter and bias, so the raw activation feeds the tensor cores directly)
//   B  contraction over H per image column   u[i,m,:] = sum_j Kx[h][i][j] u[j,m,:]   (in place, column private)
//   C  contraction over W per image row      u[i,l,:] = sum_m Ky[h][l][m] u[i,m,:]   (in place, row private)
//   D  InstanceNorm2d statistics per channel over the H*W pixels (this head's 64 channels are CTA local)
//   E  normalise, write bf16 [B,H,W,512] (channels h*64..h*64+63) -- the only HBM write; to_out's 1x1 convs follow.
// All GEMMs are mma.sync.m16n8k16 (bf16 x bf16 -> fp32) fed by ldmatrix; HBM traffic per sample drops from ~9 MB to
// 128 KB in + 1 MB out.
#include "common.cuh"

namespace lns {

namespace {
constexpr int kUS = 72;  // bf16 elements per shared-memory row (64 + 8 pad -> conflict-free ldmatrix)

__device__ __forceinline__ void ldsm_x4(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr));
}
__device__ __forceinline__ void ldsm_x2(uint32_t addr, uint32_t& r0, uint32_t& r1) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x2.shared.b16 {%0,%1}, [%2];" : "=r"(r0), "=r"(r1) : "r"(addr));
}
__device__ __forceinline__ void ldsm_x2_trans(uint32_t addr, uint32_t& r0, uint32_t& r1) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x2.trans.shared.b16 {%0,%1}, [%2];" : "=r"(r0), "=r"(r1) : "r"(addr));
}
// F16 = false: bf16 operands / storage, true: IEEE half
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
__device__ __forceinline__ uint32_t s32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// In-place axial contraction of one line (an image column or row) held in U_s:
//   out[i][c] = sum_j Kmat[i][j] * line[j][c],  line element j lives at U_s + base + j*step  (offsets in bf16 elements).
// KT = ceil(n/16) k/m tiles.  The whole line (B fragments) is read into registers before anything is written.
template <int KT, bool STATS, bool F16>
__device__ __forceinline__ void contract_line(__nv_bfloat16* U_s, int base, int step, int n, const __nv_bfloat16* K_s, int kstride,
                                              int lane, float* st) {
  uint32_t bf[KT][8][2];
#pragma unroll
  for (int kt = 0; kt < KT; ++kt) {
    int j = kt * 16 + (lane & 15);
    j = j < n ? j : n - 1;  // padded rows: multiplied by the zero-padded kernel columns
    const __nv_bfloat16* rowp = U_s + (size_t)base + (size_t)j * step;
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) ldsm_x2_trans(s32(rowp + nt * 8), bf[kt][nt][0], bf[kt][nt][1]);
  }
  __syncwarp();
  const int g = lane >> 2, t = lane & 3;
#pragma unroll
  for (int mt = 0; mt < KT; ++mt) {
    float acc[8][4];
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) acc[nt][0] = acc[nt][1] = acc[nt][2] = acc[nt][3] = 0.f;
#pragma unroll
    for (int kt = 0; kt < KT; ++kt) {
      uint32_t a[4];
      ldsm_x4(s32(K_s + (size_t)(mt * 16 + (lane & 15)) * kstride + kt * 16 + (lane >> 4) * 8), a[0], a[1], a[2], a[3]);
#pragma unroll
      for (int nt = 0; nt < 8; ++nt) mma16816<F16>(acc[nt], a, bf[kt][nt][0], bf[kt][nt][1]);
    }
    const int i0 = mt * 16 + g, i1 = i0 + 8;
    const bool v0 = i0 < n, v1 = i1 < n;
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
      if (v0)
        *reinterpret_cast<uint32_t*>(U_s + (size_t)base + (size_t)i0 * step + nt * 8 + t * 2) = pack2_h16<F16>(acc[nt][0], acc[nt][1]);
      if (v1)
        *reinterpret_cast<uint32_t*>(U_s + (size_t)base + (size_t)i1 * step + nt * 8 + t * 2) = pack2_h16<F16>(acc[nt][2], acc[nt][3]);
      if (STATS) {  // InstanceNorm sums of channels nt*8 + 2t, +1 (fp32 values, before the bf16 rounding of the store)
        const float a0 = v0 ? acc[nt][0] : 0.f, a1 = v0 ? acc[nt][1] : 0.f, a2 = v1 ? acc[nt][2] : 0.f, a3 = v1 ? acc[nt][3] : 0.f;
        st[nt * 4 + 0] += a0 + a2;
        st[nt * 4 + 1] += a1 + a3;
        st[nt * 4 + 2] = fmaf(a0, a0, fmaf(a2, a2, st[nt * 4 + 2]));
        st[nt * 4 + 3] = fmaf(a1, a1, fmaf(a3, a3, st[nt * 4 + 3]));
      }
    }
  }
  __syncwarp();
}

template <int KT, bool STATS, bool F16>
__device__ __forceinline__ void contract_axis(__nv_bfloat16* U_s, int lines, int line_mul, int step, int n, const __nv_bfloat16* K_s,
                                              int kstride, int warp, int nwarp, int lane, float* st) {
  for (int l = warp; l < lines; l += nwarp) contract_line<KT, STATS, F16>(U_s, l * line_mul, step, n, K_s, kstride, lane, st);
}
}  // namespace

// grid (heads, B), block 512 (256 when an axis needs 3 k-tiles: register budget); instruction-issue bound, so 4 warps per
// scheduler instead of 2 buy latency hiding
template <int NTHR, bool F16>
__global__ void __launch_bounds__(NTHR, 1) fablock_core_kernel(const __nv_bfloat16* __restrict__ u, int H, int W, int heads,
                                                               const float* __restrict__ gn_scale, const float* __restrict__ gn_shift,
                                                               const float* __restrict__ w_in, const float* __restrict__ Kx,
                                                               const float* __restrict__ Ky, float eps,
                                                               __nv_bfloat16* __restrict__ out) {
  extern __shared__ __align__(16) uint8_t smraw[];
  const int HW = H * W;
  const int H16 = (H + 15) & ~15, W16 = (W + 15) & ~15;
  const int kxs = H16 + 8, kys = W16 + 8;
  // u_phi_h: pixel (y, x) at y*RS + x*72 elements; RS = W*72 + 8 -- the extra 16 bytes per image row rotate the banks so that
  // the 8 rows an ldmatrix / fragment store touches along a COLUMN (stride RS) fall into 8 different 16-byte bank groups
  // (without it the column stride W*144 B is a multiple of 128 B: 8-way conflicts on every access of phase B)
  const int RS = W * kUS + 8;
  __nv_bfloat16* U_s = reinterpret_cast<__nv_bfloat16*>(smraw);          // [H][RS]  (raw input first, u_phi_h in place)
  __nv_bfloat16* Ws_s = U_s + (size_t)H * RS;                            // [64 n][72 k]
  __nv_bfloat16* Kx_s = Ws_s + 64 * kUS;                                 // [H16][H16+8]
  __nv_bfloat16* Ky_s = Kx_s + (size_t)H16 * kxs;                        // [W16][W16+8]
  float* bias_s = reinterpret_cast<float*>(Ky_s + (size_t)W16 * kys);    // [64]
  float* red_s = bias_s + 64;                                            // [<=64 pixel groups][64][2]
  float* stat_s = red_s + 64 * 64 * 2;                                   // [64][2] mean, rstd
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int nthr = blockDim.x, nwarp = nthr >> 5, ppi = nthr >> 3;  // ppi: pixels covered per iteration at 8 lanes / pixel
  const int h = blockIdx.x, b = blockIdx.y;
  const int C = heads * 64;
  const __nv_bfloat16* ub = u + (int64_t)b * HW * 64;

  // ---- phase A loads first: the raw input streams in (cp.async, four commit groups = four quarters of the image) while the
  // setup below runs.  The 1x1 in_proj is pointwise in space, so the rows are copied straight into the rows of U_s they will
  // be replaced in: no staging buffer, no block-wide barrier per tile.
  const int nblk = (HW + 15) >> 4;                 // 16-row blocks
  const int blk_per_q = (nblk + 3) >> 2;
  const float invW = 1.0f / (float)W;
  {
    const int ch = tid & 7;
    int px = tid >> 3;                             // 8 lanes x 16 B per pixel row
    int y = __float2int_rd(((float)px + 0.5f) * invW), x = px - y * W;
    const int sy = ppi / W, sx = ppi - sy * W;
    for (int q = 0; q < 4; ++q) {
      const int px_end = min(HW, (q + 1) * blk_per_q * 16);
      for (; px < px_end; px += ppi) {
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(s32(U_s + (size_t)y * RS + (size_t)x * kUS + ch * 8)),
                     "l"(ub + (int64_t)px * 64 + ch * 8) : "memory");
        y += sy;
        x += sx;
        if (x >= W) {
          x -= W;
          ++y;
        }
      }
      asm volatile("cp.async.commit_group;" ::: "memory");
    }
  }

  // ---- setup: per-sample filter (GroupNorm scale folded in), bias (GroupNorm shift folded in), kernel matrices ----
  for (int e = tid; e < 64 * 8; e += nthr) {       // (output channel n, 8-wide k chunk): 16-byte shared stores
    const int n = e >> 3, kc = e & 7;
    const float4 w0 = __ldg(reinterpret_cast<const float4*>(w_in + (int64_t)(h * 64 + n) * 64 + kc * 8));
    const float4 w1 = __ldg(reinterpret_cast<const float4*>(w_in + (int64_t)(h * 64 + n) * 64 + kc * 8 + 4));
    const float4 s0 = __ldg(reinterpret_cast<const float4*>(gn_scale + (int64_t)b * 64 + kc * 8));
    const float4 s1 = __ldg(reinterpret_cast<const float4*>(gn_scale + (int64_t)b * 64 + kc * 8 + 4));
    *reinterpret_cast<uint4*>(Ws_s + n * kUS + kc * 8) =
        make_uint4(pack2_h16<F16>(w0.x * s0.x, w0.y * s0.y), pack2_h16<F16>(w0.z * s0.z, w0.w * s0.w),
                   pack2_h16<F16>(w1.x * s1.x, w1.y * s1.y), pack2_h16<F16>(w1.z * s1.z, w1.w * s1.w));
  }
  {
    // bias[n] = sum_k W[n][k] * shift[k]: 4 threads per output channel, 16 k each, fixed-order quad reduction
    const int n = (tid >> 2) & 63, part = tid & 3;  // (threads >= 256 recompute the same values; only tid < 256 store)
    float a = 0.f;
#pragma unroll
    for (int q4 = 0; q4 < 4; ++q4) {
      float4 wv = __ldg(reinterpret_cast<const float4*>(w_in + (int64_t)(h * 64 + n) * 64 + part * 16 + q4 * 4));
      float4 sv = __ldg(reinterpret_cast<const float4*>(gn_shift + (int64_t)b * 64 + part * 16 + q4 * 4));
      a = fmaf(wv.x, sv.x, a); a = fmaf(wv.y, sv.y, a); a = fmaf(wv.z, sv.z, a); a = fmaf(wv.w, sv.w, a);
    }
    a += __shfl_xor_sync(0xffffffffu, a, 1);
    a += __shfl_xor_sync(0xffffffffu, a, 2);
    if (part == 0 && tid < 256) bias_s[n] = a;
  }
  {
    const float* kg = Kx + ((int64_t)b * heads + h) * H * H;
    for (int i = warp; i < H16; i += nwarp)
      for (int j = lane; j < H16; j += 32)
        reinterpret_cast<uint16_t*>(Kx_s)[i * kxs + j] = to_h16<F16>((i < H && j < H) ? __ldg(kg + i * H + j) : 0.f);
    kg = Ky + ((int64_t)b * heads + h) * W * W;
    for (int i = warp; i < W16; i += nwarp)
      for (int j = lane; j < W16; j += 32)
        reinterpret_cast<uint16_t*>(Ky_s)[i * kys + j] = to_h16<F16>((i < W && j < W) ? __ldg(kg + i * W + j) : 0.f);
  }

  // ---- phase A: u_phi_h = u x Ws^T + bias, in place, 16-row blocks per warp ----
  const int g = lane >> 2, t = lane & 3;
  uint32_t wf[4][8][2];  // B fragments of the per-sample filter: identical for every 16-row block -> loaded once
  for (int q = 0; q < 4; ++q) {
    if (q == 0) asm volatile("cp.async.wait_group 3;" ::: "memory");
    else if (q == 1) asm volatile("cp.async.wait_group 2;" ::: "memory");
    else if (q == 2) asm volatile("cp.async.wait_group 1;" ::: "memory");
    else asm volatile("cp.async.wait_group 0;" ::: "memory");
    __syncthreads();  // quarter q has landed for every thread (q == 0 also publishes the setup writes)
    if (q == 0) {
#pragma unroll
      for (int ks = 0; ks < 4; ++ks)
#pragma unroll
        for (int nt = 0; nt < 8; ++nt)
          ldsm_x2(s32(Ws_s + (size_t)(nt * 8 + (lane & 7)) * kUS + ks * 16 + ((lane >> 3) & 1) * 8), wf[ks][nt][0], wf[ks][nt][1]);
    }
    const int blk_end = min(nblk, (q + 1) * blk_per_q);
    for (int blk = q * blk_per_q + warp; blk < blk_end; blk += nwarp) {
      // rows of this block: pixels blk*16 .. blk*16+15 (rows >= HW do not exist: clamp reads, skip writes)
      int pr = blk * 16 + (lane & 15);
      pr = pr < HW ? pr : HW - 1;
      const int yr = __float2int_rd(((float)pr + 0.5f) * invW);  // exact: pr < 2^20
      const __nv_bfloat16* arow = U_s + (size_t)yr * RS + (size_t)(pr - yr * W) * kUS;
      uint32_t a[4][4];
#pragma unroll
      for (int ks = 0; ks < 4; ++ks) ldsm_x4(s32(arow + ks * 16 + (lane >> 4) * 8), a[ks][0], a[ks][1], a[ks][2], a[ks][3]);
      float acc[8][4];
#pragma unroll
      for (int nt = 0; nt < 8; ++nt) acc[nt][0] = acc[nt][1] = acc[nt][2] = acc[nt][3] = 0.f;
#pragma unroll
      for (int ks = 0; ks < 4; ++ks) {
#pragma unroll
        for (int nt = 0; nt < 8; ++nt) mma16816<F16>(acc[nt], a[ks], wf[ks][nt][0], wf[ks][nt][1]);
      }
      __syncwarp();  // every lane's ldmatrix of the raw rows is done before they are overwritten
      const int r0 = blk * 16 + g, r1 = r0 + 8;
      const int y0 = __float2int_rd(((float)r0 + 0.5f) * invW), y1 = __float2int_rd(((float)r1 + 0.5f) * invW);
      __nv_bfloat16* o0 = U_s + (size_t)y0 * RS + (size_t)(r0 - y0 * W) * kUS + t * 2;
      __nv_bfloat16* o1 = U_s + (size_t)y1 * RS + (size_t)(r1 - y1 * W) * kUS + t * 2;
#pragma unroll
      for (int nt = 0; nt < 8; ++nt) {
        const float2 bs = *reinterpret_cast<const float2*>(bias_s + nt * 8 + t * 2);
        if (r0 < HW) *reinterpret_cast<uint32_t*>(o0 + nt * 8) = pack2_h16<F16>(acc[nt][0] + bs.x, acc[nt][1] + bs.y);
        if (r1 < HW) *reinterpret_cast<uint32_t*>(o1 + nt * 8) = pack2_h16<F16>(acc[nt][2] + bs.x, acc[nt][3] + bs.y);
      }
    }
  }
  __syncthreads();

  // ---- phase B: contraction over H, one image column per warp at a time (element j of column m at m*72 + j*RS) ----
  float st[32];  // phase C: per-thread InstanceNorm partial sums: [nt][sum(2t), sum(2t+1), sumsq(2t), sumsq(2t+1)]
#pragma unroll
  for (int i = 0; i < 32; ++i) st[i] = 0.f;
  if (H16 == 16) contract_axis<1, false, F16>(U_s, W, kUS, RS, H, Kx_s, kxs, warp, nwarp, lane, st);
  else if (H16 == 32) contract_axis<2, false, F16>(U_s, W, kUS, RS, H, Kx_s, kxs, warp, nwarp, lane, st);
  else contract_axis<3, false, F16>(U_s, W, kUS, RS, H, Kx_s, kxs, warp, nwarp, lane, st);
  __syncthreads();
  // ---- phase C: contraction over W, one image row per warp (element m of row i at i*RS + m*72); its epilogue also
  // accumulates the InstanceNorm sums of the 16 channels a thread holds (no separate statistics pass over the tile) ----
  if (W16 == 16) contract_axis<1, true, F16>(U_s, H, RS, kUS, W, Ky_s, kys, warp, nwarp, lane, st);
  else if (W16 == 32) contract_axis<2, true, F16>(U_s, H, RS, kUS, W, Ky_s, kys, warp, nwarp, lane, st);
  else contract_axis<3, true, F16>(U_s, H, RS, kUS, W, Ky_s, kys, warp, nwarp, lane, st);

  // ---- phase D: reduce the partial sums: over the 8 lanes that share t (shuffles), then over the warps (fixed order) ----
#pragma unroll
  for (int i = 0; i < 32; ++i) {
    st[i] += __shfl_xor_sync(0xffffffffu, st[i], 4);
    st[i] += __shfl_xor_sync(0xffffffffu, st[i], 8);
    st[i] += __shfl_xor_sync(0xffffffffu, st[i], 16);
  }
  if (lane < 4) {  // g == 0: channels nt*8 + 2*lane, +1
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
      const int c = nt * 8 + lane * 2;
      red_s[(warp * 64 + c) * 2 + 0] = st[nt * 4 + 0];
      red_s[(warp * 64 + c + 1) * 2 + 0] = st[nt * 4 + 1];
      red_s[(warp * 64 + c) * 2 + 1] = st[nt * 4 + 2];
      red_s[(warp * 64 + c + 1) * 2 + 1] = st[nt * 4 + 3];
    }
  }
  __syncthreads();
  if (tid < 64) {
    double sm = 0.0, ss = 0.0;
    for (int w = 0; w < nwarp; ++w) {
      sm += (double)red_s[(w * 64 + tid) * 2 + 0];
      ss += (double)red_s[(w * 64 + tid) * 2 + 1];
    }
    const double mean = sm / (double)HW;
    double var = ss / (double)HW - mean * mean;
    if (var < 0.0) var = 0.0;
    const double rstd = 1.0 / sqrt(var + (double)eps);
    stat_s[tid * 2 + 0] = (float)rstd;                 // scale
    stat_s[tid * 2 + 1] = (float)(-mean * rstd);       // shift
  }
  __syncthreads();

  // ---- phase E: normalise and write channels [h*64, h*64+64) of out [B][HW][C]; 8 lanes x 16 B per pixel ----
  {
    const int ch = tid & 7;
    float na[8], nb[8];  // this thread always handles the same 8 channels: y = x * na + nb
#pragma unroll
    for (int j2 = 0; j2 < 8; ++j2) {
      na[j2] = stat_s[(ch * 8 + j2) * 2];
      nb[j2] = stat_s[(ch * 8 + j2) * 2 + 1];
    }
    int px = tid >> 3;
    int py = __float2int_rd(((float)px + 0.5f) * invW), pxx = px - py * W;
    const int stepy = ppi / W, stepx = ppi - stepy * W;
    const __nv_bfloat16* sp = U_s + (size_t)py * RS + (size_t)pxx * kUS + ch * 8;
    const int sadv = stepy * RS + stepx * kUS, swrap = RS - W * kUS;  // element advance per iteration / extra on row wrap
    __nv_bfloat16* gp = out + ((int64_t)b * HW + px) * C + h * 64 + ch * 8;
    const int64_t gadv = (int64_t)ppi * C;
    for (; px < HW; px += ppi) {
      const uint4 raw = *reinterpret_cast<const uint4*>(sp);
      uint32_t w[4] = {raw.x, raw.y, raw.z, raw.w};
      uint32_t o[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float2 f = unpack2_h16<F16>(w[j]);
        o[j] = pack2_h16<F16>(fmaf(f.x, na[2 * j], nb[2 * j]), fmaf(f.y, na[2 * j + 1], nb[2 * j + 1]));
      }
      *reinterpret_cast<uint4*>(gp) = make_uint4(o[0], o[1], o[2], o[3]);
      gp += gadv;
      sp += sadv;
      pxx += stepx;
      if (pxx >= W) {
        pxx -= W;
        sp += swrap;
      }
    }
  }
}

// ---- FABlock2D pre-pass: ONE read of the block input gives (1) the GroupNorm(1, C) affine and (2) the two pooled
// tensors the low-rank kernels are built from.  Means commute with the per-channel affine, so
//   mean_W(GN(u))[y][c] = scale[c] * mean_W(u)[y][c] + shift[c]   (same for mean_H): no normalised copy of u is needed.
// grid B, block 256: thread = (channel quad q, pixel lane); one image row per iteration (row sums reduced through shared
// memory in a fixed order, column sums and the totals stay in registers).  Deterministic, batch independent.
constexpr int kPreMaxX = 8;
__global__ void __launch_bounds__(256) fablock_prepass_kernel(const void* __restrict__ u, int dtype, int H, int W, int C, int64_t bstride,
                                                               float eps, const float* __restrict__ gamma, const float* __restrict__ beta,
                                                               float* __restrict__ scale, float* __restrict__ shift,
                                                               float* __restrict__ pooled_x, float* __restrict__ pooled_y) {
  extern __shared__ float smf[];
  const int cg = C >> 2, rows = 256 / cg;
  float* red = smf;                              // [rows][C] row-sum partials of the current image row
  float* rowsum = red + (size_t)rows * C;        // [H][C]
  float* colsum = rowsum + (size_t)H * C;        // [W][C]
  float* tot = colsum + (size_t)W * C;           // [rows][C][2] -> later [C][2]
  float* ab = tot + (size_t)rows * C * 2;        // [C][2] scale, shift
  const int b = blockIdx.x;
  const int q = threadIdx.x % cg, lane = threadIdx.x / cg;
  float cs[kPreMaxX][4];
#pragma unroll
  for (int i = 0; i < kPreMaxX; ++i) cs[i][0] = cs[i][1] = cs[i][2] = cs[i][3] = 0.f;
  float s[4] = {0, 0, 0, 0}, ss[4] = {0, 0, 0, 0};
  for (int y = 0; y < H; ++y) {
    float rs[4] = {0, 0, 0, 0};
    if (lane < rows) {
#pragma unroll
      for (int i = 0; i < kPreMaxX; ++i) {
        const int x = lane + i * rows;
        if (x < W) {
          float4 v = ld4_as_float(u, dtype, (int64_t)b * bstride + ((int64_t)y * W + x) * C + q * 4);
          rs[0] += v.x; rs[1] += v.y; rs[2] += v.z; rs[3] += v.w;
          cs[i][0] += v.x; cs[i][1] += v.y; cs[i][2] += v.z; cs[i][3] += v.w;
          ss[0] = fmaf(v.x, v.x, ss[0]); ss[1] = fmaf(v.y, v.y, ss[1]);
          ss[2] = fmaf(v.z, v.z, ss[2]); ss[3] = fmaf(v.w, v.w, ss[3]);
        }
      }
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        s[j] += rs[j];
        red[lane * C + q * 4 + j] = rs[j];
      }
    }
    __syncthreads();
    for (int c = threadIdx.x; c < C; c += 256) {
      float a = 0.f;
      for (int l = 0; l < rows; ++l) a += red[l * C + c];
      rowsum[y * C + c] = a;
    }
    __syncthreads();
  }
  if (lane < rows) {
#pragma unroll
    for (int i = 0; i < kPreMaxX; ++i) {
      const int x = lane + i * rows;
      if (x < W) {
#pragma unroll
        for (int j = 0; j < 4; ++j) colsum[x * C + q * 4 + j] = cs[i][j];
      }
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      tot[((lane * C) + q * 4 + j) * 2 + 0] = s[j];
      tot[((lane * C) + q * 4 + j) * 2 + 1] = ss[j];
    }
  }
  __syncthreads();
  // GroupNorm(1, C): one group over all channels; warp 0 reduces (fixed order)
  if (threadIdx.x < 32) {
    double sum = 0.0, sumsq = 0.0;
    for (int c = threadIdx.x; c < C; c += 32) {
      double a = 0.0, a2 = 0.0;
      for (int l = 0; l < rows; ++l) {
        a += (double)tot[((l * C) + c) * 2 + 0];
        a2 += (double)tot[((l * C) + c) * 2 + 1];
      }
      sum += (double)(float)a;     // per-channel sums rounded to fp32 like the generic statistics kernels
      sumsq += (double)(float)a2;
    }
    sum = warp_sum_d(sum);
    sumsq = warp_sum_d(sumsq);
    const double n = (double)C * (double)H * (double)W;
    const double mean = sum / n;
    double var = sumsq / n - mean * mean;
    if (var < 0.0) var = 0.0;
    const double rstd = 1.0 / sqrt(var + (double)eps);
    for (int c = threadIdx.x; c < C; c += 32) {
      const double ga = gamma ? (double)gamma[c] : 1.0, be = beta ? (double)beta[c] : 0.0;
      const float sc = (float)(rstd * ga), sh = (float)(be - mean * rstd * ga);
      ab[c * 2 + 0] = sc;
      ab[c * 2 + 1] = sh;
      scale[(int64_t)b * C + c] = sc;
      shift[(int64_t)b * C + c] = sh;
    }
  }
  __syncthreads();
  const float invW = 1.f / (float)W, invH = 1.f / (float)H;
  for (int e = threadIdx.x; e < H * C; e += 256) {
    const int c = e % C;
    pooled_x[(int64_t)b * H * C + e] = fmaf(rowsum[e] * invW, ab[c * 2], ab[c * 2 + 1]);
  }
  for (int e = threadIdx.x; e < W * C; e += 256) {
    const int c = e % C;
    pooled_y[(int64_t)b * W * C + e] = fmaf(colsum[e] * invH, ab[c * 2], ab[c * 2 + 1]);
  }
}

// v2 of the pre-pass: the first version walked the image rows one after the other with two block barriers per row (latency
// bound: 440 us for 4096 samples of 32x32x64 = 1.2 TB/s).  Here every warp owns whole image rows (y = warp, warp + 8, ...),
// lane = (pixel lane, channel quad), all of a row's loads are in flight together, row sums are finished with shuffles and the
// column sums / totals stay in registers until one fixed-order cross-warp reduction at the end: two barriers per SAMPLE.
// Needs C <= 128 (a warp covers all channel quads of at least one pixel) and W <= NX * (128 / C).
template <int NX>
__global__ void __launch_bounds__(256) fablock_prepass2_kernel(const void* __restrict__ u, int dtype, int H, int W, int C, int64_t bstride,
                                                                float eps, const float* __restrict__ gamma, const float* __restrict__ beta,
                                                                float* __restrict__ scale, float* __restrict__ shift,
                                                                float* __restrict__ pooled_x, float* __restrict__ pooled_y,
                                                                uint16_t* __restrict__ staged, int lgw) {
  extern __shared__ float smf[];
  float* rowsum = smf;                           // [H][C]
  float* colpart = rowsum + (size_t)H * C;       // [8 warps][W][C]
  float* tot = colpart + (size_t)8 * W * C;      // [8 warps][C][2]
  float* ab = tot + (size_t)8 * C * 2;           // [C][2] scale, shift
  const int b = blockIdx.x;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int cg = C >> 2, npl = 32 / cg;          // channel quads, pixel lanes per warp
  const int q = lane % cg, pl = lane / cg;
  float cs[NX][4];
#pragma unroll
  for (int i = 0; i < NX; ++i) cs[i][0] = cs[i][1] = cs[i][2] = cs[i][3] = 0.f;
  float s[4] = {0, 0, 0, 0}, ss[4] = {0, 0, 0, 0};
  for (int y = warp; y < H; y += 8) {
    float4 v[NX];
    {
      int64_t off[NX];
      bool ok[NX];
#pragma unroll
      for (int i = 0; i < NX; ++i) {
        const int x = pl + i * npl;
        ok[i] = x < W;
        off[i] = (int64_t)b * bstride + ((int64_t)y * W + (ok[i] ? x : W - 1)) * C + q * 4;
      }
      ld4n_as_float<NX>(u, dtype, off, ok, v);
    }
    float rs[4] = {0, 0, 0, 0};
#pragma unroll
    for (int i = 0; i < NX; ++i) {
      rs[0] += v[i].x; rs[1] += v[i].y; rs[2] += v[i].z; rs[3] += v[i].w;
      cs[i][0] += v[i].x; cs[i][1] += v[i].y; cs[i][2] += v[i].z; cs[i][3] += v[i].w;
      ss[0] = fmaf(v[i].x, v[i].x, ss[0]); ss[1] = fmaf(v[i].y, v[i].y, ss[1]);
      ss[2] = fmaf(v[i].z, v[i].z, ss[2]); ss[3] = fmaf(v[i].w, v[i].w, ss[3]);
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      s[j] += rs[j];
      for (int o = cg; o < 32; o <<= 1) rs[j] += __shfl_xor_sync(0xffffffffu, rs[j], o);  // over the pixel lanes
    }
    if (pl == 0) *reinterpret_cast<float4*>(rowsum + (size_t)y * C + q * 4) = make_float4(rs[0], rs[1], rs[2], rs[3]);
  }
#pragma unroll
  for (int i = 0; i < NX; ++i) {
    const int x = pl + i * npl;
    if (x < W) *reinterpret_cast<float4*>(colpart + ((size_t)warp * W + x) * C + q * 4) = make_float4(cs[i][0], cs[i][1], cs[i][2], cs[i][3]);
  }
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    for (int o = cg; o < 32; o <<= 1) {
      s[j] += __shfl_xor_sync(0xffffffffu, s[j], o);
      ss[j] += __shfl_xor_sync(0xffffffffu, ss[j], o);
    }
    if (pl == 0) {
      tot[((size_t)warp * C + q * 4 + j) * 2 + 0] = s[j];
      tot[((size_t)warp * C + q * 4 + j) * 2 + 1] = ss[j];
    }
  }
  __syncthreads();
  // GroupNorm(1, C): one group over all channels; warp 0 reduces (fixed order)
  if (threadIdx.x < 32) {
    double sum = 0.0, sumsq = 0.0;
    for (int c = threadIdx.x; c < C; c += 32) {
      double a = 0.0, a2 = 0.0;
      for (int w = 0; w < 8; ++w) {
        a += (double)tot[((size_t)w * C + c) * 2 + 0];
        a2 += (double)tot[((size_t)w * C + c) * 2 + 1];
      }
      sum += (double)(float)a;     // per-channel sums rounded to fp32 like the generic statistics kernels
      sumsq += (double)(float)a2;
    }
    sum = warp_sum_d(sum);
    sumsq = warp_sum_d(sumsq);
    const double n = (double)C * (double)H * (double)W;
    const double mean = sum / n;
    double var = sumsq / n - mean * mean;
    if (var < 0.0) var = 0.0;
    const double rstd = 1.0 / sqrt(var + (double)eps);
    for (int c = threadIdx.x; c < C; c += 32) {
      const double ga = gamma ? (double)gamma[c] : 1.0, be = beta ? (double)beta[c] : 0.0;
      const float sc = (float)(rstd * ga), sh = (float)(be - mean * rstd * ga);
      ab[c * 2 + 0] = sc;
      ab[c * 2 + 1] = sh;
      scale[(int64_t)b * C + c] = sc;
      shift[(int64_t)b * C + c] = sh;
    }
  }
  __syncthreads();
  const float invW = 1.f / (float)W, invH = 1.f / (float)H;
  for (int e = threadIdx.x; e < H * C; e += 256) {
    const int c = e % C;
    pooled_x[(int64_t)b * H * C + e] = fmaf(rowsum[e] * invW, ab[c * 2], ab[c * 2 + 1]);
  }
  for (int e = threadIdx.x; e < W * C; e += 256) {
    const int c = e % C;
    float a = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) a += colpart[(size_t)w * W * C + e];
    pooled_y[(int64_t)b * W * C + e] = fmaf(a * invH, ab[c * 2], ab[c * 2 + 1]);
  }
  // Staged copy for lns_fablock_full_staged (C = 64, 16-bit, power-of-two W, no pad rows): the NORMALISED sample as the byte
  // image of that kernel's shared-memory tile -- row s holds pixel s ^ ((s >> lgw) & 7) (x ^= y & 7 inside every image row),
  // 16-byte chunk ch of a row sits at chunk ch ^ (s & 7) (tcgen05's SWIZZLE_128B) -- so that a head's input is a linear bulk
  // copy.  The second read of the sample hits L1 / L2 (this CTA has just read it).
  if (staged != nullptr) {
    const uint16_t* u16 = reinterpret_cast<const uint16_t*>(u) + (int64_t)b * bstride;
    uint16_t* dst = staged + (int64_t)b * H * W * 64;
    // H * W * 8 chunks of 16 bytes, a multiple of 4 * 256: four independent loads in flight per thread
    for (int e0 = threadIdx.x; e0 < H * W * 8; e0 += 4 * 256) {
      uint4 raw[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const int e = e0 + k * 256, sl = e >> 3, ch = e & 7;
        raw[k] = __ldg(reinterpret_cast<const uint4*>(u16 + (int64_t)(sl ^ ((sl >> lgw) & 7)) * 64 + ch * 8));
      }
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const int e = e0 + k * 256, sl = e >> 3, ch = e & 7;
        const uint32_t rw[4] = {raw[k].x, raw[k].y, raw[k].z, raw[k].w};
        uint32_t o[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float2 f = unpack2_rt(dtype, rw[j]);
          const float4 a4 = *reinterpret_cast<const float4*>(ab + (ch * 8 + 2 * j) * 2);  // scale, shift of two channels
          o[j] = pack2_rt(dtype, fmaf(f.x, a4.x, a4.y), fmaf(f.y, a4.z, a4.w));
        }
        *reinterpret_cast<uint4*>(dst + (int64_t)sl * 64 + ((ch ^ (sl & 7)) << 3)) = make_uint4(o[0], o[1], o[2], o[3]);
      }
    }
  }
}

// v3 of the pre-pass, for the shapes lns_fablock_full_staged covers (C = 64, 16-bit, H, W in {16, 32}): the sample arrives in
// shared memory through bulk copies (one instruction for up to 128 KB) and every later pass reads it from there -- the
// statistics / pooled sums, and the normalised + permuted + swizzled copy the whole-block kernel fetches per head.  v2 read the
// sample twice through the LSU (once for the statistics with two loads in flight per row, once from L2 for the staged copy) and
// ran at 1.2 GB in 0.53 ms; the HBM floor of one read + one write is 0.19 ms.
//   pass A  thread = (pixel lane, 16-byte chunk), pixels pixel lane + k * NTHR/8: its image column is fixed, so it carries the
//           column partial (= its share of sum x) and sum x^2 of 8 channels in registers
//   pass B  warp = image rows, lane = (x mod 4, chunk): row sums, finished with two shuffles per channel
//   then    GroupNorm(1, 64) in fp64 by warp 0 (per-channel partials rounded to fp32 like the generic statistics kernels),
//           pooled outputs, and the staged copy straight from shared memory.
template <int NTHR>
__global__ void __launch_bounds__(NTHR) fablock_prepass3_kernel(const uint16_t* __restrict__ u, int dtype, int H, int W, int lgw, int64_t bstride,
                                                                float eps, const float* __restrict__ gamma, const float* __restrict__ beta,
                                                                float* __restrict__ scale, float* __restrict__ shift,
                                                                float* __restrict__ pooled_x, float* __restrict__ pooled_y,
                                                                uint16_t* __restrict__ staged) {
  extern __shared__ __align__(128) uint8_t sm3[];
  constexpr int nwarp = NTHR / 32, NPL = NTHR / 8;
  const int HW = H * W, NP = NPL / W;  // NP row groups contribute to one image column in pass A
  const uint8_t* S = sm3;                                                  // [HW][64] 16-bit, as in global memory
  float* colpart = reinterpret_cast<float*>(sm3 + (size_t)HW * 128);       // [NP][W][64]
  float* rowsum = colpart + (size_t)NP * W * 64;                           // [H][64]
  float* tot = rowsum + (size_t)H * 64;                                    // [nwarp][64][2]
  float* ab = tot + (size_t)nwarp * 64 * 2;                                // [64][2] scale, shift
  uint64_t* mbar = reinterpret_cast<uint64_t*>(ab + 128);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int b = blockIdx.x;
  const uint32_t S_a = (uint32_t)__cvta_generic_to_shared(sm3), bar = (uint32_t)__cvta_generic_to_shared(mbar);
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    const uint32_t bytes = (uint32_t)HW * 128u, part = bytes / 4u;
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
    const uint8_t* src = reinterpret_cast<const uint8_t*>(u + (int64_t)b * bstride);
    for (int q = 0; q < 4; ++q)
      asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(S_a + q * part),
                   "l"(src + (size_t)q * part), "r"(part), "r"(bar)
                   : "memory");
  }
  __syncthreads();  // the barrier is initialised for every waiter
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "LNSP3_WAIT_%=:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], 0;\n\t"
      "@p bra LNSP3_DONE_%=;\n\t"
      "bra LNSP3_WAIT_%=;\n\t"
      "LNSP3_DONE_%=:\n\t"
      "}" ::"r"(bar)
      : "memory");
  const int ch = lane & 7;
  // ---- pass A: column partials + sum x^2 ----
  {
    const int pl = tid >> 3;
    float cs[8], sq[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) cs[j] = sq[j] = 0.f;
    for (int p0 = pl; p0 < HW; p0 += 4 * NPL) {  // four 16-byte loads in flight
      uint4 raw[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) raw[k] = *reinterpret_cast<const uint4*>(S + (size_t)(p0 + k * NPL) * 128 + ch * 16);
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const uint32_t rw[4] = {raw[k].x, raw[k].y, raw[k].z, raw[k].w};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float2 f = unpack2_rt(dtype, rw[j]);
          cs[2 * j] += f.x; sq[2 * j] = fmaf(f.x, f.x, sq[2 * j]);
          cs[2 * j + 1] += f.y; sq[2 * j + 1] = fmaf(f.y, f.y, sq[2 * j + 1]);
        }
      }
    }
    const int x = pl & (W - 1), rg = pl >> lgw;
    float* cp = colpart + ((size_t)rg * W + x) * 64 + ch * 8;
    *reinterpret_cast<float4*>(cp) = make_float4(cs[0], cs[1], cs[2], cs[3]);
    *reinterpret_cast<float4*>(cp + 4) = make_float4(cs[4], cs[5], cs[6], cs[7]);
#pragma unroll
    for (int j = 0; j < 8; ++j) {  // the 4 pixel lanes of the warp that share this chunk
      cs[j] += __shfl_xor_sync(0xffffffffu, cs[j], 8);
      cs[j] += __shfl_xor_sync(0xffffffffu, cs[j], 16);
      sq[j] += __shfl_xor_sync(0xffffffffu, sq[j], 8);
      sq[j] += __shfl_xor_sync(0xffffffffu, sq[j], 16);
    }
    if (lane < 8) {
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        tot[((size_t)warp * 64 + ch * 8 + j) * 2 + 0] = cs[j];
        tot[((size_t)warp * 64 + ch * 8 + j) * 2 + 1] = sq[j];
      }
    }
  }
  // ---- pass B: row sums ----
  {
    const int xq = lane >> 3;
    for (int y = warp; y < H; y += nwarp) {
      float rs[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) rs[j] = 0.f;
      for (int x0 = xq; x0 < W; x0 += 16) {
        uint4 raw[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) raw[k] = *reinterpret_cast<const uint4*>(S + (size_t)(y * W + x0 + 4 * k) * 128 + ch * 16);
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const uint32_t rw[4] = {raw[k].x, raw[k].y, raw[k].z, raw[k].w};
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const float2 f = unpack2_rt(dtype, rw[j]);
            rs[2 * j] += f.x;
            rs[2 * j + 1] += f.y;
          }
        }
      }
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        rs[j] += __shfl_xor_sync(0xffffffffu, rs[j], 8);
        rs[j] += __shfl_xor_sync(0xffffffffu, rs[j], 16);
      }
      if (lane < 8) {
        *reinterpret_cast<float4*>(rowsum + (size_t)y * 64 + ch * 8) = make_float4(rs[0], rs[1], rs[2], rs[3]);
        *reinterpret_cast<float4*>(rowsum + (size_t)y * 64 + ch * 8 + 4) = make_float4(rs[4], rs[5], rs[6], rs[7]);
      }
    }
  }
  __syncthreads();
  // ---- GroupNorm(1, 64): one group over all channels; warp 0 reduces (fixed order) ----
  if (tid < 32) {
    double sum = 0.0, sumsq = 0.0;
    for (int c = tid; c < 64; c += 32) {
      double a = 0.0, a2 = 0.0;
      for (int w = 0; w < nwarp; ++w) {
        a += (double)tot[((size_t)w * 64 + c) * 2 + 0];
        a2 += (double)tot[((size_t)w * 64 + c) * 2 + 1];
      }
      sum += (double)(float)a;  // per-channel sums rounded to fp32 like the generic statistics kernels
      sumsq += (double)(float)a2;
    }
    sum = warp_sum_d(sum);
    sumsq = warp_sum_d(sumsq);
    const double n = 64.0 * (double)HW;
    const double mean = sum / n;
    double var = sumsq / n - mean * mean;
    if (var < 0.0) var = 0.0;
    const double rstd = 1.0 / sqrt(var + (double)eps);
    for (int c = tid; c < 64; c += 32) {
      const double ga = gamma ? (double)gamma[c] : 1.0, be = beta ? (double)beta[c] : 0.0;
      const float sc = (float)(rstd * ga), sh = (float)(be - mean * rstd * ga);
      ab[c * 2 + 0] = sc;
      ab[c * 2 + 1] = sh;
      scale[(int64_t)b * 64 + c] = sc;
      shift[(int64_t)b * 64 + c] = sh;
    }
  }
  __syncthreads();
  const float invW = 1.f / (float)W, invH = 1.f / (float)H;
  for (int e = tid; e < H * 64; e += NTHR) {
    const int c = e & 63;
    pooled_x[(int64_t)b * H * 64 + e] = fmaf(rowsum[e] * invW, ab[c * 2], ab[c * 2 + 1]);
  }
  for (int e = tid; e < W * 64; e += NTHR) {
    const int c = e & 63;
    float a = 0.f;
    for (int r = 0; r < NP; ++r) a += colpart[(size_t)r * W * 64 + e];
    pooled_y[(int64_t)b * W * 64 + e] = fmaf(a * invH, ab[c * 2], ab[c * 2 + 1]);
  }
  // ---- staged copy: row sl of the image = pixel sl ^ ((sl >> lgw) & 7), chunk ch at ch ^ (sl & 7), normalised ----
  {
    uint16_t* dst = staged + (int64_t)b * HW * 64;
    float4 a4[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) a4[j] = *reinterpret_cast<const float4*>(ab + (ch * 8 + 2 * j) * 2);  // scale, shift of two channels
    for (int sl0 = tid >> 3; sl0 < HW; sl0 += 4 * NPL) {
      uint4 raw[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const int sl = sl0 + k * NPL;
        raw[k] = *reinterpret_cast<const uint4*>(S + (size_t)(sl ^ ((sl >> lgw) & 7)) * 128 + ch * 16);
      }
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const int sl = sl0 + k * NPL;
        const uint32_t rw[4] = {raw[k].x, raw[k].y, raw[k].z, raw[k].w};
        uint32_t o[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float2 f = unpack2_rt(dtype, rw[j]);
          o[j] = pack2_rt(dtype, fmaf(f.x, a4[j].x, a4[j].y), fmaf(f.y, a4[j].z, a4[j].w));
        }
        *reinterpret_cast<uint4*>(dst + (int64_t)sl * 64 + ((ch ^ (sl & 7)) << 3)) = make_uint4(o[0], o[1], o[2], o[3]);
      }
    }
  }
}

static size_t prepass3_smem(int H, int W, int nthr) {
  const int np = (nthr / 8) / W;
  return (size_t)H * W * 128 + ((size_t)np * W * 64 + (size_t)H * 64 + (size_t)(nthr / 32) * 64 * 2 + 128) * sizeof(float) + 16;
}

static size_t fablock_smem(int H, int W) {
  const int HW = H * W, H16 = (H + 15) & ~15, W16 = (W + 15) & ~15;
  size_t bf = (size_t)H * (W * kUS + 8) + 64 * kUS + (size_t)H16 * (H16 + 8) + (size_t)W16 * (W16 + 8);
  return bf * sizeof(__nv_bfloat16) + (64 + 64 * 64 * 2 + 64 * 2) * sizeof(float);
}

}  // namespace lns

extern "C" {

static int fablock_prepass_impl(const void* u, int dtype, int B, int H, int W, int C, int64_t bstride, float eps, const float* gamma,
                                const float* beta, float* scale, float* shift, float* pooled_x, float* pooled_y, void* staged_v,
                                void* stream) {
  LNS_REQUIRE(u && scale && shift && pooled_x && pooled_y && B > 0 && H > 0 && W > 0, "lns_fablock_prepass: bad arguments");
  uint16_t* staged = reinterpret_cast<uint16_t*>(staged_v);
  int lgw = 0;
  if (staged) {
    LNS_REQUIRE(lns::is_h16_host(dtype) && C == 64 && (H == 16 || H == 32) && (W == 16 || W == 32) && bstride % 8 == 0 &&
                    (reinterpret_cast<uintptr_t>(u) & 15) == 0 && (reinterpret_cast<uintptr_t>(staged) & 15) == 0,
                "lns_fablock_prepass_staged: needs 16-bit [B][16|32][16|32][64] input, 16-byte aligned");
    lgw = W == 32 ? 5 : 4;
  }
  int cg = C / 4;
  LNS_REQUIRE(C % 4 == 0 && C >= 4 && C <= 256 && (cg & (cg - 1)) == 0, "lns_fablock_prepass: C must be a power of two in [4,256]");
  LNS_REQUIRE(bstride % 4 == 0, "lns_fablock_prepass: batch stride must be a multiple of 4");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (staged && !getenv("LNS_PREPASS_V2")) {  // (LNS_PREPASS_V2: the LSU pre-pass below also for the staged case, for comparison)
    const uint16_t* u16 = reinterpret_cast<const uint16_t*>(u);
    if (H * W > 256) {
      const size_t smem3 = lns::prepass3_smem(H, W, 512);
      { LNS_OPT_IN_SMEM((lns::fablock_prepass3_kernel<512>), 200 * 1024, "fablock"); }
      lns::fablock_prepass3_kernel<512><<<B, 512, smem3, st>>>(u16, dtype, H, W, lgw, bstride, eps, gamma, beta, scale, shift, pooled_x, pooled_y, staged);
    } else {
      const size_t smem3 = lns::prepass3_smem(H, W, 256);
      { LNS_OPT_IN_SMEM((lns::fablock_prepass3_kernel<256>), 200 * 1024, "fablock"); }
      lns::fablock_prepass3_kernel<256><<<B, 256, smem3, st>>>(u16, dtype, H, W, lgw, bstride, eps, gamma, beta, scale, shift, pooled_x, pooled_y, staged);
    }
    return lns::check_launch("fablock_prepass3_kernel");
  }
  if (C <= 128) {
    const int npl = 128 / C, nx = (W + npl - 1) / npl;
    size_t smem2 = ((size_t)H * C + 8 * (size_t)W * C + 8 * (size_t)C * 2 + 2 * (size_t)C) * sizeof(float);
    if (nx <= 24 && smem2 <= 200 * 1024) {
      {
        LNS_OPT_IN_SMEM((lns::fablock_prepass2_kernel<8>), 200 * 1024, "fablock");
        LNS_OPT_IN_SMEM((lns::fablock_prepass2_kernel<16>), 200 * 1024, "fablock");
        LNS_OPT_IN_SMEM((lns::fablock_prepass2_kernel<24>), 200 * 1024, "fablock");
      }
      if (nx <= 8)
        lns::fablock_prepass2_kernel<8><<<B, 256, smem2, st>>>(u, dtype, H, W, C, bstride, eps, gamma, beta, scale, shift, pooled_x, pooled_y, staged, lgw);
      else if (nx <= 16)
        lns::fablock_prepass2_kernel<16><<<B, 256, smem2, st>>>(u, dtype, H, W, C, bstride, eps, gamma, beta, scale, shift, pooled_x, pooled_y, staged, lgw);
      else
        lns::fablock_prepass2_kernel<24><<<B, 256, smem2, st>>>(u, dtype, H, W, C, bstride, eps, gamma, beta, scale, shift, pooled_x, pooled_y, staged, lgw);
      return lns::check_launch("fablock_prepass2_kernel");
    }
  }
  LNS_REQUIRE(staged == nullptr, "lns_fablock_prepass_staged: shape not covered by the staging pre-pass");
  int rows = 256 / cg;
  LNS_REQUIRE(W <= lns::kPreMaxX * rows, "lns_fablock_prepass: W=%d too wide for C=%d", W, C);
  size_t smem = ((size_t)rows * C + (size_t)H * C + (size_t)W * C + (size_t)rows * C * 2 + 2 * (size_t)C) * sizeof(float);
  LNS_REQUIRE(smem <= 200 * 1024, "lns_fablock_prepass: %dx%dx%d needs %zu B shared memory", H, W, C, smem);
  { LNS_OPT_IN_SMEM((lns::fablock_prepass_kernel), 200 * 1024, "fablock"); }
  lns::fablock_prepass_kernel<<<B, 256, smem, st>>>(u, dtype, H, W, C, bstride, eps, gamma, beta,
                                                    scale, shift, pooled_x, pooled_y);
  return lns::check_launch("fablock_prepass_kernel");
}

int lns_fablock_prepass(const void* u, int dtype, int B, int H, int W, int C, int64_t bstride, float eps, const float* gamma,
                        const float* beta, float* scale, float* shift, float* pooled_x, float* pooled_y, void* stream) {
  return fablock_prepass_impl(u, dtype, B, H, W, C, bstride, eps, gamma, beta, scale, shift, pooled_x, pooled_y, nullptr, stream);
}

int lns_fablock_prepass_staged(const void* u, int dtype, int B, int H, int W, int C, int64_t bstride, float eps, const float* gamma,
                               const float* beta, float* scale, float* shift, float* pooled_x, float* pooled_y, void* staged,
                               void* stream) {
  LNS_REQUIRE(staged, "lns_fablock_prepass_staged: staged must not be null");
  return fablock_prepass_impl(u, dtype, B, H, W, C, bstride, eps, gamma, beta, scale, shift, pooled_x, pooled_y, staged, stream);
}

int lns_fablock_core_supported(int H, int W, int dim, int dim_head) {
  return dim == 64 && dim_head == 64 && H <= 48 && W <= 48 && H >= 1 && W >= 1 && lns::fablock_smem(H, W) <= 227 * 1024;
}

int lns_fablock_core(const void* u, int dtype, int B, int H, int W, int heads, const float* gn_scale, const float* gn_shift,
                     const float* w_in_proj, const float* Kx, const float* Ky, float eps, void* out, void* stream) {
  LNS_REQUIRE(u && gn_scale && gn_shift && w_in_proj && Kx && Ky && out && B > 0 && heads > 0, "lns_fablock_core: bad arguments");
  LNS_REQUIRE(lns::is_h16_host(dtype), "lns_fablock_core: u/out must be LNS_BF16 or LNS_F16 (got dtype %d)", dtype);
  LNS_REQUIRE(lns_fablock_core_supported(H, W, 64, 64), "lns_fablock_core: %dx%d does not fit the fused kernel (use the unfused ops)", H, W);
  LNS_REQUIRE(B <= 65535, "lns_fablock_core: batch %d exceeds grid limit, chunk the call", B);
  LNS_REQUIRE((reinterpret_cast<uintptr_t>(u) & 15) == 0 && (reinterpret_cast<uintptr_t>(out) & 15) == 0, "lns_fablock_core: alignment");
  size_t smem = lns::fablock_smem(H, W);
  {
    LNS_OPT_IN_SMEM((lns::fablock_core_kernel<512, false>), 227 * 1024, "fablock");
    LNS_OPT_IN_SMEM((lns::fablock_core_kernel<256, false>), 227 * 1024, "fablock");
    LNS_OPT_IN_SMEM((lns::fablock_core_kernel<512, true>), 227 * 1024, "fablock");
    LNS_OPT_IN_SMEM((lns::fablock_core_kernel<256, true>), 227 * 1024, "fablock");
  }
  dim3 grid(heads, B);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const __nv_bfloat16* up = reinterpret_cast<const __nv_bfloat16*>(u);
  __nv_bfloat16* op = reinterpret_cast<__nv_bfloat16*>(out);
  const bool f16 = dtype == LNS_F16;
  if (H <= 32 && W <= 32 && H * W > 256) {  // <= 2 k-tiles per axis: fits 128 registers per thread; small images: 2+ CTAs/SM
    auto kern = f16 ? lns::fablock_core_kernel<512, true> : lns::fablock_core_kernel<512, false>;
    kern<<<grid, 512, smem, st>>>(up, H, W, heads, gn_scale, gn_shift, w_in_proj, Kx, Ky, eps, op);
  } else {
    auto kern = f16 ? lns::fablock_core_kernel<256, true> : lns::fablock_core_kernel<256, false>;
    kern<<<grid, 256, smem, st>>>(up, H, W, heads, gn_scale, gn_shift, w_in_proj, Kx, Ky, eps, op);
  }
  return lns::check_launch("fablock_core_kernel");
}

}  // extern "C"
