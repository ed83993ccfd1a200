// Shared device/host helpers for liblns_b200.so (sm_100a only).
#pragma once
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include <atomic>

#include "../../include/lns_b200.h"

namespace lns {

// ---- error plumbing ------------------------------------------------------------------------------
void set_error(const char* fmt, ...);
int check_launch(const char* what);  // cudaGetLastError -> LNS_E_CUDA (+ message) or LNS_OK

// Dynamic shared memory above 48 KB must be opted into per (kernel, DEVICE): one flag bit per device ordinal, set after the
// first successful cudaFuncSetAttribute on that device (idempotent, so a race between two host threads is harmless).
struct SmemOptIn {
  std::atomic<uint64_t> done[4];
};
int opt_in_smem(const void* func, int bytes, SmemOptIn& st, const char* what);
// one static flag set per call site (inside a template: per instantiation)
#define LNS_OPT_IN_SMEM(kern, bytes, what)                                                     \
  do {                                                                                         \
    static lns::SmemOptIn lns_opt_in_state_;                                                   \
    const int lns_opt_in_rc_ = lns::opt_in_smem(reinterpret_cast<const void*>(kern), (bytes), lns_opt_in_state_, (what)); \
    if (lns_opt_in_rc_ != LNS_OK) return lns_opt_in_rc_;                                       \
  } while (0)
int device_sm_count();  // SM count of the CURRENT device (cached per device ordinal)

#define LNS_REQUIRE(cond, ...)        \
  do {                                \
    if (!(cond)) {                    \
      lns::set_error(__VA_ARGS__);    \
      return LNS_E_INVALID;           \
    }                                 \
  } while (0)

// ---- dtype helpers -------------------------------------------------------------------------------
// LNS_TF32 storage = fp32 words whose values were rounded to nearest TF32 when written (loads are plain fp32 loads)
__device__ __forceinline__ float round_tf32(float v) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(v));
  return __uint_as_float(r);
}
// 16-bit storage: LNS_BF16 (8-bit mantissa, fp32 range) or LNS_F16 (11-bit mantissa; conversions SATURATE to +-65504
// instead of overflowing to inf).  Both feed tcgen05.mma.kind::f16 / mma.sync at the same rate; kernels take the format as
// a template flag (F16) or, in the byte-moving ones, as the runtime dtype.  Pointers to either are typed __nv_bfloat16*
// (an opaque 16-bit element) -- only these helpers interpret the bits.
__device__ __forceinline__ bool is_h16(int dtype) { return dtype == LNS_BF16 || dtype == LNS_F16; }
static inline bool is_h16_host(int dtype) { return dtype == LNS_BF16 || dtype == LNS_F16; }
template <bool F16>
__device__ __forceinline__ uint32_t pack2_h16(float lo, float hi) {
  uint32_t r;
  if (F16) {
    asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
  } else {
    __nv_bfloat162 h = __floats2bfloat162_rn(lo, hi);
    r = *reinterpret_cast<uint32_t*>(&h);
  }
  return r;
}
template <bool F16>
__device__ __forceinline__ float2 unpack2_h16(uint32_t raw) {
  if (F16) return __half22float2(*reinterpret_cast<__half2*>(&raw));
  return __bfloat1622float2(*reinterpret_cast<__nv_bfloat162*>(&raw));
}
template <bool F16>
__device__ __forceinline__ uint16_t to_h16(float v) { return (uint16_t)(pack2_h16<F16>(v, 0.f) & 0xFFFFu); }
template <bool F16>
__device__ __forceinline__ float from_h16(uint16_t raw) { return unpack2_h16<F16>((uint32_t)raw).x; }
__device__ __forceinline__ uint32_t pack2_rt(int dtype, float lo, float hi) {
  return dtype == LNS_F16 ? pack2_h16<true>(lo, hi) : pack2_h16<false>(lo, hi);
}
__device__ __forceinline__ float2 unpack2_rt(int dtype, uint32_t raw) {
  return dtype == LNS_F16 ? unpack2_h16<true>(raw) : unpack2_h16<false>(raw);
}

__device__ __forceinline__ float ld_as_float(const void* p, int dtype, int64_t i) {
  if (!is_h16(dtype)) return __ldg(reinterpret_cast<const float*>(p) + i);
  const uint16_t raw = reinterpret_cast<const uint16_t*>(p)[i];
  return dtype == LNS_F16 ? from_h16<true>(raw) : from_h16<false>(raw);
}
__device__ __forceinline__ void st_from_float(void* p, int dtype, int64_t i, float v) {
  if (dtype == LNS_F32)
    reinterpret_cast<float*>(p)[i] = v;
  else if (dtype == LNS_TF32)
    reinterpret_cast<float*>(p)[i] = round_tf32(v);
  else
    reinterpret_cast<uint16_t*>(p)[i] = dtype == LNS_F16 ? to_h16<true>(v) : to_h16<false>(v);
}
// 4 consecutive elements (i must be a multiple of 4 and the pointer suitably aligned)
__device__ __forceinline__ float4 ld4_as_float(const void* p, int dtype, int64_t i) {
  if (!is_h16(dtype)) return __ldg(reinterpret_cast<const float4*>(reinterpret_cast<const float*>(p) + i));
  uint2 raw = __ldg(reinterpret_cast<const uint2*>(reinterpret_cast<const uint16_t*>(p) + i));
  float2 fa = unpack2_rt(dtype, raw.x), fb = unpack2_rt(dtype, raw.y);
  return make_float4(fa.x, fa.y, fb.x, fb.y);
}
__device__ __forceinline__ void st4_from_float(void* p, int dtype, int64_t i, float4 v) {
  if (dtype == LNS_F32) {
    *reinterpret_cast<float4*>(reinterpret_cast<float*>(p) + i) = v;
  } else if (dtype == LNS_TF32) {
    *reinterpret_cast<float4*>(reinterpret_cast<float*>(p) + i) =
        make_float4(round_tf32(v.x), round_tf32(v.y), round_tf32(v.z), round_tf32(v.w));
  } else {
    uint2 raw;
    raw.x = pack2_rt(dtype, v.x, v.y);
    raw.y = pack2_rt(dtype, v.z, v.w);
    *reinterpret_cast<uint2*>(reinterpret_cast<uint16_t*>(p) + i) = raw;
  }
}
// N independent 4-element loads with ALL loads issued before the first conversion.  (Calling ld4_as_float in an unrolled
// loop under a per-element predicate makes ptxas emit load -> convert -> next load: one exposed memory latency per element,
// found with ncu in the GroupNorm and FABlock pre-pass kernels.)  off[i] must be a valid (clamped) element offset even when
// ok[i] is false; invalid elements come back as zeros.
template <int N>
__device__ __forceinline__ void ld4n_as_float(const void* p, int dtype, const int64_t (&off)[N], const bool (&ok)[N], float4 (&v)[N]) {
  if (is_h16(dtype)) {
    uint2 raw[N];
#pragma unroll
    for (int i = 0; i < N; ++i) raw[i] = __ldg(reinterpret_cast<const uint2*>(reinterpret_cast<const uint16_t*>(p) + off[i]));
#pragma unroll
    for (int i = 0; i < N; ++i) {
      const float2 fa = unpack2_rt(dtype, raw[i].x), fb = unpack2_rt(dtype, raw[i].y);
      v[i] = ok[i] ? make_float4(fa.x, fa.y, fb.x, fb.y) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
  } else {
#pragma unroll
    for (int i = 0; i < N; ++i) v[i] = __ldg(reinterpret_cast<const float4*>(reinterpret_cast<const float*>(p) + off[i]));
#pragma unroll
    for (int i = 0; i < N; ++i)
      if (!ok[i]) v[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  }
}
__host__ __device__ __forceinline__ int dtype_size(int dtype) { return (dtype == LNS_BF16 || dtype == LNS_F16) ? 2 : 4; }

// ---- activations (exact forms, matching torch) ----------------------------------------------------
// Swish  x*sigmoid(x)   modules/basics.py:27-29 ; nn.GELU() exact erf  train_stage2_ns2d.py:36
__device__ __forceinline__ float act_silu(float x) { return x / (1.0f + expf(-x)); }
__device__ __forceinline__ float act_gelu(float x) { return 0.5f * x * (1.0f + erff(x * 0.70710678118654752440f)); }
__device__ __forceinline__ float apply_act(float x, int act) {
  if (act == LNS_ACT_SILU) return act_silu(x);
  if (act == LNS_ACT_GELU) return act_gelu(x);
  return x;
}
// bf16 path: results are rounded to bf16 (2^-9) anyway, so the activations use the SFU fast paths
// (ex2.approx / rcp.approx, ~2^-21 relative) and a 1.5e-7-accurate erf (Abramowitz-Stegun 7.1.26).  The exact forms
// above stay on the fp32 validation path.  (The exact expf/erff made the elementwise kernels compute bound.)
__device__ __forceinline__ float act_silu_fast(float x) { return __fdividef(x, 1.0f + __expf(-x)); }
__device__ __forceinline__ float act_gelu_fast(float x) {
  const float z = x * 0.70710678118654752440f;
  const float a = fabsf(z);
  const float t = __fdividef(1.0f, fmaf(0.3275911f, a, 1.0f));
  float p = fmaf(t, 1.061405429f, -1.453152027f);
  p = fmaf(t, p, 1.421413741f);
  p = fmaf(t, p, -0.284496736f);
  p = fmaf(t, p, 0.254829592f);
  const float e = 1.0f - p * t * __expf(-a * a);
  return 0.5f * x * (1.0f + copysignf(e, z));
}
__device__ __forceinline__ float apply_act_fast(float x, int act) {
  if (act == LNS_ACT_SILU) return act_silu_fast(x);
  if (act == LNS_ACT_GELU) return act_gelu_fast(x);
  return x;
}
// exact on fp32 storage, fast on 16-bit storage
__device__ __forceinline__ float apply_act_for(float x, int act, int storage_dtype) {
  return is_h16(storage_dtype) ? apply_act_fast(x, act) : apply_act(x, act);
}

// ---- reductions ----------------------------------------------------------------------------------
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ double warp_sum_d(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

static inline int cdiv(int64_t a, int64_t b) { return (int)((a + b - 1) / b); }

// ---- conv index map shared by both conv engines -----------------------------------------------------
struct ConvGeom {
  int B, Hin, Win, Cin, Hv, Wv;
  int KH, KW, stride, dil, pad_t, pad_l, circ_h, circ_w;
  int Hout, Wout, Cout;
  int64_t x_bstride, y_bstride;
};
// source pixel (ys,xs) of virtual coordinate (yv,xv); returns false when the tap falls in zero padding
__device__ __forceinline__ bool conv_src(const ConvGeom& g, int yv, int xv, int& ys, int& xs) {
  if (g.circ_h) {
    yv = yv % g.Hv;
    if (yv < 0) yv += g.Hv;
  } else if (yv < 0 || yv >= g.Hv) {
    return false;
  }
  if (g.circ_w) {
    xv = xv % g.Wv;
    if (xv < 0) xv += g.Wv;
  } else if (xv < 0 || xv >= g.Wv) {
    return false;
  }
  ys = (g.Hv == g.Hin) ? yv : (int)(((int64_t)yv * g.Hin) / g.Hv);
  xs = (g.Wv == g.Win) ? xv : (int)(((int64_t)xv * g.Win) / g.Wv);
  return true;
}

int conv2d_simt(const LnsConvDesc* d, cudaStream_t stream);
int conv2d_umma(const LnsConvDesc* d, cudaStream_t stream);
int conv2d_halo(const LnsConvDesc* d, cudaStream_t stream);
bool conv_halo_supported(const LnsConvDesc* d);
int conv2d_latent(const LnsConvDesc* d, cudaStream_t stream);
int conv2d_coarse(const LnsConvDesc* d, cudaStream_t stream);
bool conv_coarse_supported(const LnsConvDesc* d);
bool conv_latent_supported(const LnsConvDesc* d);
int validate_conv(const LnsConvDesc* d);
ConvGeom make_geom(const LnsConvDesc* d);

}  // namespace lns
