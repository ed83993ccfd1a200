// Shared device/host helpers for liblns_b200.so (sm_100a only).
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/lns_b200.h"

namespace lns {

// ---- error plumbing ------------------------------------------------------------------------------
void set_error(const char* fmt, ...);
int check_launch(const char* what);  // cudaGetLastError -> LNS_E_CUDA (+ message) or LNS_OK

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
__device__ __forceinline__ float ld_as_float(const void* p, int dtype, int64_t i) {
  if (dtype != LNS_BF16) return __ldg(reinterpret_cast<const float*>(p) + i);
  return __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(p)[i]);
}
__device__ __forceinline__ void st_from_float(void* p, int dtype, int64_t i, float v) {
  if (dtype == LNS_F32)
    reinterpret_cast<float*>(p)[i] = v;
  else if (dtype == LNS_TF32)
    reinterpret_cast<float*>(p)[i] = round_tf32(v);
  else
    reinterpret_cast<__nv_bfloat16*>(p)[i] = __float2bfloat16_rn(v);
}
// 4 consecutive elements (i must be a multiple of 4 and the pointer suitably aligned)
__device__ __forceinline__ float4 ld4_as_float(const void* p, int dtype, int64_t i) {
  if (dtype != LNS_BF16) return __ldg(reinterpret_cast<const float4*>(reinterpret_cast<const float*>(p) + i));
  uint2 raw = __ldg(reinterpret_cast<const uint2*>(reinterpret_cast<const __nv_bfloat16*>(p) + i));
  __nv_bfloat162 a = *reinterpret_cast<__nv_bfloat162*>(&raw.x);
  __nv_bfloat162 b = *reinterpret_cast<__nv_bfloat162*>(&raw.y);
  float2 fa = __bfloat1622float2(a), fb = __bfloat1622float2(b);
  return make_float4(fa.x, fa.y, fb.x, fb.y);
}
__device__ __forceinline__ void st4_from_float(void* p, int dtype, int64_t i, float4 v) {
  if (dtype == LNS_F32) {
    *reinterpret_cast<float4*>(reinterpret_cast<float*>(p) + i) = v;
  } else if (dtype == LNS_TF32) {
    *reinterpret_cast<float4*>(reinterpret_cast<float*>(p) + i) =
        make_float4(round_tf32(v.x), round_tf32(v.y), round_tf32(v.z), round_tf32(v.w));
  } else {
    __nv_bfloat162 a = __floats2bfloat162_rn(v.x, v.y);
    __nv_bfloat162 b = __floats2bfloat162_rn(v.z, v.w);
    uint2 raw;
    raw.x = *reinterpret_cast<uint32_t*>(&a);
    raw.y = *reinterpret_cast<uint32_t*>(&b);
    *reinterpret_cast<uint2*>(reinterpret_cast<__nv_bfloat16*>(p) + i) = raw;
  }
}
__host__ __device__ __forceinline__ int dtype_size(int dtype) { return dtype == LNS_BF16 ? 2 : 4; }

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
// exact on fp32 storage, fast on bf16 storage
__device__ __forceinline__ float apply_act_for(float x, int act, int storage_dtype) {
  return storage_dtype == LNS_BF16 ? apply_act_fast(x, act) : apply_act(x, act);
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
int validate_conv(const LnsConvDesc* d);
ConvGeom make_geom(const LnsConvDesc* d);

}  // namespace lns
