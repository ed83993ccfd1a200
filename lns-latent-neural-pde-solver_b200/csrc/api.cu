// liblns_b200.so: error plumbing, version, device info and the lns_conv2d dispatcher.
#include <stdarg.h>
#include <string.h>

#include "common.cuh"

namespace lns {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int check_launch(const char* what) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    set_error("%s: CUDA error %d (%s)", what, (int)e, cudaGetErrorString(e));
    return LNS_E_CUDA;
  }
  return LNS_OK;
}

int opt_in_smem(const void* func, int bytes, SmemOptIn& st, const char* what) {
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) {
    set_error("%s: cudaGetDevice: %s", what, cudaGetErrorString(e));
    return LNS_E_CUDA;
  }
  const uint64_t bit = 1ull << (dev & 63);
  std::atomic<uint64_t>& word = st.done[(dev >> 6) & 3];
  if (word.load(std::memory_order_acquire) & bit) return LNS_OK;
  e = cudaFuncSetAttribute(func, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
  if (e != cudaSuccess) {
    set_error("%s: cudaFuncSetAttribute(%d B of dynamic shared memory) on device %d: %s", what, bytes, dev, cudaGetErrorString(e));
    return LNS_E_CUDA;
  }
  word.fetch_or(bit, std::memory_order_release);
  return LNS_OK;
}

int device_sm_count() {
  static std::atomic<int> cache[256];
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return 148;
  std::atomic<int>& c = cache[dev & 255];
  int v = c.load(std::memory_order_relaxed);
  if (v <= 0) {
    if (cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || v <= 0) v = 148;
    c.store(v, std::memory_order_relaxed);
  }
  return v;
}

ConvGeom make_geom(const LnsConvDesc* d) {
  ConvGeom g;
  g.B = d->B; g.Hin = d->Hin; g.Win = d->Win; g.Cin = d->Cin; g.Hv = d->Hv; g.Wv = d->Wv;
  g.KH = d->KH; g.KW = d->KW; g.stride = d->stride; g.dil = d->dil; g.pad_t = d->pad_t; g.pad_l = d->pad_l;
  g.circ_h = d->pad_mode_h == LNS_PAD_CIRCULAR; g.circ_w = d->pad_mode_w == LNS_PAD_CIRCULAR;
  g.Hout = d->Hout; g.Wout = d->Wout; g.Cout = d->Cout;
  g.x_bstride = d->x_bstride; g.y_bstride = d->y_bstride;
  return g;
}

int validate_conv(const LnsConvDesc* d) {
  LNS_REQUIRE(d != nullptr, "lns_conv2d: null descriptor");
  LNS_REQUIRE(d->x && d->y && d->w, "lns_conv2d: null x/y/w pointer");
  LNS_REQUIRE(d->B > 0 && d->Hin > 0 && d->Win > 0 && d->Cin > 0, "lns_conv2d: bad input shape %d %d %d %d", d->B,
              d->Hin, d->Win, d->Cin);
  LNS_REQUIRE(d->Hv >= d->Hin && d->Wv >= d->Win, "lns_conv2d: virtual size (%d,%d) smaller than input (%d,%d)",
              d->Hv, d->Wv, d->Hin, d->Win);
  LNS_REQUIRE(d->Hout > 0 && d->Wout > 0 && d->Cout > 0, "lns_conv2d: bad output shape");
  LNS_REQUIRE(d->KH >= 1 && d->KH <= 7 && d->KW >= 1 && d->KW <= 7, "lns_conv2d: unsupported filter %dx%d", d->KH,
              d->KW);
  LNS_REQUIRE(d->stride >= 1 && d->dil >= 1, "lns_conv2d: bad stride/dilation");
  LNS_REQUIRE(d->x_dtype >= LNS_F32 && d->x_dtype <= LNS_F16 && d->y_dtype >= LNS_F32 && d->y_dtype <= LNS_F16,
              "lns_conv2d: bad dtype");
  // every output pixel's taps must stay inside the padded virtual input
  int64_t ymax = (int64_t)(d->Hout - 1) * d->stride + (int64_t)(d->KH - 1) * d->dil - d->pad_t;
  int64_t xmax = (int64_t)(d->Wout - 1) * d->stride + (int64_t)(d->KW - 1) * d->dil - d->pad_l;
  LNS_REQUIRE(d->pad_mode_h == LNS_PAD_CIRCULAR || ymax < (int64_t)d->Hv + d->Hv,
              "lns_conv2d: output height %d inconsistent with input", d->Hout);
  LNS_REQUIRE(d->pad_mode_w == LNS_PAD_CIRCULAR || xmax < (int64_t)d->Wv + d->Wv,
              "lns_conv2d: output width %d inconsistent with input", d->Wout);
  if (d->x_layout == LNS_NCHW) LNS_REQUIRE(!lns::is_h16_host(d->x_dtype), "lns_conv2d: NCHW input must be fp32");
  if (d->y_layout == LNS_NCHW) LNS_REQUIRE(!lns::is_h16_host(d->y_dtype), "lns_conv2d: NCHW output must be fp32");
  return LNS_OK;
}

}  // namespace lns

extern "C" {

const char* lns_version(void) { return "lns_b200 0.2.1 (sm_100a)"; }
const char* lns_last_error(void) { return lns::g_err; }

int lns_device_info(int* sm_count, int* cc_major, int* cc_minor) {
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) {
    lns::set_error("lns_device_info: %s", cudaGetErrorString(e));
    return LNS_E_CUDA;
  }
  cudaDeviceProp p;
  e = cudaGetDeviceProperties(&p, dev);
  if (e != cudaSuccess) {
    lns::set_error("lns_device_info: %s", cudaGetErrorString(e));
    return LNS_E_CUDA;
  }
  if (sm_count) *sm_count = p.multiProcessorCount;
  if (cc_major) *cc_major = p.major;
  if (cc_minor) *cc_minor = p.minor;
  return LNS_OK;
}

int lns_conv_stats_chunks(int Hout, int Wout) { return ((Hout + 7) / 8) * ((Wout + 7) / 8) * 4; }

int lns_conv2d(const LnsConvDesc* d, void* stream) {
  int rc = lns::validate_conv(d);
  if (rc != LNS_OK) return rc;
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  if (d->stats && d->engine != LNS_ENGINE_COARSE) {
    lns::set_error("lns_conv2d: output statistics (LnsConvDesc.stats) are produced by LNS_ENGINE_COARSE only");
    return LNS_E_UNSUPPORTED;
  }
  if (d->engine == LNS_ENGINE_HALO) return lns::conv2d_halo(d, s);
  if (d->engine == LNS_ENGINE_LATENT) return lns::conv2d_latent(d, s);
  if (d->engine == LNS_ENGINE_COARSE) return lns::conv2d_coarse(d, s);
  if (d->engine == LNS_ENGINE_UMMA) return lns::conv2d_umma(d, s);
  if (d->engine == LNS_ENGINE_SIMT) return lns::conv2d_simt(d, s);
  lns::set_error("lns_conv2d: unknown engine %d", d->engine);
  return LNS_E_INVALID;
}

}  // extern "C"
