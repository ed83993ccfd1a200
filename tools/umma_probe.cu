// Probe: can a tcgen05 K-major SWIZZLE_128B A-descriptor start at an arbitrary 128-byte multiple inside a larger,
// address-swizzled pixel array (implicit-GEMM "halo" trick)?  Tests base_offset variants and SBO values.
//   smem halo: NPIX pixels x 128 B (64 bf16), chunk c of pixel q stored at q*128 + ((c ^ (q & 7)) << 4)
//   B = 64x64 identity (K-major, swizzled) -> D[m][n] = A[m][n];  expected A row m = pixel(shift + (m/8)*pitch + m%8)
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 tools/umma_probe.cu -o /tmp/umma_probe && /tmp/umma_probe
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <vector>

#define NPIX 512

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__global__ void probe(const __nv_bfloat16* X, float* D, int shift, int pitch, int bo_mode, int kstep) {
  extern __shared__ uint8_t raw[];
  uint32_t base = (smem_u32(raw) + 1023u) & ~1023u;
  uint8_t* gen = raw + (base - smem_u32(raw));
  uint8_t* halo = gen;                   // NPIX * 128
  uint8_t* bt = gen + NPIX * 128;        // 64 * 128
  uint64_t* bar = (uint64_t*)(gen + NPIX * 128 + 64 * 128);
  uint32_t* tslot = (uint32_t*)(bar + 1);
  int tid = threadIdx.x;
  for (int e = tid; e < NPIX * 8; e += blockDim.x) {
    int q = e >> 3, c = e & 7;
    *(uint4*)(halo + q * 128 + ((c ^ (q & 7)) << 4)) = *(const uint4*)(X + q * 64 + c * 8);
  }
  for (int e = tid; e < 64 * 8; e += blockDim.x) {
    int n = e >> 3, c = e & 7;
    __nv_bfloat16 v[8];
    for (int j = 0; j < 8; ++j) v[j] = __float2bfloat16((c * 8 + j) == n ? 1.f : 0.f);
    *(uint4*)(bt + n * 128 + ((c ^ (n & 7)) << 4)) = *(uint4*)v;
  }
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(bar)));
    asm volatile("fence.mbarrier_init.release.cluster;");
  }
  if (tid < 32) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 64;" ::"r"(smem_u32(tslot)));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;");
  uint32_t tmem = *(volatile uint32_t*)tslot;
  if (tid == 0) {
    uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((64u >> 3) << 17) | ((128u >> 4) << 24);
    for (int k = 0; k < 4; ++k) {
      int kk = kstep ? k : k;  // K advance: +32 B per step
      uint32_t a_addr = base + shift * 128 + kk * 32;
      uint64_t ad = (uint64_t)((a_addr & 0x3FFFFu) >> 4);
      ad |= (uint64_t)((pitch * 128u) >> 4) << 32;
      ad |= 1ull << 46;
      ad |= 2ull << 61;
      uint32_t bo = 0;
      if (bo_mode == 1) bo = (a_addr >> 7) & 7;
      if (bo_mode == 2) bo = ((base + shift * 128) >> 7) & 7;
      ad |= (uint64_t)bo << 49;
      uint32_t b_addr = base + NPIX * 128 + kk * 32;
      uint64_t bd = (uint64_t)((b_addr & 0x3FFFFu) >> 4);
      bd |= (uint64_t)(1024u >> 4) << 32;
      bd |= 1ull << 46;
      bd |= 2ull << 61;
      uint32_t acc = k > 0;
      asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}"
                   ::"r"(tmem), "l"(ad), "l"(bd), "r"(idesc), "r"(acc) : "memory");
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
  }
  // wait
  {
    uint32_t b = smem_u32(bar);
    asm volatile("{\n.reg .pred p;\nW:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], 0;\n@p bra Dn;\nbra W;\nDn:\n}" ::"r"(b) : "memory");
  }
  asm volatile("tcgen05.fence::after_thread_sync;");
  if (tid < 128) {
    int warp = tid >> 5;
    for (int c0 = 0; c0 < 64; c0 += 16) {
      uint32_t r[16];
      uint32_t ta = tmem + ((uint32_t)(warp * 32) << 16) + c0;
      asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                   : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                     "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                   : "r"(ta) : "memory");
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      for (int j = 0; j < 16; ++j) D[tid * 64 + c0 + j] = __uint_as_float(r[j]);
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  if (tid < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 64;" ::"r"(tmem));
}

int main() {
  std::vector<__nv_bfloat16> hx(NPIX * 64);
  for (int q = 0; q < NPIX; ++q)
    for (int c = 0; c < 64; ++c) hx[q * 64 + c] = __float2bfloat16((float)(q + c * 0.001953125f * 0 + (c % 8) * 0.0f) + 0.0f * c);
  // value encodes pixel index only in the integer part; channel in a second pass below
  for (int q = 0; q < NPIX; ++q)
    for (int c = 0; c < 64; ++c) hx[q * 64 + c] = __float2bfloat16((float)((q * 7 + c * 3) % 251));
  __nv_bfloat16* dX; float* dD;
  cudaMalloc(&dX, hx.size() * 2); cudaMalloc(&dD, 128 * 64 * 4);
  cudaMemcpy(dX, hx.data(), hx.size() * 2, cudaMemcpyHostToDevice);
  int smem = NPIX * 128 + 64 * 128 + 64 + 1024;
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  std::vector<float> hd(128 * 64);
  int pitches[] = {8, 16, 10, 12, 24};
  for (int bo_mode = 0; bo_mode < 3; ++bo_mode)
    for (int pi = 0; pi < 5; ++pi) {
      int pitch = pitches[pi];
      printf("bo_mode %d pitch %2d: ", bo_mode, pitch);
      for (int shift = 0; shift < 20; ++shift) {
        cudaMemset(dD, 0, 128 * 64 * 4);
        probe<<<1, 128, smem>>>(dX, dD, shift, pitch, bo_mode, 1);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("CUDA error %s\n", cudaGetErrorString(e)); return 1; }
        cudaMemcpy(hd.data(), dD, 128 * 64 * 4, cudaMemcpyDeviceToHost);
        int bad = 0;
        for (int m = 0; m < 128; ++m) {
          int q = shift + (m / 8) * pitch + (m % 8);
          for (int n = 0; n < 64; ++n) {
            float exp = (float)((q * 7 + n * 3) % 251);
            if (hd[m * 64 + n] != exp) ++bad;
          }
        }
        printf("%s", bad == 0 ? "." : "X");
      }
      printf("\n");
    }
  return 0;
}
