// Probe for the tensor-core FABlock kernel (csrc/fablock_tc.cu): tcgen05.mma with an MN-MAJOR B operand.
//   D[128 x 64] = A[128 x 128] * B[128 x 64]
//   A: K-major SWIZZLE_128B, two 64-column slabs of [128 rows x 128 B] (what the block-diagonal axial kernels use)
//   B: the pixel-row image U [K = 128 pixel rows][N = 64 channels], one 128-byte row per pixel, 16-byte chunk c of row p stored at
//      p*128 + ((c ^ (p & 7)) << 4) -- i.e. exactly the layout every other kernel writes.  Read as an MN-major SWIZZLE_128B operand
//      (instruction-descriptor bit 16): canonical form ((8,n),(8,k)):((1,LBO),(8,SBO)) in 16-byte units -> one 128-byte row per k,
//      8-row atoms SBO = 1024 B apart, K advance of 16 = +2048 B.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 tools/umma_mn_probe.cu -o /tmp/umma_mn_probe && /tmp/umma_mn_probe
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <vector>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// A_g: [128][128] half row-major; U_g: [128][64] half row-major (pixel rows)
__global__ void probe(const __half* A_g, const __half* U_g, float* D, uint32_t lbo_bytes, uint32_t sbo_bytes, int a_mn) {
  extern __shared__ uint8_t raw[];
  uint32_t base = (smem_u32(raw) + 1023u) & ~1023u;
  uint8_t* gen = raw + (base - smem_u32(raw));
  uint8_t* At = gen;                 // 2 slabs x 128 rows x 128 B = 32 KB
  uint8_t* Ut = gen + 32768;         // 128 rows x 128 B = 16 KB
  uint64_t* bar = (uint64_t*)(gen + 32768 + 16384);
  uint32_t* tslot = (uint32_t*)(bar + 1);
  int tid = threadIdx.x;
  for (int e = tid; e < 128 * 16; e += blockDim.x) {  // A: row r, 16-byte chunk kc (8 halves), slab = kc / 8
    int r = e >> 4, kc = e & 15, slab = kc >> 3, c = kc & 7;
    *(uint4*)(At + slab * 16384 + r * 128 + ((c ^ (r & 7)) << 4)) = *(const uint4*)(A_g + r * 128 + kc * 8);
  }
  for (int e = tid; e < 128 * 8; e += blockDim.x) {
    int p = e >> 3, c = e & 7;
    *(uint4*)(Ut + p * 128 + ((c ^ (p & 7)) << 4)) = *(const uint4*)(U_g + p * 64 + c * 8);
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
    // kind::f16, D = f32, A = B = f16 (format 0), A K-major (bit 15 = 0), B MN-major (bit 16 = 1), N = 64, M = 128
    uint32_t idesc = (1u << 4) | (1u << 16) | ((64u >> 3) << 17) | ((128u >> 4) << 24);
    (void)a_mn;
    for (int ks = 0; ks < 8; ++ks) {
      uint32_t a_addr = base + (ks >> 2) * 16384 + (ks & 3) * 32;
      uint64_t ad = (uint64_t)((a_addr & 0x3FFFFu) >> 4) | ((uint64_t)(1024u >> 4) << 32) | (1ull << 46) | (2ull << 61);
      uint32_t b_addr = base + 32768 + ks * 2048;  // 16 pixel rows per K step
      uint64_t bd = (uint64_t)((b_addr & 0x3FFFFu) >> 4) | ((uint64_t)(lbo_bytes >> 4) << 16) | ((uint64_t)(sbo_bytes >> 4) << 32) |
                    (1ull << 46) | (2ull << 61);
      uint32_t acc = ks > 0;
      asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}"
                   ::"r"(tmem), "l"(ad), "l"(bd), "r"(idesc), "r"(acc) : "memory");
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
  }
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
  std::vector<__half> hA(128 * 128), hU(128 * 64);
  std::vector<float> fA(128 * 128), fU(128 * 64);
  srand(7);
  for (size_t i = 0; i < hA.size(); ++i) { float v = (rand() % 17 - 8) / 8.0f; hA[i] = __float2half(v); fA[i] = __half2float(hA[i]); }
  for (size_t i = 0; i < hU.size(); ++i) { float v = (rand() % 33 - 16) / 16.0f; hU[i] = __float2half(v); fU[i] = __half2float(hU[i]); }
  std::vector<double> ref(128 * 64, 0.0);
  for (int m = 0; m < 128; ++m)
    for (int k = 0; k < 128; ++k)
      for (int n = 0; n < 64; ++n) ref[m * 64 + n] += (double)fA[m * 128 + k] * fU[k * 64 + n];
  __half *dA, *dU; float* dD;
  cudaMalloc(&dA, hA.size() * 2); cudaMalloc(&dU, hU.size() * 2); cudaMalloc(&dD, 128 * 64 * 4);
  cudaMemcpy(dA, hA.data(), hA.size() * 2, cudaMemcpyHostToDevice);
  cudaMemcpy(dU, hU.data(), hU.size() * 2, cudaMemcpyHostToDevice);
  int smem = 32768 + 16384 + 64 + 1024;
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  std::vector<float> hd(128 * 64);
  struct { uint32_t lbo, sbo; } cfgs[] = {{0, 1024}, {1024, 1024}, {16, 1024}, {2048, 1024}, {1024, 2048}, {128, 1024}, {1024, 128}};
  for (auto c : cfgs) {
    cudaMemset(dD, 0, 128 * 64 * 4);
    probe<<<1, 128, smem>>>(dA, dU, dD, c.lbo, c.sbo, 0);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("lbo %u sbo %u: CUDA error %s\n", c.lbo, c.sbo, cudaGetErrorString(e)); return 1; }
    cudaMemcpy(hd.data(), dD, 128 * 64 * 4, cudaMemcpyDeviceToHost);
    double worst = 0; int bad = 0;
    for (int i = 0; i < 128 * 64; ++i) { double d = fabs(hd[i] - ref[i]); if (d > worst) worst = d; if (d > 1e-3) ++bad; }
    printf("B MN-major  LBO %4u  SBO %4u : max |err| %.3e, %d / 8192 wrong %s\n", c.lbo, c.sbo, worst, bad, bad == 0 ? "OK" : "");
  }
  return 0;
}
