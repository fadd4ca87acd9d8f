// Microbenchmark: do tcgen05.mma instructions issued by DIFFERENT warps of one CTA overlap?  (M = 128, K = 8, tf32;
// every issuing warp has its own accumulator.)  One warp: ~102 cycles per MMA regardless of N <= 128 (ubench_mma.cu).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o tools/ubench_mma_mw tools/ubench_mma_mw.cu
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t make_desc(uint32_t addr, uint32_t lbo, uint32_t sbo, uint32_t layout) {
  uint64_t d = 0;
  d |= (uint64_t)((addr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)(layout & 7) << 61;
  return d;
}
__host__ __device__ constexpr uint32_t make_idesc_tf32(int M, int N, int a_mn, int b_mn) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)a_mn << 15) | ((uint32_t)b_mn << 16) | ((uint32_t)(N >> 3) << 17) |
         ((uint32_t)(M >> 4) << 24);
}

__global__ void __launch_bounds__(256) k_mma(int N, int nissuers, int nmma, long long* out) {
  __shared__ __align__(1024) float sA[128 * 8 * 2];
  __shared__ __align__(1024) float sB[256 * 8];
  __shared__ __align__(8) uint64_t bar[8];
  __shared__ uint32_t tmem_s;
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int i = tid; i < 128 * 16; i += 256) sA[i] = 0.001f * i;
  for (int i = tid; i < 256 * 8; i += 256) sB[i] = 0.002f * i;
  if (tid == 0) {
    for (int i = 0; i < 8; ++i) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar[i])));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_s)), "r"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tmem_s;
  if (warp < nissuers && (tid & 31) == 0) {
    const uint32_t idesc = make_idesc_tf32(128, N, 0, 0);
    const uint64_t da = make_desc(smem_u32(sA), 128, 256, 0);
    const uint64_t db = make_desc(smem_u32(sB), 128, 256, 0);
    const uint32_t d = tmem + warp * N;
    const long long t0 = clock64();
    for (int i = 0; i < nmma; ++i)
      asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::
                   "r"(d), "l"(da), "l"(db), "r"(idesc), "r"(1u) : "memory");
    const long long t1 = clock64();
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar[warp])) : "memory");
    uint32_t ok = 0;
    while (!ok)
      asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                   : "=r"(ok) : "r"(smem_u32(&bar[warp])), "r"(0) : "memory");
    const long long t2 = clock64();
    if (blockIdx.x == 0 && warp == 0) {
      out[0] = t1 - t0;
      out[1] = t2 - t0;
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512) : "memory");
}

int main() {
  long long* d;
  cudaMalloc(&d, 16);
  const int nmma = 4096;
  for (int N = 32; N <= 64; N *= 2)
    for (int nw = 1; nw <= 8; nw *= 2) {
      if (nw * N > 512) continue;
      long long h[2];
      for (int rep = 0; rep < 2; ++rep) {
        k_mma<<<148, 256>>>(N, nw, nmma, d);
        cudaDeviceSynchronize();
      }
      cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
      printf("1 CTA/SM, %d issuing warps, N %3d : issue %.1f cyc/MMA/warp, complete %.1f cyc/MMA/warp -> %.1f cyc/MMA per SM  (%s)\n",
             nw, N, (double)h[0] / nmma, (double)h[1] / nmma, (double)h[1] / nmma / nw, cudaGetErrorString(cudaGetLastError()));
    }
  return 0;
}
