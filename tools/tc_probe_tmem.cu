// Probe: tcgen05.mma kind::tf32 with the A operand in TENSOR MEMORY (written by tcgen05.st from registers).
// Verifies the layout assumption "A[m][k] = TMEM lane m, column a_base + k" for M = 128, and the 3xTF32 split done in
// registers.   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o tools/tc_probe_tmem tools/tc_probe_tmem.cu
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <math.h>

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
template <int NROWS>
__device__ __host__ inline int kmajor_plain_index(int n, int k) {
  return ((k >> 3) * (NROWS / 8) * 256 + ((k & 7) >> 2) * 128 + (n >> 3) * 256 + (n & 7) * 16 + (k & 3) * 4) >> 2;
}

constexpr int K = 16, N = 32;

__global__ void __launch_bounds__(128) k_probe(const float* __restrict__ A, const float* __restrict__ bimg, float* __restrict__ out,
                                               int split) {
  __shared__ __align__(1024) float sB[2 * N * K];  // hi image, lo image
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t tmem_s;
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int i = tid; i < 2 * N * K; i += 128) sB[i] = bimg[i];
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_s)), "r"(128));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tmem_s;
  const uint32_t lane_base = tmem + ((uint32_t)(warp * 32) << 16);
  // D: cols [0, 32); A_hi: cols [32, 48); A_lo: cols [48, 64)
  uint32_t hi[K], lo[K];
#pragma unroll
  for (int k = 0; k < K; ++k) {
    const float x = A[tid * K + k];
    const float h = __uint_as_float(__float_as_uint(x) & 0xFFFFE000u);
    hi[k] = __float_as_uint(x);  // the tensor core ignores the low 13 mantissa bits
    lo[k] = __float_as_uint(x - h);
  }
#define ST16(addr, r)                                                                                                    \
  asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};" ::"r"( \
                   addr),                                                                                                \
               "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]),        \
               "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])                         \
               : "memory")
  ST16(lane_base + 32, hi);
  ST16(lane_base + 48, lo);
  asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (tid == 0) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    constexpr uint32_t idesc = make_idesc_tf32(128, N, 0, 0);
    const uint32_t bh = smem_u32(sB), bl = smem_u32(sB + N * K);
    int first = 1;
    for (int g = 0; g < K / 8; ++g) {
      const uint64_t dbh = make_desc(bh + g * (N / 8) * 256, 128, 256, 0);
      const uint64_t dbl = make_desc(bl + g * (N / 8) * 256, 128, 256, 0);
      const uint32_t a_hi = tmem + 32 + 8 * g, a_lo = tmem + 48 + 8 * g;
#define MMA(a, b, accum)                                                                                            \
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}" :: \
               "r"(tmem), "r"(a), "l"(b), "r"(idesc), "r"((uint32_t)(accum)) : "memory")
      MMA(a_hi, dbh, !first);
      first = 0;
      if (split) {
        MMA(a_lo, dbh, 1);
        MMA(a_hi, dbl, 1);
      }
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
  }
  uint32_t ok = 0;
  while (!ok)
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok) : "r"(smem_u32(&bar)), "r"(0) : "memory");
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  uint32_t r[32];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,"
      "%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(lane_base));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
  for (int n = 0; n < N; ++n) out[tid * N + n] = __uint_as_float(r[n]);
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(128) : "memory");
}

int main() {
  static float A[128 * K], B[N * K], bimg[2 * N * K], out[128 * N];
  srand(1);
  for (int i = 0; i < 128 * K; ++i) A[i] = (float)rand() / RAND_MAX * 2.f - 1.f;
  for (int i = 0; i < N * K; ++i) B[i] = (float)rand() / RAND_MAX * 2.f - 1.f;
  for (int n = 0; n < N; ++n)
    for (int k = 0; k < K; ++k) {
      const float v = B[n * K + k];
      uint32_t u;
      memcpy(&u, &v, 4);
      u = (u + 0x1000u) & 0xFFFFE000u;
      float h;
      memcpy(&h, &u, 4);
      bimg[kmajor_plain_index<N>(n, k)] = h;
      bimg[N * K + kmajor_plain_index<N>(n, k)] = v - h;
    }
  float *dA, *dB, *dO;
  cudaMalloc(&dA, sizeof(A));
  cudaMalloc(&dB, sizeof(bimg));
  cudaMalloc(&dO, sizeof(out));
  cudaMemcpy(dA, A, sizeof(A), cudaMemcpyHostToDevice);
  cudaMemcpy(dB, bimg, sizeof(bimg), cudaMemcpyHostToDevice);
  for (int split = 0; split < 2; ++split) {
    cudaMemset(dO, 0, sizeof(out));
    k_probe<<<1, 128>>>(dA, dB, dO, split);
    cudaError_t e = cudaDeviceSynchronize();
    cudaMemcpy(out, dO, sizeof(out), cudaMemcpyDeviceToHost);
    double maxerr = 0, maxref = 0;
    for (int m = 0; m < 128; ++m)
      for (int n = 0; n < N; ++n) {
        double ref = 0;
        for (int k = 0; k < K; ++k) ref += (double)A[m * K + k] * B[n * K + k];
        maxerr = fmax(maxerr, fabs(ref - out[m * N + n]));
        maxref = fmax(maxref, fabs(ref));
      }
    printf("[A in TMEM, %s] %s  max |err| %.3e  (max |ref| %.3f)\n", split ? "3xTF32" : "1xTF32", cudaGetErrorString(e), maxerr,
           maxref);
  }
  return 0;
}
