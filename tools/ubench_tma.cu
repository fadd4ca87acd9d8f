// Microbenchmark: HBM read bandwidth of TMA box loads in the access pattern of k_tc_stream (no compute):
// persistent CTAs, one thread keeps NST boxes of [ROWS][32 floats] x NB column blocks in flight per CTA.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o tools/ubench_tma tools/ubench_tma.cu && tools/ubench_tma
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok = 0;
  while (!ok)
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, 20000;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma_load_3d(void* dst, const CUtensorMap* tm, int c0, int c1, int c2, uint64_t* bar) {
  asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];" ::
               "r"(smem_u32(dst)), "l"(tm), "r"(c0), "r"(c1), "r"(c2), "r"(smem_u32(bar)) : "memory");
}

// work item = (tile of NB*32 columns, chunk of ROWS rows); items dealt round-robin over CTAs by tile
__global__ void __launch_bounds__(64) k_tma(const __grid_constant__ CUtensorMap tm, int nst, int rows, int nb, int boxw,
                                            int chunks_per_tile, int tiles_per_slab, int total_tiles, float* sink) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ __align__(8) uint64_t full[64];
  const int stage_bytes = rows * boxw * 4 * nb;
  if (threadIdx.x == 0) {
    for (int s = 0; s < nst; ++s) mbar_init(&full[s], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (threadIdx.x != 0) return;
  const int my_tiles = total_tiles > (int)blockIdx.x ? (total_tiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
  const long items = (long)my_tiles * chunks_per_tile;
  float acc = 0.f;
  auto issue = [&](long it) {
    const int ti = it / chunks_per_tile, c = it - (long)ti * chunks_per_tile;
    const int tile = blockIdx.x + ti * gridDim.x;
    const int g = tile / tiles_per_slab, m0 = (tile - g * tiles_per_slab) * nb * boxw;
    const int s = it % nst;
    mbar_expect_tx(&full[s], stage_bytes);
    for (int j = 0; j < nb; ++j) tma_load_3d(smem + s * stage_bytes + j * rows * boxw * 4, &tm, m0 + boxw * j, c * rows, g, &full[s]);
  };
  for (long it = 0; it < nst && it < items; ++it) issue(it);
  for (long it = 0; it < items; ++it) {
    const int s = it % nst;
    mbar_wait(&full[s], (uint32_t)((it / nst) & 1));
    acc += *reinterpret_cast<volatile float*>(smem + s * stage_bytes);
    if (it + nst < items) issue(it + nst);
  }
  if (acc == 12345.678f) *sink = acc;
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main() {
  const int G = 2, C = 48;
  const long S = 121L * 9440;
  const long n = (long)G * C * S;
  float *a, *sink;
  cudaMalloc(&a, n * 4);
  cudaMalloc(&sink, 4);
  cudaMemset(a, 0, n * 4);
  void* fp = nullptr;
  cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fp, cudaEnableDefault, &q);
  EncodeTiledFn enc = (EncodeTiledFn)fp;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  cudaFuncSetAttribute(k_tma, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
  struct Cfg { int per_sm, nst, rows, nb, boxw, swz, promo; };
  const Cfg cfgs[] = {
      {3, 4, 8, 4, 32, 2, 3},
      {3, 4, 16, 4, 32, 2, 3},
      {3, 4, 24, 4, 32, 2, 3},
      {3, 3, 48, 4, 32, 2, 3},
      {2, 4, 48, 4, 32, 2, 3},
      {4, 2, 48, 4, 32, 2, 3},
      {3, 4, 24, 1, 128, 0, 3},
      {3, 4, 48, 1, 128, 0, 3},
      {3, 4, 24, 2, 64, 0, 3},
      {3, 8, 24, 8, 16, 0, 3},
      {3, 4, 24, 2, 32, 2, 3},
      {3, 4, 24, 8, 32, 2, 3},
      {6, 4, 24, 2, 32, 2, 3},
      {6, 2, 24, 4, 32, 2, 3},
      {8, 2, 24, 4, 32, 2, 3},
  };
  for (const Cfg& c : cfgs) {
    CUtensorMap tm;
    cuuint64_t dims[3] = {(cuuint64_t)S, (cuuint64_t)C, (cuuint64_t)G};
    cuuint64_t strides[2] = {(cuuint64_t)S * 4, (cuuint64_t)C * S * 4};
    cuuint32_t box[3] = {(cuuint32_t)c.boxw, (cuuint32_t)c.rows, 1}, es[3] = {1, 1, 1};
    CUtensorMapSwizzle sw = c.swz == 2 ? CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B : c.swz == 1 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_NONE;
    CUtensorMapL2promotion pr = c.promo == 0 ? CU_TENSOR_MAP_L2_PROMOTION_NONE : c.promo == 1 ? CU_TENSOR_MAP_L2_PROMOTION_L2_64B
                                : c.promo == 2 ? CU_TENSOR_MAP_L2_PROMOTION_L2_128B : CU_TENSOR_MAP_L2_PROMOTION_L2_256B;
    CUresult rc = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, a, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, sw, pr,
                      CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (rc != CUDA_SUCCESS) { printf("encode failed %d\n", (int)rc); continue; }
    const int tile_cols = c.nb * c.boxw;
    const int tiles_per_slab = (int)(S / tile_cols);
    const int total = tiles_per_slab * G;
    const int chunks = C / c.rows;
    const size_t smem = (size_t)c.nst * c.rows * c.boxw * 4 * c.nb + 1024;
    float ms = 0;
    for (int rep = 0; rep < 3; ++rep) {
      cudaEventRecord(e0);
      k_tma<<<148 * c.per_sm, 64, smem>>>(tm, c.nst, c.rows, c.nb, c.boxw, chunks, tiles_per_slab, total, sink);
      cudaEventRecord(e1);
      cudaEventSynchronize(e1);
      cudaEventElapsedTime(&ms, e0, e1);
    }
    const double bytes = (double)total * tile_cols * C * 4;
    printf("CTAs/SM %d  stages %2d  box %3d floats x %2d rows x %d  swz %d promo %d  in-flight/SM %4zu KB : %.3f ms  %.0f GB/s  (%s)\n",
           c.per_sm, c.nst, c.boxw, c.rows, c.nb, c.swz, c.promo, (size_t)c.per_sm * c.nst * c.rows * c.boxw * 4 * c.nb / 1024, ms,
           bytes / ms / 1e6, cudaGetErrorString(cudaGetLastError()));
  }
  return 0;
}
