// Microbenchmark: how fast can a [G][N][P] fp32 tensor be WRITTEN in the access pattern of the D-axis synthesis
// epilogue (tile = 128 consecutive voxels x all N rows, one thread per voxel)?
//   mode 0  direct     one STG.32 per (thread, row): a warp store covers 128 B of one row
//   mode 1  tma store  rows staged in shared memory [32 rows][128 voxels], one thread issues cp.async.bulk.tensor stores
//   mode 2  tma reduce the same with cp.reduce.async.bulk.tensor .add (accumulate epilogue without reading the old values)
//   mode 3  direct rmw out[...] += v with plain loads / stores (32 loads in flight per thread)
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o tools/ubench_store tools/ubench_store.cu -lcuda
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

struct Geo {
  int G, N, P, tiles_per_slab, total_tiles;
};

template <int MODE>
__global__ void __launch_bounds__(128) k_store(const __grid_constant__ CUtensorMap tm, float* out, Geo g) {
  __shared__ __align__(128) float stage[2][32][128];
  const int tid = threadIdx.x;
  int nbuf = 0;
  for (int tile = blockIdx.x; tile < g.total_tiles; tile += gridDim.x) {
    const int slab = tile / g.tiles_per_slab;
    const int m0 = (tile - slab * g.tiles_per_slab) * 128;
    const int m = m0 + tid;
    const bool ok = m < g.P;
    float* po = out + ((long)slab * g.N) * g.P + m;
    const float v0 = 0.001f * tid + tile;
    if (MODE == 0) {
      if (ok)
#pragma unroll 8
        for (int n = 0; n < g.N; ++n) po[(long)n * g.P] = v0 + n;
    } else if (MODE == 3) {
      if (ok)
        for (int n0 = 0; n0 < g.N; n0 += 32) {
          float old[32];
#pragma unroll
          for (int j = 0; j < 32; ++j)
            if (n0 + j < g.N) old[j] = __ldcs(po + (long)(n0 + j) * g.P);
#pragma unroll
          for (int j = 0; j < 32; ++j)
            if (n0 + j < g.N) __stcs(po + (long)(n0 + j) * g.P, old[j] + v0 + j);
        }
    } else {
      for (int n0 = 0; n0 < g.N; n0 += 32) {
        const int b = nbuf & 1;
        // the bulk store that last read this buffer (two blocks ago) must have finished READING shared memory
        if (tid == 0) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
        __syncthreads();
#pragma unroll
        for (int j = 0; j < 32; ++j) stage[b][j][tid] = v0 + n0 + j;
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncthreads();
        if (tid == 0) {
          if (MODE == 1)
            asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];" ::"l"(&tm),
                         "r"(smem_u32(&stage[b][0][0])), "r"(m0), "r"(n0), "r"(slab)
                         : "memory");
          else
            asm volatile("cp.reduce.async.bulk.tensor.3d.global.shared::cta.add.bulk_group [%0, {%2, %3, %4}], [%1];" ::"l"(&tm),
                         "r"(smem_u32(&stage[b][0][0])), "r"(m0), "r"(n0), "r"(slab)
                         : "memory");
          asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        }
        ++nbuf;
      }
    }
  }
  if (MODE == 1 || MODE == 2) {
    if (tid == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
  }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main() {
  Geo g;
  g.G = 48; g.N = 121; g.P = 9440;
  g.tiles_per_slab = (g.P + 127) / 128;
  g.total_tiles = g.tiles_per_slab * g.G;
  const size_t bytes = (size_t)g.G * g.N * g.P * 4;
  float* out;
  cudaMalloc(&out, bytes);
  cudaMemset(out, 0, bytes);
  void* fp = nullptr;
  cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fp, cudaEnableDefault, &q);
  EncodeTiledFn enc = (EncodeTiledFn)fp;
  CUtensorMap tm;
  cuuint64_t dims[3] = {(cuuint64_t)g.P, (cuuint64_t)g.N, (cuuint64_t)g.G};
  cuuint64_t strides[2] = {(cuuint64_t)g.P * 4, (cuuint64_t)g.P * g.N * 4};
  cuuint32_t box[3] = {128, 32, 1}, es[3] = {1, 1, 1};
  CUresult rc = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, out, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                    CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  printf("encode rc %d\n", (int)rc);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  const char* names[4] = {"direct STG.32   ", "TMA store       ", "TMA reduce-add  ", "direct load+add "};
  for (int mode = 0; mode < 4; ++mode)
    for (int ctas = 2; ctas <= 8; ctas *= 2) {
      float best = 1e9f;
      for (int rep = 0; rep < 4; ++rep) {
        cudaEventRecord(e0);
        const int grid = 148 * ctas;
        if (mode == 0) k_store<0><<<grid, 128>>>(tm, out, g);
        if (mode == 1) k_store<1><<<grid, 128>>>(tm, out, g);
        if (mode == 2) k_store<2><<<grid, 128>>>(tm, out, g);
        if (mode == 3) k_store<3><<<grid, 128>>>(tm, out, g);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms;
        cudaEventElapsedTime(&ms, e0, e1);
        if (rep > 0 && ms < best) best = ms;
      }
      const double traffic = (mode >= 2 ? 2.0 : 1.0) * bytes;
      printf("%s CTAs/SM %d : %.4f ms  %.0f GB/s written (%.0f GB/s incl. the read of the old values)  (%s)\n", names[mode],
             ctas, best, bytes / best * 1e-6, traffic / best * 1e-6, cudaGetErrorString(cudaGetLastError()));
    }
  // correctness of the TMA paths: out was zero, then modes ran; spot-check one element after a fresh run
  cudaMemset(out, 0, bytes);
  k_store<1><<<148, 128>>>(tm, out, g);
  k_store<2><<<148, 128>>>(tm, out, g);
  cudaDeviceSynchronize();
  float h[2];
  const long idx = ((long)5 * g.N + 77) * g.P + 130;  // slab 5, row 77, voxel 130 -> tile 5*74+1, tid 2
  cudaMemcpy(h, out + idx, 4, cudaMemcpyDeviceToHost);
  const float expect = 2.f * (0.001f * 2 + (5 * 74 + 1) + 77);
  printf("check: got %.4f expected %.4f  (%s)\n", h[0], expect, cudaGetErrorString(cudaGetLastError()));
  return 0;
}
