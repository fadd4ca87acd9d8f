// Microbenchmark: HBM read bandwidth of the channel-planar access pattern of the streamed kernels.
// A [G][C][S] fp32 tensor is read in tiles: a CTA takes tile t (SEG contiguous bytes of every one of the C rows of a slab),
// tiles are dealt round-robin (tile = blockIdx.x + i * gridDim.x).  SEG = 512 B is what k_tc_stream does today.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o tools/ubench_strided tools/ubench_strided.cu && tools/ubench_strided
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>

template <int UNROLL>
__global__ void __launch_bounds__(256) k_read(const float4* __restrict__ a, long S4, int C, int seg4, long tiles_per_slab,
                                              long total_tiles, int blocked, float* sink) {
  // S4: row length in float4; seg4: float4 per segment
  const int tpr = seg4;                 // threads per row segment
  const int rows_per_pass = 256 / tpr;  // rows covered by one pass of the CTA
  const int col = threadIdx.x % tpr, r0 = threadIdx.x / tpr;
  float acc = 0.f;
  const long my = total_tiles > blockIdx.x ? (total_tiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
  for (long i = 0; i < my; ++i) {
    const long tile = blocked ? (long)blockIdx.x * ((total_tiles + gridDim.x - 1) / gridDim.x) + i : blockIdx.x + i * gridDim.x;
    if (tile >= total_tiles) break;
    const long g = tile / tiles_per_slab, t = tile - g * tiles_per_slab;
    const float4* base = a + g * C * S4 + t * seg4 + col;
    for (int r = r0; r < C; r += rows_per_pass * UNROLL) {
      float4 v[UNROLL];
#pragma unroll
      for (int u = 0; u < UNROLL; ++u) {
        const int rr = r + u * rows_per_pass;
        v[u] = rr < C ? __ldcs(base + (long)rr * S4) : make_float4(0, 0, 0, 0);
      }
#pragma unroll
      for (int u = 0; u < UNROLL; ++u) acc += v[u].x + v[u].y + v[u].z + v[u].w;
    }
  }
  if (acc == 12345.678f) *sink = acc;
}

int main() {
  const int G = 2, C = 48;
  const long S = 121L * 9440;  // floats per row (BASELINE grid, padded planes)
  const long n = (long)G * C * S;
  float* a;
  float* sink;
  cudaMalloc(&a, n * 4);
  cudaMalloc(&sink, 4);
  cudaMemset(a, 0, n * 4);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  printf("bytes per pass: %.1f MB\n", n * 4 / 1e6);
  for (int blocked = 0; blocked < 2; ++blocked)
    for (int per_sm = 2; per_sm <= 8; per_sm *= 2)
      for (int seg = 128; seg <= 4096; seg *= 2) {
        const int seg4 = seg / 16;
        if (seg4 > 256) continue;
        const long tps = (S * 4 + seg - 1) / seg;  // last partial tile reads past the row: keep it simple, drop it
        const long tiles_per_slab = S * 4 / seg;
        const long total = tiles_per_slab * G;
        (void)tps;
        const int grid = 148 * per_sm;
        for (int rep = 0; rep < 2; ++rep) {
          cudaEventRecord(e0);
          k_read<4><<<grid, 256>>>((const float4*)a, S / 4, C, seg4, tiles_per_slab, total, blocked, sink);
          cudaEventRecord(e1);
          cudaEventSynchronize(e1);
        }
        float ms;
        cudaEventElapsedTime(&ms, e0, e1);
        printf("%s  CTAs/SM %d  SEG %4d B : %.3f ms  %.0f GB/s\n", blocked ? "blocked    " : "round-robin", per_sm, seg, ms,
               n * 4 / ms / 1e6);
      }
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
