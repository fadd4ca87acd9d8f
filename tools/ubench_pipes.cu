// Micro-benchmark: issue throughput of FFMA, FFMA2 and legacy mma.sync TF32 on sm_100a (per SM per clock).
#include <cstdio>
#include <cuda_runtime.h>

__global__ void k_ffma(float* out, int iters) {
  float a[16];
  for (int i = 0; i < 16; ++i) a[i] = threadIdx.x * 0.001f + i;
  float b = out[0], c = out[1];
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 16; ++i) a[i] = fmaf(a[i], b, c);
  }
  float s = 0;
  for (int i = 0; i < 16; ++i) s += a[i];
  if (s == 1234.5f) out[2] = s;
}
__global__ void k_ffma2(float* out, int iters) {
  float2 a[16];
  for (int i = 0; i < 16; ++i) a[i] = make_float2(threadIdx.x * 0.001f + i, i);
  float2 b = make_float2(out[0], out[1]), c = make_float2(out[1], out[0]);
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 16; ++i) a[i] = __ffma2_rn(a[i], b, c);
  }
  float s = 0;
  for (int i = 0; i < 16; ++i) s += a[i].x + a[i].y;
  if (s == 1234.5f) out[2] = s;
}
__global__ void k_mma_tf32(float* out, int iters) {
  float d[8][4];
  for (int i = 0; i < 8; ++i)
    for (int j = 0; j < 4; ++j) d[i][j] = 0.f;
  unsigned a0 = __float_as_uint(out[0]), a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, b0 = __float_as_uint(out[1]), b1 = b0 + 5;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i)
      asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
                   : "+f"(d[i][0]), "+f"(d[i][1]), "+f"(d[i][2]), "+f"(d[i][3])
                   : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
  }
  float s = 0;
  for (int i = 0; i < 8; ++i) s += d[i][0] + d[i][1] + d[i][2] + d[i][3];
  if (s == 1234.5f) out[2] = s;
}
__global__ void k_mma_bf16(float* out, int iters) {
  float d[8][4];
  for (int i = 0; i < 8; ++i)
    for (int j = 0; j < 4; ++j) d[i][j] = 0.f;
  unsigned a0 = __float_as_uint(out[0]), a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, b0 = __float_as_uint(out[1]), b1 = b0 + 5;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i)
      asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
                   : "+f"(d[i][0]), "+f"(d[i][1]), "+f"(d[i][2]), "+f"(d[i][3])
                   : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
  }
  float s = 0;
  for (int i = 0; i < 8; ++i) s += d[i][0] + d[i][1] + d[i][2] + d[i][3];
  if (s == 1234.5f) out[2] = s;
}

template <typename K>
void run(const char* name, K kern, double fma_per_thread_iter, int warps_per_sm) {
  float* out;
  cudaMalloc(&out, 64);
  cudaMemset(out, 0, 64);
  int iters = 20000;
  int threads = 256, blocks = 148 * warps_per_sm / 8;
  kern<<<blocks, threads>>>(out, 100);
  cudaDeviceSynchronize();
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  cudaEventRecord(e0);
  kern<<<blocks, threads>>>(out, iters);
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  double fma = fma_per_thread_iter * iters * (double)threads * blocks;
  printf("%-12s warps/SM %2d: %8.3f ms  %8.2f TFMA/s  (%.0f FMA/clk/SM at 1.965 GHz)\n", name, warps_per_sm, ms,
         fma / ms / 1e9, fma / (ms * 1e-3) / 148 / 1.965e9);
  cudaFree(out);
}

int main() {
  for (int w : {8, 16, 32}) {
    run("ffma", k_ffma, 16, w);
    run("ffma2", k_ffma2, 32, w);
    run("mma.tf32", k_mma_tf32, 8.0 * 1024 / 32, w);
    run("mma.bf16", k_mma_bf16, 8.0 * 2048 / 32, w);
  }
  return 0;
}
