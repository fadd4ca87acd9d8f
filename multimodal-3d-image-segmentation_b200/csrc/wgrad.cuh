// Register-tiled weight-gradient accumulation shared by the pointwise-conv and stem backward kernels.
#pragma once
#include "common.cuh"

namespace hno {

constexpr int kPwThreads = 256;

template <int CO, int CI>
struct WgTile {
  static constexpr int TO = (CO % 8 == 0) ? CO / 8 : ((CO > 8 && CO % 4 == 0) ? CO / 4 : 1);
  static constexpr int NOT_RAW = CO / TO;
  static constexpr int N_OT = NOT_RAW <= 1 ? 1 : (NOT_RAW <= 2 ? 2 : (NOT_RAW <= 4 ? 4 : 8));
  static constexpr int N_IT = CI % 8 == 0 ? 8 : 4;   // column threads per group (12-channel tensors: 4 x 3 rows)
  static constexpr int TI = CI / N_IT;
  static constexpr int G = N_OT * N_IT;         // threads per group
  static constexpr int NG = kPwThreads / G;     // groups per CTA
  static_assert(CI % 4 == 0, "input channels must be a multiple of 4");
  static_assert(CO <= 8 || CO % 4 == 0, "output channels must be <= 8 or a multiple of 4");
};

// Accumulate dW[o][i] += sum_v dpre[o][v] * x[i][v] over this group's share of a TV-voxel tile.
template <int CO, int CI, int TV, int TVS>
__device__ __forceinline__ void wgrad_tile(const float* __restrict__ sdp, const float* __restrict__ sx,
                                           float2 (&accW)[WgTile<CO, CI>::TO][WgTile<CO, CI>::TI],
                                           float (&accB)[WgTile<CO, CI>::TO], bool with_bias) {
  using T = WgTile<CO, CI>;
  const int g = threadIdx.x / T::G;
  const int l = threadIdx.x - g * T::G;
  const int ot = l / T::N_IT;
  const int it = l - ot * T::N_IT;
  constexpr int TVG = TV / T::NG;
  static_assert(TVG % 4 == 0, "tile share must be a multiple of 4 voxels");
  if (ot * T::TO >= CO) return;
  const int v0 = g * TVG;
  const bool do_bias = with_bias && it == 0;
#pragma unroll 2
  for (int v = v0; v < v0 + TVG; v += 4) {
    float4 d[T::TO], x[T::TI];
#pragma unroll
    for (int q = 0; q < T::TO; ++q) d[q] = *reinterpret_cast<const float4*>(sdp + (ot * T::TO + q) * TVS + v);
#pragma unroll
    // input rows are dealt to the N_IT column threads round-robin (row r * N_IT + it): the 8 distinct rows a warp reads
    // per LDS.128 are then TVS floats apart (TVS % 32 == 4 -> 8 x 4 distinct banks) instead of TI * TVS apart
    // (4-way conflicts: 55 M excess wavefronts in the stem weight gradient)
    for (int r = 0; r < T::TI; ++r) x[r] = *reinterpret_cast<const float4*>(sx + (r * T::N_IT + it) * TVS + v);
#pragma unroll
    for (int q = 0; q < T::TO; ++q)
#pragma unroll
      for (int r = 0; r < T::TI; ++r) {
        // two packed FMAs per 4 voxels; the two lanes are summed once, in wgrad_flush
        float2 a = accW[q][r];
        a = ffma2(make_float2(d[q].x, d[q].y), make_float2(x[r].x, x[r].y), a);
        a = ffma2(make_float2(d[q].z, d[q].w), make_float2(x[r].z, x[r].w), a);
        accW[q][r] = a;
      }
    if (do_bias) {
#pragma unroll
      for (int q = 0; q < T::TO; ++q) accB[q] += (d[q].x + d[q].y) + (d[q].z + d[q].w);
    }
  }
}

// Cross-group reduction of the register tiles through shared memory, then one partial row per CTA.
template <int CO, int CI>
__device__ __forceinline__ void wgrad_flush(float* __restrict__ scratch /* >= NG*CO*CI floats */,
                                            const float2 (&accW)[WgTile<CO, CI>::TO][WgTile<CO, CI>::TI],
                                            float* __restrict__ dst /* [CO][ldw] */, int ldw, int col0) {
  using T = WgTile<CO, CI>;
  const int g = threadIdx.x / T::G;
  const int l = threadIdx.x - g * T::G;
  const int ot = l / T::N_IT;
  const int it = l - ot * T::N_IT;
  __syncthreads();
  if (ot * T::TO < CO) {
#pragma unroll
    for (int q = 0; q < T::TO; ++q)
#pragma unroll
      for (int r = 0; r < T::TI; ++r)
        scratch[g * (CO * CI) + (ot * T::TO + q) * CI + r * T::N_IT + it] = accW[q][r].x + accW[q][r].y;
  }
  __syncthreads();
  for (int idx = threadIdx.x; idx < CO * CI; idx += kPwThreads) {
    float s = 0.f;
#pragma unroll
    for (int gg = 0; gg < T::NG; ++gg) s += scratch[gg * (CO * CI) + idx];
    const int o = idx / CI, i = idx - o * CI;
    dst[o * ldw + col0 + i] = s;
  }
  __syncthreads();
}


// partials [nrows][nw+nb] -> dweight[nw], dbias[nb] (may be null), summed in fp64.
int reduce_partials(const float* partials, int nrows, int nw, int nb, float* dweight, float* dbias, int accumulate,
                    cudaStream_t st);

}  // namespace hno
