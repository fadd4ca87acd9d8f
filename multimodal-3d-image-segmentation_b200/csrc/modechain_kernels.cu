// The n_XS shared-weight frequency-domain mixes of one HNO-XS block as ONE kernel, forward and backward (sm_100a).
//
// Replaces the loop of nets/hnosegxs.py:261-262 over NeuralOperatorBlock.forward (:307-329) with
// HartleyOperator._call3d_notransform, weights_type == 'shared' (nets/hartley_operator.py:287-292):
//     z_l = selu( W_l z_{l-1} + z_{l-1} ),   l = 1..L,   W_l one C x C matrix shared by all retained modes.
// The mode tensor is tiny (B x 24 x 15,680 fp32 = 3 MB at the BASELINE config), so the per-layer kernels were pure
// launch latency and tail: 6 launches of ~20-45 us per block.  Here a thread owns one mode (all C channels in
// registers) and walks all L layers; the outputs of every layer are stored because the backward needs them.
// Backward: per layer d(pre) = dz_l * selu'(z_l), dW_l += d(pre) z_{l-1}^T (register-tiled outer products out of a
// shared-memory tile, per-CTA partial sums, fp64 final reduction by k_reduce_partials), dz_{l-1} = W_l^T d(pre) + d(pre).
#include "common.cuh"
#include "hno_b200.h"

namespace hno {

constexpr int kMcThreads = 128;
constexpr int kMcMaxLayers = 8;

struct McPtrs {
  const float* w[kMcMaxLayers];
  float* dw[kMcMaxLayers];
};

// ------------------------------------------------------------------------------------------------ forward
template <int C>
__global__ void __launch_bounds__(kMcThreads) k_modechain_fwd(const float* __restrict__ z0, float* __restrict__ zs,
                                                             const McPtrs P, int L, long M, long total) {
  // wt[l][i][o]: transposed so that one LDS.128 yields W[o..o+3][i] (broadcast across the warp)
  extern __shared__ float4 smem4[];
  float* wt = reinterpret_cast<float*>(smem4);
  for (int idx = threadIdx.x; idx < L * C * C; idx += kMcThreads) {
    const int l = idx / (C * C), r = idx - l * C * C;
    const int i = r / C, o = r - i * C;
    wt[idx] = __ldg(P.w[l] + o * C + i);
  }
  __syncthreads();
  const long t = blockIdx.x * (long)kMcThreads + threadIdx.x;
  if (t >= total) return;
  const long b = t / M, m = t - b * M;
  const long BCM = (total / M) * C * M;  // elements of one layer's output
  float x[C];
  const float* ip = z0 + b * C * M + m;
#pragma unroll
  for (int c = 0; c < C; ++c) x[c] = __ldg(ip + (long)c * M);
  for (int l = 0; l < L; ++l) {
    float2 acc[C / 2];
#pragma unroll
    for (int q = 0; q < C / 2; ++q) acc[q] = make_float2(x[2 * q], x[2 * q + 1]);  // residual
    const float4* w4 = reinterpret_cast<const float4*>(wt + l * C * C);
#pragma unroll
    for (int i = 0; i < C; ++i) {
      const float2 xi = dup2(x[i]);
#pragma unroll
      for (int q = 0; q < C / 4; ++q) {
        const float4 w = w4[i * (C / 4) + q];
        acc[2 * q] = ffma2(make_float2(w.x, w.y), xi, acc[2 * q]);
        acc[2 * q + 1] = ffma2(make_float2(w.z, w.w), xi, acc[2 * q + 1]);
      }
    }
    float* op = zs + (long)l * BCM + b * C * M + m;
#pragma unroll
    for (int q = 0; q < C / 2; ++q) {
      const float2 y = selu2(acc[q]);
      x[2 * q] = y.x;
      x[2 * q + 1] = y.y;
      op[(long)(2 * q) * M] = y.x;
      op[(long)(2 * q + 1) * M] = y.y;
    }
  }
}

// ------------------------------------------------------------------------------------------------ backward
template <int C>
__global__ void __launch_bounds__(kMcThreads) k_modechain_bwd(const float* __restrict__ dzL, const float* __restrict__ z0,
                                                             const float* __restrict__ zs, float* __restrict__ dz0,
                                                             float* __restrict__ partials, const McPtrs P, int L,
                                                             long M, long total) {
  constexpr int TV = kMcThreads;
  constexpr int TVS = TV + 4;
  extern __shared__ float4 smem4[];
  float* w = reinterpret_cast<float*>(smem4);  // [L][o][i] row-major as stored
  float* sdp = w + L * C * C;                   // [C][TVS]
  float* sx = sdp + C * TVS;                    // [C][TVS]
  for (int idx = threadIdx.x; idx < L * C * C; idx += kMcThreads) {
    const int l = idx / (C * C), r = idx - l * C * C;
    w[idx] = __ldg(P.w[l] + r);
  }
  const int tid = threadIdx.x;
  const long t = blockIdx.x * (long)kMcThreads + tid;
  const bool valid = t < total;
  const long b = valid ? t / M : 0, m = valid ? t - b * M : 0;
  const long BCM = (total / M) * C * M;
  const long off = b * C * M + m;
  float g[C];
#pragma unroll
  for (int c = 0; c < C; ++c) g[c] = valid ? __ldg(dzL + off + (long)c * M) : 0.f;
  // weight-gradient ownership: thread -> output row o = tid / NI, input columns [ic0, ic0 + TI)
  constexpr int TI = C % 6 == 0 ? 6 : 4;
  constexpr int NI = C / TI;            // threads per output row (4 for C = 24)
  static_assert(C % TI == 0 && C * NI <= kMcThreads, "weight-gradient tiling");
  const int wo = tid / NI, wi0 = (tid - wo * NI) * TI;
  const bool wactive = tid < C * NI;
  __syncthreads();
  for (int l = L - 1; l >= 0; --l) {
    const float* yp = zs + (long)l * BCM + off;                         // z_{l+1} (output of layer l)
    const float* xp = l == 0 ? z0 + off : zs + (long)(l - 1) * BCM + off;  // input of layer l
    float dp[C];
#pragma unroll
    for (int c = 0; c < C; ++c) {
      const float y = valid ? __ldg(yp + (long)c * M) : 0.f;
      dp[c] = g[c] * selu_grad_from_out(y);
      sdp[c * TVS + tid] = valid ? dp[c] : 0.f;
      sx[c * TVS + tid] = valid ? __ldg(xp + (long)c * M) : 0.f;
    }
    __syncthreads();
    // ---- weight gradient of this CTA's tile
    if (wactive) {
      float acc[TI];
#pragma unroll
      for (int r = 0; r < TI; ++r) acc[r] = 0.f;
      const float4* d4 = reinterpret_cast<const float4*>(sdp + wo * TVS);
#pragma unroll 2
      for (int v = 0; v < TV / 4; ++v) {
        const float4 d = d4[v];
#pragma unroll
        for (int r = 0; r < TI; ++r) {
          const float4 xv = *reinterpret_cast<const float4*>(sx + (wi0 + r) * TVS + 4 * v);
          acc[r] = fmaf(d.x, xv.x, acc[r]);
          acc[r] = fmaf(d.y, xv.y, acc[r]);
          acc[r] = fmaf(d.z, xv.z, acc[r]);
          acc[r] = fmaf(d.w, xv.w, acc[r]);
        }
      }
      float* pr = partials + ((long)blockIdx.x * L + l) * C * C + wo * C + wi0;
#pragma unroll
      for (int r = 0; r < TI; ++r) pr[r] = acc[r];
    }
    // ---- input gradient: g_new = W^T dp + dp (residual)
    {
      float2 a2[C / 2];
#pragma unroll
      for (int q = 0; q < C / 2; ++q) a2[q] = make_float2(dp[2 * q], dp[2 * q + 1]);
      const float4* w4 = reinterpret_cast<const float4*>(w + l * C * C);
#pragma unroll
      for (int o = 0; o < C; ++o) {
        const float2 d = dup2(dp[o]);
#pragma unroll
        for (int q = 0; q < C / 4; ++q) {
          const float4 wv = w4[o * (C / 4) + q];
          a2[2 * q] = ffma2(make_float2(wv.x, wv.y), d, a2[2 * q]);
          a2[2 * q + 1] = ffma2(make_float2(wv.z, wv.w), d, a2[2 * q + 1]);
        }
      }
#pragma unroll
      for (int q = 0; q < C / 2; ++q) {
        g[2 * q] = a2[q].x;
        g[2 * q + 1] = a2[q].y;
      }
    }
    __syncthreads();  // the staging tiles are rewritten by the next layer
  }
  if (valid) {
#pragma unroll
    for (int c = 0; c < C; ++c) dz0[off + (long)c * M] = g[c];
  }
}

// one warp per (layer, element): sum over the per-CTA partial rows
__global__ void __launch_bounds__(256) k_modechain_reduce(const float* __restrict__ partials, int nrows, int L, int n,
                                                          const McPtrs P, int accumulate) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (warp >= L * n) return;
  const int l = warp / n, e = warp - l * n;
  double s = 0.0;
  for (int r = lane; r < nrows; r += 32) s += (double)partials[((long)r * L + l) * n + e];
  s = warp_sum_d(s);
  if (lane == 0) {
    float* dst = P.dw[l] + e;
    *dst = accumulate ? *dst + (float)s : (float)s;
  }
}

// partials [nrows][L][n] -> dweights[l][n] (fp64 sums; used by the fused spectral core, spectral_core.cu)
int reduce_chain_partials(const float* partials, int nrows, int L, int n, float* const* dweights, int accumulate,
                          cudaStream_t st) {
  McPtrs P{};
  for (int l = 0; l < L; ++l) {
    HNO_CHECK(dweights[l] != nullptr, "reduce_chain_partials: null weight-gradient pointer");
    P.dw[l] = dweights[l];
  }
  k_modechain_reduce<<<ceil_div((long)L * n * 32, 256), 256, 0, st>>>(partials, nrows, L, n, P, accumulate);
  HNO_LAUNCH_CHECK();
  return 0;
}

// ------------------------------------------------------------------------------------------------ host side
int modechain_supported(int C) { return C == 24 || C == 8; }

static int check(const float* const* weights, int L, int C, long M, int B) {
  HNO_CHECK(modechain_supported(C), "modechain: channel count %d is not supported (8 or 24)", C);
  HNO_CHECK(L >= 1 && L <= kMcMaxLayers && weights, "modechain: between 1 and %d layers", kMcMaxLayers);
  HNO_CHECK(B >= 1 && M >= 1, "modechain: bad sizes");
  for (int l = 0; l < L; ++l) HNO_CHECK(weights[l] != nullptr, "modechain: null weight pointer");
  return 0;
}

int modechain_forward(const float* z0, const float* const* weights, float* zs, int B, int C, long M, int L,
                      cudaStream_t st) {
  if (check(weights, L, C, M, B)) return -1;
  HNO_CHECK(z0 && zs, "modechain_forward: null pointer");
  McPtrs P{};
  for (int l = 0; l < L; ++l) P.w[l] = weights[l];
  const long total = (long)B * M;
  const size_t smem = (size_t)L * C * C * sizeof(float);
  if (C == 24)
    k_modechain_fwd<24><<<ceil_div(total, kMcThreads), kMcThreads, smem, st>>>(z0, zs, P, L, M, total);
  else
    k_modechain_fwd<8><<<ceil_div(total, kMcThreads), kMcThreads, smem, st>>>(z0, zs, P, L, M, total);
  HNO_LAUNCH_CHECK();
  return 0;
}

size_t modechain_backward_workspace_bytes(int B, int C, long M, int L) {
  return (size_t)(ceil_div((long)B * M, kMcThreads) + 1) * L * C * C * sizeof(float);
}

int modechain_backward(const float* dzL, const float* z0, const float* zs, const float* const* weights, float* dz0,
                       float* const* dweights, void* workspace, int B, int C, long M, int L, int accumulate,
                       cudaStream_t st) {
  if (check(weights, L, C, M, B)) return -1;
  HNO_CHECK(dzL && z0 && zs && dz0 && dweights && workspace, "modechain_backward: null pointer");
  McPtrs P{};
  for (int l = 0; l < L; ++l) {
    P.w[l] = weights[l];
    HNO_CHECK(dweights[l] != nullptr, "modechain_backward: null weight-gradient pointer");
  }
  const long total = (long)B * M;
  const int grid = ceil_div(total, kMcThreads);
  float* partials = reinterpret_cast<float*>(workspace);
  const size_t smem = (size_t)(L * C * C + 2 * C * (kMcThreads + 4)) * sizeof(float);
  if (C == 24) {
    if (smem > 48 * 1024)
      HNO_CUDA(cudaFuncSetAttribute(k_modechain_bwd<24>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k_modechain_bwd<24><<<grid, kMcThreads, smem, st>>>(dzL, z0, zs, dz0, partials, P, L, M, total);
  } else {
    k_modechain_bwd<8><<<grid, kMcThreads, smem, st>>>(dzL, z0, zs, dz0, partials, P, L, M, total);
  }
  HNO_LAUNCH_CHECK();
  // partials [grid][L][C*C] -> dweights[l][C*C], summed over the CTAs in fp64 (deterministic, no atomics)
  for (int l = 0; l < L; ++l) P.dw[l] = dweights[l];
  k_modechain_reduce<<<ceil_div((long)L * C * C * 32, 256), 256, 0, st>>>(partials, grid, L, C * C, P, accumulate);
  HNO_LAUNCH_CHECK();
  return 0;
}

}  // namespace hno
