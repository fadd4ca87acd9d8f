// Pointwise (1x1x1) channel mixing on channel-planar tensors, forward and backward, sm_100a.
//
// Replaces (see include/hno_b200.h for the full list):
//   nets/nets_utils.py:120-174      ConvNormAct(kernel_size=1) = Conv3d 1x1x1 + bias + SELU
//   nets/hartley_operator.py:287-292 einsum('oi,bidhw->bodhw') on the cropped modes
//   nets/hnosegxs.py:307-329        NeuralOperatorBlock: selu(op(x) + x)
//   nets/hnosegxs.py:254-255,273-275 mapping conv / concat conv on torch.cat([a, b], dim=1)
// The concat is virtual: the two halves arrive as two pointers, torch.cat never materialises.
//
// Forward: one thread owns V adjacent voxels and all CO outputs (register accumulators); the
// weights are broadcast from shared memory.  Backward: a persistent CTA walks voxel tiles; phase A
// computes d(pre-activation) and the input gradients per voxel, phase B accumulates the weight
// gradient as register-tiled outer products out of shared memory.  Per-CTA partial sums are
// written to a workspace and reduced in fp64 by a second tiny kernel (deterministic, no atomics).
#include "common.cuh"
#include "hno_b200.h"
#include "wgrad.cuh"
#include "tc_stream.h"

namespace hno {

// pipelined tensor-core variant for the 24(+24) -> 24 SELU layers (pwconv_bwd_tc.cu)
bool pwconv_bwd_tc_eligible(const float* dy, const float* y, const float* in1, const float* in2, const float* din1,
                            const float* din2, int ci1, int ci2, int co, long S, int act, int residual);
int pwconv_bwd_tc(const float* dy, const float* y, const float* in1, const float* in2, const float* w, float* din1,
                  float* din2, float* dweight, float* dbias, void* ws, int B, int ci2, long S, long P, long HW, int flags,
                  cudaStream_t st);

template <int ACT>
__device__ __forceinline__ float act_f(float x) {
  return ACT == 1 ? selu_f(x) : x;
}
template <int ACT>
__device__ __forceinline__ float act_grad_from_out(float y) {
  return ACT == 1 ? selu_grad_from_out(y) : 1.f;
}

// ------------------------------------------------------------------------------------------ forward
constexpr int kFwdRows = 8;    // input-channel rows per pipeline stage
constexpr int kFwdStages = 4;  // stages of the per-thread cp.async ring

// Persistent CTAs; every thread owns V adjacent voxels of a tile and streams the CI input rows of that tile
// through a PRIVATE 4-stage cp.async ring in shared memory (it only ever reads back the bytes it copied itself,
// so the pipeline needs no barriers at all).  The ring keeps 3 x 8 rows x V x 4 B per thread in flight
// (96 KB per SM at 2 CTAs), and it runs across tile boundaries, so HBM latency is paid once per kernel.
template <int CI1, int CI2, int CO, int ACT, bool RES, int V>
__global__ void __launch_bounds__(kPwThreads, 2) k_pwconv_fwd(const float* __restrict__ in1,
                                                              const float* __restrict__ in2,
                                                              const float* __restrict__ weight,
                                                              const float* __restrict__ bias, float* __restrict__ out,
                                                              long S, int tiles_per_sample, long total_tiles) {
  static_assert(!RES || (CI1 == CO && CI2 == 0), "residual needs CI1 == CO and a single input");
  constexpr int CI = CI1 + CI2;
  constexpr int FR = (CI1 % kFwdRows == 0 && CI2 % kFwdRows == 0) ? kFwdRows : 4;  // rows per stage (12-channel tensors: 4)
  static_assert(CI1 % FR == 0 && CI2 % FR == 0, "input channels must be a multiple of 4");
  constexpr int NCH = CI / FR;
  constexpr int COp = (CO + 3) & ~3;
  constexpr int TV = kPwThreads * V;
  extern __shared__ float4 smem4[];
  float* wt = reinterpret_cast<float*>(smem4);      // [CI][COp] transposed weights
  float* sbias = wt + CI * COp;                      // [COp]
  float* ring = sbias + COp;                         // [kFwdStages][FR][TV]
  for (int idx = threadIdx.x; idx < CI * COp; idx += kPwThreads) {
    int i = idx / COp, o = idx - i * COp;
    wt[idx] = o < CO ? weight[o * CI + i] : 0.f;
  }
  for (int o = threadIdx.x; o < COp; o += kPwThreads) sbias[o] = (bias != nullptr && o < CO) ? bias[o] : 0.f;
  __syncthreads();

  const int lv = threadIdx.x * V;
  const long n_my = total_tiles > blockIdx.x ? (total_tiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
  const long G = n_my * NCH;

  auto issue = [&](long g) {
    if (g < G) {
      const long k = g / NCH;
      const int ch = (int)(g - k * NCH);
      const long tile = blockIdx.x + k * gridDim.x;
      const int b = (int)(tile / tiles_per_sample);
      const long s0 = (tile - (long)b * tiles_per_sample) * TV + lv;
      if (s0 < S) {
        float* dst = ring + ((int)(g % kFwdStages) * FR) * TV + lv;
#pragma unroll
        for (int r = 0; r < FR; ++r) {
          const int c = ch * FR + r;
          const float* src = (CI2 == 0 || c < CI1) ? in1 + ((long)b * CI1 + c) * S + s0
                                                   : in2 + ((long)b * CI2 + (c - CI1)) * S + s0;
          cp_async_vec<V>(dst + r * TV, src);
        }
      }
    }
    cp_async_commit();
  };

#pragma unroll
  for (int g = 0; g < kFwdStages - 1; ++g) issue(g);

  // acc[op][v] = (out[2 op][v], out[2 op + 1][v]): output channels are paired so that the weight pair comes
  // straight out of one LDS.128 and each FFMA2 retires two FMAs
  float2 acc[COp / 2][V];
  for (long g = 0; g < G; ++g) {
    issue(g + kFwdStages - 1);
    cp_async_wait_group<kFwdStages - 1>();
    const long k = g / NCH;
    const int ch = (int)(g - k * NCH);
    if (ch == 0) {
#pragma unroll
      for (int op = 0; op < COp / 2; ++op)
#pragma unroll
        for (int v = 0; v < V; ++v) acc[op][v] = make_float2(sbias[2 * op], sbias[2 * op + 1]);
    }
    const float* src = ring + ((int)(g % kFwdStages) * FR) * TV + lv;
#pragma unroll
    for (int r = 0; r < FR; ++r) {
      float2 xd[V];
      if (V == 4) {
        const float4 t = *reinterpret_cast<const float4*>(src + r * TV);
        xd[0] = dup2(t.x); xd[1 % V] = dup2(t.y); xd[2 % V] = dup2(t.z); xd[3 % V] = dup2(t.w);
      } else if (V == 2) {
        const float2 t = *reinterpret_cast<const float2*>(src + r * TV);
        xd[0] = dup2(t.x); xd[1 % V] = dup2(t.y);
      } else {
        xd[0] = dup2(src[r * TV]);
      }
      const float4* w4 = reinterpret_cast<const float4*>(wt + (ch * FR + r) * COp);
#pragma unroll
      for (int q = 0; q < COp / 4; ++q) {
        const float4 w = w4[q];
#pragma unroll
        for (int v = 0; v < V; ++v) {
          acc[2 * q + 0][v] = ffma2(make_float2(w.x, w.y), xd[v], acc[2 * q + 0][v]);
          acc[2 * q + 1][v] = ffma2(make_float2(w.z, w.w), xd[v], acc[2 * q + 1][v]);
        }
      }
    }
    if (ch == NCH - 1) {
      const long tile = blockIdx.x + k * gridDim.x;
      const int b = (int)(tile / tiles_per_sample);
      const long s0 = (tile - (long)b * tiles_per_sample) * TV + lv;
      if (s0 < S) {
        float* po = out + (long)b * CO * S + s0;
#pragma unroll
        for (int o = 0; o < CO; ++o) {
          Vec<V> r;
#pragma unroll
          for (int v = 0; v < V; ++v) r.v[v] = (o & 1) ? acc[o / 2][v].y : acc[o / 2][v].x;
          if (RES) {
            Vec<V> x = Vec<V>::ld(in1 + ((long)b * CI1 + o) * S + s0);  // just streamed: L2 hit
#pragma unroll
            for (int v = 0; v < V; ++v) r.v[v] += x.v[v];
          }
#pragma unroll
          for (int v = 0; v < V; ++v) r.v[v] = act_f<ACT>(r.v[v]);
          r.st(po + (long)o * S);
        }
      }
    }
  }
}

// ------------------------------------------------------------------------------------------ backward

// global [C][S] columns (lv .. lv+VV) of this thread -> shared [C][TVS], asynchronously (LDGSTS)
template <int C, int VV, int TVS>
__device__ __forceinline__ void stage_input(float* __restrict__ sx, const float* __restrict__ src, long S, int lv,
                                            bool valid) {
  if (valid) {
    const unsigned dst = (unsigned)__cvta_generic_to_shared(sx + lv);
#pragma unroll
    for (int i = 0; i < C; ++i) {
      if (VV == 2)
        asm volatile("cp.async.ca.shared.global [%0], [%1], 8;\n" ::"r"(dst + i * TVS * 4), "l"(src + (long)i * S)
                     : "memory");
      else
        asm volatile("cp.async.ca.shared.global [%0], [%1], 4;\n" ::"r"(dst + i * TVS * 4), "l"(src + (long)i * S)
                     : "memory");
    }
  } else {
#pragma unroll
    for (int i = 0; i < C; ++i)
#pragma unroll
      for (int v = 0; v < VV; ++v) sx[i * TVS + lv + v] = 0.f;
  }
}

template <int CI1, int CI2, int CO, int ACT, bool RES, int VV>
__global__ void __launch_bounds__(kPwThreads, 2) k_pwconv_bwd(
    const float* __restrict__ dy, const float* __restrict__ y, const float* __restrict__ in1,
    const float* __restrict__ in2, const float* __restrict__ weight, float* __restrict__ din1,
    float* __restrict__ din2, float* __restrict__ partials, long S, long P, long HW, int tiles_per_sample,
    long total_tiles, int flags) {
  constexpr int CI = CI1 + CI2;
  constexpr int CIM = CI1 > CI2 ? CI1 : CI2;
  constexpr int COp = (CO + 3) & ~3;
  constexpr int TV = kPwThreads * VV;
  constexpr int TVS = TV + 4;  // (TVS/4) odd -> float4 rows of different channels land on different bank groups
  constexpr int PSTRIDE = CO * CI + CO;
  extern __shared__ float4 smem4[];
  float* wt = reinterpret_cast<float*>(smem4);  // [CI][COp]
  float* sdp = wt + CI * COp;                   // [CO][TVS]
  float* sx = sdp + CO * TVS;                   // [CIM][TVS]
  for (int idx = threadIdx.x; idx < CI * COp; idx += kPwThreads) {
    int i = idx / COp, o = idx - i * COp;
    wt[idx] = o < CO ? weight[o * CI + i] : 0.f;
  }
  using T1 = WgTile<CO, CI1>;
  float2 accW1[T1::TO][T1::TI];
  float accB[T1::TO];
#pragma unroll
  for (int q = 0; q < T1::TO; ++q) {
    accB[q] = 0.f;
#pragma unroll
    for (int r = 0; r < T1::TI; ++r) accW1[q][r] = make_float2(0.f, 0.f);
  }
  constexpr int CI2s = CI2 > 0 ? CI2 : 8;
  using T2 = WgTile<CO, CI2s>;
  float2 accW2[T2::TO][T2::TI];
  float accB2[T2::TO];
#pragma unroll
  for (int q = 0; q < T2::TO; ++q) {
    accB2[q] = 0.f;
#pragma unroll
    for (int r = 0; r < T2::TI; ++r) accW2[q][r] = make_float2(0.f, 0.f);
  }
  __syncthreads();

  const int lv = threadIdx.x * VV;
  for (long tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
    const int b = (int)(tile / tiles_per_sample);
    const long s0 = (tile - (long)b * tiles_per_sample) * TV + lv;
    const bool valid = s0 < S;
    bool live[VV];
#pragma unroll
    for (int v = 0; v < VV; ++v) live[v] = valid && ((s0 + v) % P) < HW;

    // input 1 goes global -> shared with cp.async (no registers, all CI1 requests in flight at once) while the
    // d(pre-activation) loads below are outstanding too: one exposed memory latency per tile instead of one per channel
    stage_input<CI1, VV, TVS>(sx, in1 + (long)b * CI1 * S + s0, S, lv, valid);

    // ---- phase A1: d(pre-activation), staged to shared memory; kept in registers as output-channel pairs
    pk2 dp2[COp / 2][VV];  // output-channel pairs (2q, 2q+1), packed (see pack2 in common.cuh)
#pragma unroll
    for (int op = 0; op < COp / 2; ++op) {
      float d[2][VV];
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int o = 2 * op + h;
        if (o < CO && valid) {
          Vec<VV> g = Vec<VV>::ld(dy + ((long)b * CO + o) * S + s0);
          Vec<VV> yy;
          if (ACT != 0) yy = Vec<VV>::ld(y + ((long)b * CO + o) * S + s0);
#pragma unroll
          for (int v = 0; v < VV; ++v)
            d[h][v] = live[v] ? g.v[v] * (ACT != 0 ? act_grad_from_out<ACT>(yy.v[v]) : 1.f) : 0.f;
        } else {
#pragma unroll
          for (int v = 0; v < VV; ++v) d[h][v] = 0.f;
        }
        if (o < CO) {
#pragma unroll
          for (int v = 0; v < VV; ++v) sdp[o * TVS + lv + v] = d[h][v];
        }
      }
#pragma unroll
      for (int v = 0; v < VV; ++v) dp2[op][v] = pack2(d[0][v], d[1][v]);
    }
    cp_async_wait_all();
    // ---- input gradient of input 1 (own columns of sx / sdp only: no barrier needed yet)
    if (din1 != nullptr && valid) {
      float* dst = din1 + (long)b * CI1 * S + s0;
#pragma unroll 4
      for (int i = 0; i < CI1; ++i) {
        pk2 a2[VV];
#pragma unroll
        for (int v = 0; v < VV; ++v) a2[v] = pack2(RES ? sdp[i * TVS + lv + v] : 0.f, 0.f);
        const ulonglong2* w4 = reinterpret_cast<const ulonglong2*>(wt + i * COp);
#pragma unroll
        for (int q = 0; q < COp / 4; ++q) {
          const ulonglong2 w = w4[q];
#pragma unroll
          for (int v = 0; v < VV; ++v) {
            a2[v] = ffma2p(w.x, dp2[2 * q + 0][v], a2[v]);
            a2[v] = ffma2p(w.y, dp2[2 * q + 1][v], a2[v]);
          }
        }
        Vec<VV> r;
#pragma unroll
        for (int v = 0; v < VV; ++v) {
          const float2 t = unpack2(a2[v]);
          r.v[v] = t.x + t.y;
        }
        if (flags & 4) {
#pragma unroll
          for (int v = 0; v < VV; ++v) r.v[v] *= selu_grad_from_out(sx[i * TVS + lv + v]);
        }
        if (flags & 1) {
          Vec<VV> old = Vec<VV>::ld(dst + (long)i * S);
#pragma unroll
          for (int v = 0; v < VV; ++v) r.v[v] += old.v[v];
        }
        r.st(dst + (long)i * S);
      }
    }
    __syncthreads();
    wgrad_tile<CO, CI1, TV, TVS>(sdp, sx, accW1, accB, true);
    if (CI2 > 0) {
      __syncthreads();  // everyone is done reading input 1 from sx
      stage_input<CI2s, VV, TVS>(sx, in2 + (long)b * CI2 * S + s0, S, lv, valid);
      if (din2 != nullptr && valid) {
        float* dst = din2 + (long)b * CI2 * S + s0;
#pragma unroll 4
        for (int i = 0; i < CI2; ++i) {
          pk2 a2[VV];
#pragma unroll
          for (int v = 0; v < VV; ++v) a2[v] = 0ull;
          const ulonglong2* w4 = reinterpret_cast<const ulonglong2*>(wt + (CI1 + i) * COp);
#pragma unroll
          for (int q = 0; q < COp / 4; ++q) {
            const ulonglong2 w = w4[q];
#pragma unroll
            for (int v = 0; v < VV; ++v) {
              a2[v] = ffma2p(w.x, dp2[2 * q + 0][v], a2[v]);
              a2[v] = ffma2p(w.y, dp2[2 * q + 1][v], a2[v]);
            }
          }
          Vec<VV> r;
#pragma unroll
          for (int v = 0; v < VV; ++v) {
            const float2 t = unpack2(a2[v]);
            r.v[v] = t.x + t.y;
          }
          if (flags & 2) {
            Vec<VV> old = Vec<VV>::ld(dst + (long)i * S);
#pragma unroll
            for (int v = 0; v < VV; ++v) r.v[v] += old.v[v];
          }
          r.st(dst + (long)i * S);
        }
      }
      cp_async_wait_all();  // the input-2 copy overlapped the input-gradient math above
      __syncthreads();
      wgrad_tile<CO, CI2s, TV, TVS>(sdp, sx, accW2, accB2, false);
    }
    __syncthreads();  // tiles are overwritten by the next iteration
  }

  // ---- per-CTA partial sums -> workspace row
  float* prow = partials + (long)blockIdx.x * PSTRIDE;
  wgrad_flush<CO, CI1>(sdp, accW1, prow, CI, 0);
  if (CI2 > 0) wgrad_flush<CO, CI2s>(sdp, accW2, prow, CI, CI1);
  {  // bias: threads with it == 0 hold partial sums per group
    const int g = threadIdx.x / T1::G;
    const int l = threadIdx.x - g * T1::G;
    const int ot = l / T1::N_IT;
    const int it = l - ot * T1::N_IT;
    if (it == 0 && ot * T1::TO < CO) {
#pragma unroll
      for (int q = 0; q < T1::TO; ++q) sx[g * CO + ot * T1::TO + q] = accB[q];
    }
    __syncthreads();
    for (int o = threadIdx.x; o < CO; o += kPwThreads) {
      float s = 0.f;
#pragma unroll
      for (int gg = 0; gg < T1::NG; ++gg) s += sx[gg * CO + o];
      prow[CO * CI + o] = s;
    }
  }
}

// partials [nrows][n] -> dst[n] in fp64; flags bit3 accumulates.  dbias may be null.
__global__ void __launch_bounds__(256) k_reduce_partials(const float* __restrict__ partials, int nrows, int n,
                                                         int nw, float* __restrict__ dweight,
                                                         float* __restrict__ dbias, int accumulate) {
  // one warp per output element
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");  // a streamed tensor-core kernel usually follows
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (warp >= n) return;
  double s = 0.0;
  for (int r = lane; r < nrows; r += 32) s += (double)partials[(long)r * n + warp];
  s = warp_sum_d(s);
  if (lane == 0) {
    float* dst = warp < nw ? dweight + warp : (dbias ? dbias + (warp - nw) : nullptr);
    if (dst) *dst = accumulate ? *dst + (float)s : (float)s;
  }
}

int reduce_partials(const float* partials, int nrows, int nw, int nb, float* dweight, float* dbias, int accumulate,
                    cudaStream_t st) {
  const int n = nw + nb;
  k_reduce_partials<<<ceil_div((long)n * 32, 256), 256, 0, st>>>(partials, nrows, n, nw, dweight, dbias, accumulate);
  HNO_LAUNCH_CHECK();
  return 0;
}

// ------------------------------------------------------------------------------------------ dispatch
static int bwd_grid(long total_tiles) {
  long g = (long)sm_count() * 2;
  return (int)(total_tiles < g ? total_tiles : g);
}

template <int CI1, int CI2, int CO, int V>
static size_t fwd_smem() {
  constexpr int COp = (CO + 3) & ~3;
  return (size_t)((CI1 + CI2) * COp + COp + kFwdStages * kFwdRows * kPwThreads * V) * sizeof(float);
}

template <int CI1, int CI2, int CO, int ACT, bool RES, int V>
static int fwd_launch(const float* in1, const float* in2, const float* w, const float* bias, float* out, int B, long S,
                      cudaStream_t st) {
  constexpr int TV = kPwThreads * V;
  const int tps = ceil_div(S, TV);
  const long total = (long)tps * B;
  const long gmax = (long)sm_count() * 2;
  const int grid = (int)(total < gmax ? total : gmax);
  auto kern = k_pwconv_fwd<CI1, CI2, CO, ACT, RES, V>;
  const size_t smem = fwd_smem<CI1, CI2, CO, V>();
  HNO_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  kern<<<grid, kPwThreads, smem, st>>>(in1, in2, w, bias, out, S, tps, total);
  HNO_LAUNCH_CHECK();
  return 0;
}

template <int CI1, int CI2, int CO, int ACT, bool RES>
static int fwd_t(const float* in1, const float* in2, const float* w, const float* bias, float* out, int B, long S,
                 cudaStream_t st) {
  const void* ptrs[3] = {in1, in2, out};
  const long cnt[1] = {S};
  int v = pick_vec(ptrs, 3, cnt, 1);
  // 2 voxels per thread keep the CO/2 x V packed accumulators + operands at ~110 registers (2 CTAs / SM)
  if (v == 4) v = 2;
  if (v == 2) return fwd_launch<CI1, CI2, CO, ACT, RES, 2>(in1, in2, w, bias, out, B, S, st);
  return fwd_launch<CI1, CI2, CO, ACT, RES, 1>(in1, in2, w, bias, out, B, S, st);
}

template <int CI1, int CI2, int CO, int VV>
static size_t bwd_smem() {
  constexpr int CI = CI1 + CI2;
  constexpr int CIM = CI1 > CI2 ? CI1 : CI2;
  constexpr int COp = (CO + 3) & ~3;
  constexpr int TVS = kPwThreads * VV + 4;
  return (size_t)(CI * COp + (CO + CIM) * TVS) * sizeof(float);
}

template <int CI1, int CI2, int CO, int ACT, bool RES>
static int bwd_t(const float* dy, const float* y, const float* in1, const float* in2, const float* w, float* din1,
                 float* din2, float* dweight, float* dbias, void* ws, int B, long S, long P, long HW, int flags,
                 cudaStream_t st) {
  constexpr int CI = CI1 + CI2;
  const void* ptrs[6] = {dy, y, in1, in2, din1, din2};
  const long cnt[1] = {S};
  const int v = pick_vec(ptrs, 6, cnt, 1);
  float* partials = reinterpret_cast<float*>(ws);
  int grid;
  if (v >= 2) {
    constexpr int TV = kPwThreads * 2;
    const int tps = ceil_div(S, TV);
    const long total = (long)tps * B;
    grid = bwd_grid(total);
    auto kern = k_pwconv_bwd<CI1, CI2, CO, ACT, RES, 2>;
    const size_t smem = bwd_smem<CI1, CI2, CO, 2>();
    HNO_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kern<<<grid, kPwThreads, smem, st>>>(dy, y, in1, in2, w, din1, din2, partials, S, P, HW, tps, total, flags);
  } else {
    constexpr int TV = kPwThreads;
    const int tps = ceil_div(S, TV);
    const long total = (long)tps * B;
    grid = bwd_grid(total);
    auto kern = k_pwconv_bwd<CI1, CI2, CO, ACT, RES, 1>;
    const size_t smem = bwd_smem<CI1, CI2, CO, 1>();
    HNO_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kern<<<grid, kPwThreads, smem, st>>>(dy, y, in1, in2, w, din1, din2, partials, S, P, HW, tps, total, flags);
  }
  HNO_LAUNCH_CHECK();
  return reduce_partials(partials, grid, CO * CI, CO, dweight, dbias, (flags & 8) ? 1 : 0, st);
}

// (CI1, CI2, CO, ACT, RES) combinations the model needs, for filters F in {8, 12, 24} and heads of 2..4 classes; the 4 -> F
// rows are conv1 of the use_resize=False models reading the (4-channel, or zero-padded to 4) image directly.
#define HNO_PW_CONFIGS(X) \
  X(4, 0, 8, 1, false)    \
  X(4, 0, 12, 1, false)   \
  X(4, 0, 24, 1, false)   \
  X(8, 0, 8, 1, true)     \
  X(8, 0, 8, 1, false)    \
  X(8, 8, 8, 1, false)    \
  X(8, 8, 8, 0, false)    \
  X(8, 0, 8, 0, false)    \
  X(8, 0, 2, 0, false)    \
  X(8, 0, 3, 0, false)    \
  X(8, 0, 4, 0, false)    \
  X(12, 0, 12, 1, true)   \
  X(12, 0, 12, 1, false)  \
  X(12, 12, 12, 1, false) \
  X(12, 12, 12, 0, false) \
  X(12, 0, 12, 0, false)  \
  X(12, 0, 2, 0, false)   \
  X(12, 0, 3, 0, false)   \
  X(12, 0, 4, 0, false)   \
  X(24, 0, 24, 1, true)   \
  X(24, 0, 24, 1, false)  \
  X(24, 24, 24, 1, false) \
  X(24, 24, 24, 0, false) \
  X(24, 0, 24, 0, false)  \
  X(24, 0, 2, 0, false)   \
  X(24, 0, 3, 0, false)   \
  X(24, 0, 4, 0, false)

int pwconv_supported(int ci1, int ci2, int co, int act, int residual) {
#define X(A, B_, C, D, E) \
  if (ci1 == A && ci2 == B_ && co == C && (act < 0 || act == D) && (residual < 0 || (residual != 0) == E)) return 1;
  HNO_PW_CONFIGS(X)
#undef X
  return 0;
}

int pwconv_forward(const float* in1, const float* in2, const float* w, const float* bias, float* out, int B, int ci1,
                   int ci2, int co, long S, int act, int residual, cudaStream_t st) {
  HNO_CHECK(in1 && w && out && (ci2 == 0 || in2), "pwconv_forward: null pointer");
  HNO_CHECK(B >= 1 && B <= 65535 && S >= 1, "pwconv_forward: bad sizes B=%d S=%ld", B, S);
  if (!residual && (ci2 == 0 || ci2 == ci1) && S >= 4096 && ci1 >= 8) {
    // tensor-core path (tcgen05, 3xTF32): the activation streams through TMA exactly once
    TcStreamArgs a{};
    a.a[0] = in1;
    a.a[1] = in2;
    a.lda[0] = a.lda[1] = S;
    a.gsa[0] = (long)ci1 * S;
    a.gsa[1] = (long)ci2 * S;
    a.rows[0] = ci1;
    a.rows[1] = ci2;
    a.nsrc = ci2 > 0 ? 2 : 1;
    a.mext = S;
    a.G = B;
    // channel counts that are no multiple of 8 (12-channel HartleyMHASeg tensors) ride in 16-row chunks whose missing rows
    // the TMA zero-fills (the B image skips the gap, tc_stream.cu)
    a.kc = ci1 % 32 == 0 ? 32 : (ci1 % 24 == 0 ? 24 : (ci1 % 16 == 0 ? 16 : (ci1 % 8 == 0 ? 8 : (ci1 < 16 ? 16 : 0))));
    a.chunks_per_src = a.kc ? (ci1 + a.kc - 1) / a.kc : 0;
    a.b = w;
    a.ldbn = ci1 + ci2;
    a.ldbk = 1;
    a.kvalid = ci1 + ci2;
    a.scale = 1.f;
    a.bias = bias;
    a.out = out;
    a.ldo = S;
    a.gso = (long)co * S;
    a.nout = co;
    a.valid_m = S;
    a.act = act;
    a.epi = 0;
    // two-source (concat) convolutions: plain 512-byte-row TMA boxes + hi/lo images written by the split warps
    // (pw48f 0.179 -> 0.145 ms); one-source ones tie and keep the cp.async loader
    a.loader = a.nsrc == 2 ? 2 : 0;
    if (a.kc && (act == 0 || act == 1) && reinterpret_cast<uintptr_t>(out) % 16 == 0 && tc_stream_eligible(a))
      return tc_stream_launch(a, st);
  }
#define X(A, B_, C, D, E)                                                     \
  if (ci1 == A && ci2 == B_ && co == C && act == D && (residual != 0) == E)   \
    return fwd_t<A, B_, C, D, E>(in1, in2, w, bias, out, B, S, st);
  HNO_PW_CONFIGS(X)
#undef X
  set_error("pwconv_forward: unsupported configuration ci1=%d ci2=%d co=%d act=%d residual=%d", ci1, ci2, co, act,
            residual);
  return -1;
}

size_t pwconv_backward_workspace_bytes(int ci1, int ci2, int co) {
  return (size_t)(sm_count() * 2 + 8) * ((size_t)co * (ci1 + ci2) + co) * sizeof(float);
}

int pwconv_backward(const float* dy, const float* y, const float* in1, const float* in2, const float* w, float* din1,
                    float* din2, float* dweight, float* dbias, void* ws, int B, int ci1, int ci2, int co, long S,
                    long P, long HW, int act, int residual, int flags, cudaStream_t st) {
  HNO_CHECK(dy && in1 && w && dweight && ws && (ci2 == 0 || in2) && (act == 0 || y),
            "pwconv_backward: null pointer");
  HNO_CHECK(B >= 1 && S >= 1 && P >= 1 && HW >= 1 && HW <= P, "pwconv_backward: bad sizes");
  if (pwconv_bwd_tc_eligible(dy, y, in1, in2, din1, din2, ci1, ci2, co, S, act, residual))
    return pwconv_bwd_tc(dy, y, in1, in2, w, din1, din2, dweight, dbias, ws, B, ci2, S, P, HW, flags, st);
#define X(A, B_, C, D, E)                                                                                   \
  if (ci1 == A && ci2 == B_ && co == C && act == D && (residual != 0) == E)                                 \
    return bwd_t<A, B_, C, D, E>(dy, y, in1, in2, w, din1, din2, dweight, dbias, ws, B, S, P, HW, flags, st);
  HNO_PW_CONFIGS(X)
#undef X
  set_error("pwconv_backward: unsupported configuration ci1=%d ci2=%d co=%d act=%d residual=%d", ci1, ci2, co, act,
            residual);
  return -1;
}

}  // namespace hno
