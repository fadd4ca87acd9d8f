// Stem of HNOSeg-XS: Conv3d(kernel 2, stride 2, padding 1) + bias + SELU, forward and weight gradient.
//
// Replaces nets/hnosegxs.py:102-105,150-151 -> nets/nets_utils.py:156-163 (ConvNormAct with
// kernel_size=2, stride=2: padding = kernel_size // 2 = 1).  With stride == kernel the windows do not
// overlap, so the op is a gather of a 2x2x2xCIN patch per output voxel followed by a (8*CIN -> F)
// channel mix: every input voxel is read exactly once.  The model input needs no gradient, so the
// backward is the weight/bias gradient only.
#include "common.cuh"
#include "hno_b200.h"
#include "wgrad.cuh"

#include <stdlib.h>

namespace hno {

struct StemGeom {
  int Dx, Hx, Wx, D, H, W;
  long P, S, HWx;
};

__device__ __forceinline__ float stem_tap(const float* __restrict__ xc, const StemGeom& g, int d, int h, int w, int kd,
                                          int kh, int kw) {
  const int zd = 2 * d - 1 + kd, zh = 2 * h - 1 + kh, zw = 2 * w - 1 + kw;
  const bool inb = zd >= 0 && zd < g.Dx && zh >= 0 && zh < g.Hx && zw >= 0 && zw < g.Wx;
  return inb ? __ldg(xc + (zd * (int)g.HWx + zh * g.Wx + zw)) : 0.f;  // one channel volume fits 32-bit offsets
}

template <int CIN, int F>
__global__ void __launch_bounds__(256) k_stem_fwd(const float* __restrict__ x, const float* __restrict__ weight,
                                                  const float* __restrict__ bias, float* __restrict__ out,
                                                  StemGeom g) {
  constexpr int Q = 8 * CIN;
  constexpr int Fp = (F + 3) & ~3;
  __shared__ __align__(16) float wt[Q * Fp];  // wt[q][o]
  __shared__ float sb[Fp];
  for (int idx = threadIdx.x; idx < Q * Fp; idx += 256) {
    int q = idx / Fp, o = idx - q * Fp;
    wt[idx] = o < F ? weight[o * Q + q] : 0.f;
  }
  for (int o = threadIdx.x; o < Fp; o += 256) sb[o] = (bias && o < F) ? bias[o] : 0.f;
  __syncthreads();
  const long s = blockIdx.x * 256L + threadIdx.x;
  if (s >= g.S) return;
  const int b = blockIdx.y;
  float* po = out + (long)b * F * g.S + s;
  const int d = (int)s / (int)g.P;  // S < 2^31 (make_geom)
  const int p = (int)s - d * (int)g.P;
  if (p >= g.H * g.W) {
#pragma unroll
    for (int o = 0; o < F; ++o) po[(long)o * g.S] = 0.f;
    return;
  }
  const int h = p / g.W, w = p - h * g.W;
  // packed FFMA2 (two output channels per issue slot; scalar FFMA issues at half rate on sm_100): the 8 * CIN * F
  // multiply-adds per voxel were the kernel's bound (0.23 ms against 0.08 ms of HBM time)
  float2 acc[Fp / 2];
#pragma unroll
  for (int o = 0; o < Fp / 2; ++o) acc[o] = make_float2(sb[2 * o], sb[2 * o + 1]);
#pragma unroll
  for (int i = 0; i < CIN; ++i) {
    const float* xc = x + ((long)b * CIN + i) * g.Dx * g.HWx;
    float xv[8];
#pragma unroll
    for (int t = 0; t < 8; ++t) xv[t] = stem_tap(xc, g, d, h, w, t >> 2, (t >> 1) & 1, t & 1);  // 8 loads in flight
#pragma unroll
    for (int t = 0; t < 8; ++t) {
      const float2 x2 = dup2(xv[t]);
      const float4* w4 = reinterpret_cast<const float4*>(wt + (i * 8 + t) * Fp);
#pragma unroll
      for (int q = 0; q < Fp / 4; ++q) {
        const float4 ww = w4[q];
        acc[2 * q] = ffma2(make_float2(ww.x, ww.y), x2, acc[2 * q]);
        acc[2 * q + 1] = ffma2(make_float2(ww.z, ww.w), x2, acc[2 * q + 1]);
      }
    }
  }
#pragma unroll
  for (int o = 0; o < F / 2; ++o) {
    const float2 y = selu2(acc[o]);
    po[(long)(2 * o) * g.S] = y.x;
    po[(long)(2 * o + 1) * g.S] = y.y;
  }
  if (F & 1) po[(long)(F - 1) * g.S] = selu_f(acc[F / 2].x);
}

// dW[o][q] = sum_v dpre[o][v] * patch_q(v),  db[o] = sum_v dpre[o][v]
template <int CIN, int F>
__global__ void __launch_bounds__(kPwThreads, 3) k_stem_wgrad(const float* __restrict__ dpre,
                                                              const float* __restrict__ x,
                                                              float* __restrict__ partials, StemGeom g,
                                                              int tiles_per_sample, long total_tiles) {
  constexpr int Q = 8 * CIN;
  constexpr int TV = kPwThreads;
  constexpr int TVS = TV + 4;
  constexpr int PSTRIDE = F * Q + F;
  extern __shared__ float4 smem4[];
  float* sdp = reinterpret_cast<float*>(smem4);  // [F][TVS]
  float* sx = sdp + F * TVS;                     // [Q][TVS]
  using T = WgTile<F, Q>;
  float2 accW[T::TO][T::TI];
  float accB[T::TO];
#pragma unroll
  for (int q = 0; q < T::TO; ++q) {
    accB[q] = 0.f;
#pragma unroll
    for (int r = 0; r < T::TI; ++r) accW[q][r] = make_float2(0.f, 0.f);
  }
  for (long tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
    const int b = (int)(tile / tiles_per_sample);
    const long s = (tile - (long)b * tiles_per_sample) * TV + threadIdx.x;
    int d = 0, h = 0, w = 0;
    bool live = false;
    if (s < g.S) {
      d = (int)(s / g.P);
      const int p = (int)(s - (long)d * g.P);
      if (p < g.H * g.W) {
        live = true;
        h = p / g.W;
        w = p - h * g.W;
      }
    }
#pragma unroll 4
    for (int o = 0; o < F; ++o)
      sdp[o * TVS + threadIdx.x] = live ? __ldg(dpre + ((long)b * F + o) * g.S + s) : 0.f;
#pragma unroll
    for (int i = 0; i < CIN; ++i) {
      const float* xc = x + ((long)b * CIN + i) * g.Dx * g.HWx;
#pragma unroll
      for (int t = 0; t < 8; ++t)
        sx[(i * 8 + t) * TVS + threadIdx.x] = live ? stem_tap(xc, g, d, h, w, t >> 2, (t >> 1) & 1, t & 1) : 0.f;
    }
    __syncthreads();
    wgrad_tile<F, Q, TV, TVS>(sdp, sx, accW, accB, true);
    __syncthreads();
  }
  float* prow = partials + (long)blockIdx.x * PSTRIDE;
  wgrad_flush<F, Q>(sdp, accW, prow, Q, 0);
  {
    const int gi = threadIdx.x / T::G;
    const int l = threadIdx.x - gi * T::G;
    const int ot = l / T::N_IT;
    const int it = l - ot * T::N_IT;
    if (it == 0 && ot * T::TO < F) {
#pragma unroll
      for (int q = 0; q < T::TO; ++q) sx[gi * F + ot * T::TO + q] = accB[q];
    }
    __syncthreads();
    for (int o = threadIdx.x; o < F; o += kPwThreads) {
      float sum = 0.f;
#pragma unroll
      for (int gg = 0; gg < T::NG; ++gg) sum += sx[gg * F + o];
      prow[F * Q + o] = sum;
    }
  }
}

// Pipelined variant (round 2).  The kernel above is phase serialised (gather a tile, barrier, FFMA, barrier: HBM idle while it
// computes and vice versa) and its 3 x 4 register tile needs 7 LDS.128 per 24 FFMA2, which makes the shared-memory data
// pipe (4 wavefronts per LDS.128) its real bound: 3,584 wavefronts per 256 voxels against 1,536 FMA-pipe cycles.  Here
//   * tiles of 128 voxels are DOUBLE BUFFERED and filled with 4-byte cp.async (zero-filled outside the volume / in the plane
//     padding), issued for tile t+1 before the weight gradient of tile t runs;
//   * a warp owns 16 voxels of the tile and all F x Q outputs: lane (ot, it) accumulates F/4 x Q/8 outputs (6 x 4 for the
//     24-filter, 4-modality stem): 10 LDS.128 per 48 FFMA2.
template <int CIN, int F>
__global__ void __launch_bounds__(256, 3) k_stem_wgrad_pipe(const float* __restrict__ dpre, const float* __restrict__ x,
                                                            float* __restrict__ partials, StemGeom g,
                                                            int tiles_per_sample, long total_tiles) {
  constexpr int Q = 8 * CIN;
  constexpr int TV = 128, TVS = TV + 4;
  constexpr int TO = F / 4, TI = Q / 8;
  constexpr int STAGE = (F + Q) * TVS;  // floats per buffer: d(pre) rows then patch rows
  constexpr int PSTRIDE = F * Q + F;
  static_assert(F % 4 == 0 && 8 * (F * Q) <= 2 * STAGE, "stem weight gradient: unsupported shape");
  extern __shared__ float4 smem4[];
  float* buf = reinterpret_cast<float*>(smem4);  // [2][STAGE]
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int ot = lane >> 3, it = lane & 7;
  float2 accW[TO][TI];
  float accB[TO];
#pragma unroll
  for (int q = 0; q < TO; ++q) {
    accB[q] = 0.f;
#pragma unroll
    for (int r = 0; r < TI; ++r) accW[q][r] = make_float2(0.f, 0.f);
  }
  // loader role of this thread: voxel tid & 127 of the tile, half (tid >> 7) of the rows
  const int lv = tid & (TV - 1), lh = tid >> 7;
  // Address arithmetic of the gather is kept to one integer add, one widening add and one select per tap (the first version
  // spent 13 integer instructions per cp.async, more than the weight-gradient math of the tile): the 8 tap offsets are
  // constants of the launch, the in-bounds mask is 2 + 2 + 2 comparisons per voxel.
  int toff[8];
#pragma unroll
  for (int t = 0; t < 8; ++t) toff[t] = (t >> 2) * (int)g.HWx + ((t >> 1) & 1) * g.Wx + (t & 1);
  const long xcs = (long)g.Dx * g.HWx;  // channel stride of x
  const int P32 = (int)g.P, S32 = (int)g.S;  // 32-bit index arithmetic (the launcher checks the ranges): the 64-bit
                                             // divisions of the first version were two subroutine-sized sequences per tile
  auto issue = [&](long tile64, float* dst) {
    const unsigned tile = (unsigned)tile64;
    const int b = (int)(tile / (unsigned)tiles_per_sample);
    const int s = (int)(tile - (unsigned)b * (unsigned)tiles_per_sample) * TV + lv;
    int d = 0, h = 0, w = 0;
    bool live = false;
    if (s < S32) {
      d = s / P32;
      const int p = s - d * P32;
      if (p < g.H * g.W) {
        live = true;
        h = p / g.W;
        w = p - h * g.W;
      }
    }
    const unsigned d0 = (unsigned)__cvta_generic_to_shared(dst) + 4u * lv;
    const float* dp = dpre + ((long)b * F + lh * (F / 2)) * g.S + (live ? s : 0);
    const int dsz = live ? 4 : 0;
#pragma unroll
    for (int o = 0; o < F / 2; ++o)
      asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(d0 + 4u * ((lh * (F / 2) + o) * TVS)),
                   "l"(dp + (long)o * g.S), "r"(dsz)
                   : "memory");
    // tap (kd, kh, kw) reads x[2d - 1 + kd][2h - 1 + kh][2w - 1 + kw]
    const int zd = 2 * d - 1, zh = 2 * h - 1, zw = 2 * w - 1;
    const bool okd[2] = {live && zd >= 0, live && zd + 1 < g.Dx};
    const bool okh[2] = {zh >= 0, zh + 1 < g.Hx};
    const bool okw[2] = {zw >= 0, zw + 1 < g.Wx};
    const int base = zd * (int)g.HWx + zh * g.Wx + zw;  // may be negative at the border: only used when the tap is inside
    const float* xb = x + (long)b * CIN * xcs;
    if constexpr ((Q / 2) % 8 == 0) {
      // even channel counts: this thread's half of the rows is whole channels, so the tap index is a compile-time constant
      // and the shared-memory / channel offsets are immediates on two per-tile bases
      const unsigned dq = d0 + 4u * ((F + lh * (Q / 2)) * TVS);
      const float* xh = xb + (long)(lh * (CIN / 2)) * xcs;
#pragma unroll
      for (int ii = 0; ii < CIN / 2; ++ii) {
        const float* xi = xh + (long)ii * xcs;
#pragma unroll
        for (int t = 0; t < 8; ++t) {
          const bool inb = okd[t >> 2] && okh[(t >> 1) & 1] && okw[t & 1];
          asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(dq + 4u * ((ii * 8 + t) * TVS)),
                       "l"(xi + (inb ? base + toff[t] : 0)), "r"(inb ? 4 : 0)
                       : "memory");
        }
      }
    } else {
#pragma unroll
      for (int qq = 0; qq < Q / 2; ++qq) {
        const int q = lh * (Q / 2) + qq, i = q >> 3, t = q & 7;
        const bool inb = okd[t >> 2] && okh[(t >> 1) & 1] && okw[t & 1];
        const float* src = xb + (long)i * xcs + (inb ? base + toff[t] : 0);
        asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(d0 + 4u * ((F + q) * TVS)), "l"(src),
                     "r"(inb ? 4 : 0)
                     : "memory");
      }
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };
  long tile = blockIdx.x;
  if (tile < total_tiles) issue(tile, buf);
  int cur = 0;
  for (; tile < total_tiles; tile += gridDim.x, cur ^= 1) {
    const long next = tile + gridDim.x;
    if (next < total_tiles) {
      issue(next, buf + (cur ^ 1) * STAGE);
      asm volatile("cp.async.wait_group 1;" ::: "memory");
    } else {
      asm volatile("cp.async.wait_group 0;" ::: "memory");
    }
    __syncthreads();
    const float* sdp = buf + cur * STAGE + (ot * TO) * TVS + warp * 16;
    const float* sx = buf + cur * STAGE + (F + it) * TVS + warp * 16;
#pragma unroll 2
    for (int v = 0; v < 16; v += 4) {
      float4 dd[TO], xx[TI];
#pragma unroll
      for (int q = 0; q < TO; ++q) dd[q] = *reinterpret_cast<const float4*>(sdp + q * TVS + v);
#pragma unroll
      for (int r = 0; r < TI; ++r) xx[r] = *reinterpret_cast<const float4*>(sx + (r * 8) * TVS + v);  // rows r * 8 + it
#pragma unroll
      for (int q = 0; q < TO; ++q) {
#pragma unroll
        for (int r = 0; r < TI; ++r) {
          float2 a = accW[q][r];
          a = ffma2(make_float2(dd[q].x, dd[q].y), make_float2(xx[r].x, xx[r].y), a);
          a = ffma2(make_float2(dd[q].z, dd[q].w), make_float2(xx[r].z, xx[r].w), a);
          accW[q][r] = a;
        }
        if (it == 0) accB[q] += (dd[q].x + dd[q].y) + (dd[q].z + dd[q].w);
      }
    }
    __syncthreads();  // the buffer is refilled by the next iteration's prefetch
  }
  // ---- per-CTA partial row: the 8 warps' tiles through shared memory
  float* prow = partials + (long)blockIdx.x * PSTRIDE;
  float* scratch = buf;  // [8][F * Q], then [8][F]
#pragma unroll
  for (int q = 0; q < TO; ++q) {
#pragma unroll
    for (int r = 0; r < TI; ++r) scratch[warp * (F * Q) + (ot * TO + q) * Q + r * 8 + it] = accW[q][r].x + accW[q][r].y;
    if (it == 0) scratch[8 * F * Q + warp * F + ot * TO + q] = accB[q];
  }
  __syncthreads();
  for (int idx = tid; idx < F * Q + F; idx += 256) {
    float sum = 0.f;
    if (idx < F * Q) {
#pragma unroll
      for (int ww = 0; ww < 8; ++ww) sum += scratch[ww * (F * Q) + idx];
    } else {
#pragma unroll
      for (int ww = 0; ww < 8; ++ww) sum += scratch[8 * F * Q + ww * F + (idx - F * Q)];
    }
    prow[idx] = sum;
  }
}

#define HNO_STEM_CONFIGS(X) X(1, 8) X(2, 8) X(4, 8) X(1, 12) X(2, 12) X(4, 12) X(1, 24) X(2, 24) X(3, 24) X(4, 24)

int stem_supported(int cin, int f) {
#define X(A, B_) \
  if (cin == A && f == B_) return 1;
  HNO_STEM_CONFIGS(X)
#undef X
  return 0;
}

static int make_geom(StemGeom* g, int Dx, int Hx, int Wx, long P) {
  g->Dx = Dx; g->Hx = Hx; g->Wx = Wx;
  g->D = Dx / 2 + 1; g->H = Hx / 2 + 1; g->W = Wx / 2 + 1;
  g->P = P;
  g->S = (long)g->D * P;
  g->HWx = (long)Hx * Wx;
  HNO_CHECK(Dx >= 1 && Hx >= 1 && Wx >= 1, "stem: bad input size %dx%dx%d", Dx, Hx, Wx);
  HNO_CHECK((long)Dx * Hx * Wx < (1L << 31), "stem: input volume too large for 32-bit offsets");
  HNO_CHECK(P >= (long)g->H * g->W, "stem: plane pitch %ld < H*W = %ld", P, (long)g->H * g->W);
  HNO_CHECK(g->S < (1L << 31) - 256, "stem: output volume too large for 32-bit voxel indices");
  return 0;
}

int stem_forward(const float* x, const float* weight, const float* bias, float* out, int B, int cin, int f, int Dx,
                 int Hx, int Wx, long P, cudaStream_t st) {
  HNO_CHECK(x && weight && out, "stem_forward: null pointer");
  StemGeom g;
  if (make_geom(&g, Dx, Hx, Wx, P)) return -1;
  dim3 grid(ceil_div(g.S, 256), B);
#define X(A, B_)                                                                  \
  if (cin == A && f == B_) {                                                      \
    k_stem_fwd<A, B_><<<grid, 256, 0, st>>>(x, weight, bias, out, g);             \
    HNO_LAUNCH_CHECK();                                                           \
    return 0;                                                                     \
  }
  HNO_STEM_CONFIGS(X)
#undef X
  set_error("stem_forward: unsupported configuration in_channels=%d filters=%d", cin, f);
  return -1;
}

size_t stem_backward_workspace_bytes(int cin, int f) {
  return (size_t)(sm_count() * 3 + 8) * ((size_t)f * 8 * cin + f) * sizeof(float);
}

template <int CIN, int F>
static int stem_bwd_t(const float* dpre, const float* x, float* dweight, float* dbias, void* ws, int B,
                      const StemGeom& g, int accumulate, cudaStream_t st) {
  constexpr int Q = 8 * CIN;
  float* partials = reinterpret_cast<float*>(ws);
  const long gmax = (long)sm_count() * 3;  // <= 60 KB of shared memory per CTA: three fit, and the phases of one hide behind the others
  static const bool pipe = !(getenv("HNO_STEM_PIPE") && atoi(getenv("HNO_STEM_PIPE")) == 0);  // 0: the phase-serialised kernel
  int grid;
  if (pipe) {
    constexpr int TV = 128;
    const int tps = ceil_div(g.S, TV);
    const long total = (long)tps * B;
    HNO_CHECK(g.S + TV < (1L << 31) && total < (1L << 31), "stem_backward: volume too large for 32-bit tile indices");
    grid = (int)(total < gmax ? total : gmax);
    const size_t smem = (size_t)2 * (F + Q) * (TV + 4) * sizeof(float) + (8 * F + 8) * sizeof(float);
    auto kern = k_stem_wgrad_pipe<CIN, F>;
    HNO_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kern<<<grid, 256, smem, st>>>(dpre, x, partials, g, tps, total);
  } else {
    constexpr int TV = kPwThreads;
    const int tps = ceil_div(g.S, TV);
    const long total = (long)tps * B;
    grid = (int)(total < gmax ? total : gmax);
    const size_t smem = (size_t)(F + Q) * (TV + 4) * sizeof(float);
    auto kern = k_stem_wgrad<CIN, F>;
    HNO_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kern<<<grid, kPwThreads, smem, st>>>(dpre, x, partials, g, tps, total);
  }
  HNO_LAUNCH_CHECK();
  return reduce_partials(partials, grid, F * Q, F, dweight, dbias, accumulate, st);
}

int stem_backward(const float* dpre, const float* x, float* dweight, float* dbias, void* ws, int B, int cin, int f,
                  int Dx, int Hx, int Wx, long P, int accumulate, cudaStream_t st) {
  HNO_CHECK(dpre && x && dweight && ws, "stem_backward: null pointer");
  StemGeom g;
  if (make_geom(&g, Dx, Hx, Wx, P)) return -1;
#define X(A, B_) \
  if (cin == A && f == B_) return stem_bwd_t<A, B_>(dpre, x, dweight, dbias, ws, B, g, accumulate, st);
  HNO_STEM_CONFIGS(X)
#undef X
  set_error("stem_backward: unsupported configuration in_channels=%d filters=%d", cin, f);
  return -1;
}

// Input gradient (the transposed convolution).  Stride == kernel, so every input voxel z feeds exactly one output voxel
// d = (z + 1) >> 1 through tap k = (z + 1) & 1 per axis:  dx[i][z] = sum_o W[o][i][k] dpre[o][d].  One thread per input voxel;
// the eight voxels of a patch read the same F values of dpre (cache hits).  Only needed when the caller asks for the
// gradient w.r.t. the image (autograd-differentiable w.r.t. the input); training never does.
template <int CIN, int F>
__global__ void __launch_bounds__(256) k_stem_dx(const float* __restrict__ dpre, const float* __restrict__ weight,
                                                 float* __restrict__ dx, StemGeom g) {
  constexpr int Q = 8 * CIN;
  __shared__ float wt[Q * F];  // wt[q][o]
  for (int idx = threadIdx.x; idx < Q * F; idx += 256) {
    const int q = idx / F, o = idx - q * F;
    wt[idx] = weight[o * Q + q];
  }
  __syncthreads();
  const long Nx = (long)g.Dx * g.HWx;
  const long v = blockIdx.x * 256L + threadIdx.x;
  if (v >= Nx) return;
  const int b = blockIdx.y;
  const unsigned vu = (unsigned)v;  // Nx < 2^31 (make_geom)
  const unsigned r = vu / (unsigned)g.Wx;
  const int zw = (int)(vu - r * (unsigned)g.Wx);
  const int zd = (int)(r / (unsigned)g.Hx);
  const int zh = (int)(r - (unsigned)zd * (unsigned)g.Hx);
  const int d = (zd + 1) >> 1, h = (zh + 1) >> 1, w = (zw + 1) >> 1;
  const int t = (((zd + 1) & 1) << 2) | (((zh + 1) & 1) << 1) | ((zw + 1) & 1);
  const float* dp = dpre + (long)b * F * g.S + (long)d * g.P + (long)h * g.W + w;
  float acc[CIN];
#pragma unroll
  for (int i = 0; i < CIN; ++i) acc[i] = 0.f;
#pragma unroll 4
  for (int o = 0; o < F; ++o) {
    const float gv = __ldg(dp + (long)o * g.S);
#pragma unroll
    for (int i = 0; i < CIN; ++i) acc[i] = fmaf(wt[(i * 8 + t) * F + o], gv, acc[i]);
  }
#pragma unroll
  for (int i = 0; i < CIN; ++i) dx[((long)b * CIN + i) * Nx + v] = acc[i];
}

int stem_backward_input(const float* dpre, const float* weight, float* dx, int B, int cin, int f, int Dx, int Hx, int Wx,
                        long P, cudaStream_t st) {
  HNO_CHECK(dpre && weight && dx, "stem_backward_input: null pointer");
  HNO_CHECK(B >= 1 && B <= 65535, "stem_backward_input: bad batch size");
  StemGeom g;
  if (make_geom(&g, Dx, Hx, Wx, P)) return -1;
  dim3 grid(ceil_div((long)Dx * g.HWx, 256), B);
#define X(A, B_)                                                      \
  if (cin == A && f == B_) {                                          \
    k_stem_dx<A, B_><<<grid, 256, 0, st>>>(dpre, weight, dx, g);      \
    HNO_LAUNCH_CHECK();                                               \
    return 0;                                                         \
  }
  HNO_STEM_CONFIGS(X)
#undef X
  set_error("stem_backward_input: unsupported configuration in_channels=%d filters=%d", cin, f);
  return -1;
}

}  // namespace hno
