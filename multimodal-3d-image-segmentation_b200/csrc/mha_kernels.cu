// Hartley multi-head attention on the retained modes (reference nets/hartley_mha.py:136-222, 310-334, 473-524).
//
// The frequency-domain tensors are tiny next to the activations (1.5 MB per 24 channels and sample), so everything around
// the two big contractions is plain CUDA-core code whose only job is to hand the GEMM engine (gemm_tc.cu) K-major
// operands:
//   project   : per-head 1x1x1 convolution of the mode tensor (freq_conv3d, :310-334; the eight corner einsums of the
//               reference are one convolution over the cropped block) + patch grouping (grouping3d, :473-498) written
//               straight into the attention layouts  X_tok [b, head][token][feature]  and  X_chan [b, head][feature][token],
//               zero padded to multiples of 128 tokens / 32 features (zero rows are inert: selu(0) = 0);
//   attention : P = selu(Q K^T / sqrt(F)),  O = P V                          (:198-204), three GEMM launches backward -> six;
//   output    : ungrouping (ungrouping3d, :501-524) + the head-mixing einsum 'oi,bidhw->bodhw' (:213-216) back onto the
//               mode tensor [b][channel][mode].
// feature f = c * P + patch_offset,  token t = patch index,  exactly the channel / position order grouping3d produces.
#include "common.cuh"
#include "gemm_tc.h"

#include <stdlib.h>

namespace hno {

struct MhaGeom {
  int B, H;            // samples, heads
  int Ld, Lh, Lw;      // retained mode block
  int pd, ph, pw;      // patch
  int T, Tp;           // tokens, padded tokens (multiple of 128)
  long M;              // Ld * Lh * Lw
};

__device__ __forceinline__ void mha_locate(const MhaGeom& g, int m, int& t, int& po) {
  const int w = m % g.Lw;
  const int r = m / g.Lw;
  const int h = r % g.Lh;
  const int d = r / g.Lh;
  const int nh = g.Lh / g.ph, nw = g.Lw / g.pw;
  t = ((d / g.pd) * nh + h / g.ph) * nw + w / g.pw;
  po = ((d % g.pd) * g.ph + h % g.ph) * g.pw + w % g.pw;
}

// Token decomposition without per-element divisions: token t -> (td, th, tw); mode index of patch offset (id, ih, iw).
struct MhaTok {
  int td, th, tw;
};
__device__ __forceinline__ MhaTok mha_token(const MhaGeom& g, int t) {
  const int nh = g.Lh / g.ph, nw = g.Lw / g.pw;
  MhaTok k;
  k.tw = t % nw;
  const int r = t / nw;
  k.th = r % nh;
  k.td = r / nh;
  return k;
}

constexpr int kMhaPMax = 8;  // patch sizes up to 8 modes (2 x 2 x 2) take the register-resident fast paths
// mode offsets of the (up to kMhaPMax) patch positions of token k, in grouping3d's (id, ih, iw) order
__device__ __forceinline__ void mha_patch_offsets(const MhaGeom& g, const MhaTok& k, int (&moff)[kMhaPMax]) {
  const int base = ((k.td * g.pd) * g.Lh + k.th * g.ph) * g.Lw + k.tw * g.pw;
#pragma unroll
  for (int po = 0; po < kMhaPMax; ++po) {
    const int iw = po % g.pw, r = po / g.pw;
    const int ih = r % g.ph, id = r / g.ph;
    moff[po] = base + (id * g.Lh + ih) * g.Lw + iw;
  }
}

// Weight element (h, c, i) lives at w[h * sh + c * sc + i * si]: the projections read weight_{query,key,value} [H][cd][cin]
// (sh = cd cin, sc = cin, si = 1), the backward of the output projection reads weight_out [co][H cd] transposed
// (sh = cd, sc = 1, si = H cd).
struct MhaW {
  long sh, sc, si;
};

// src [B][cin][M] (mode tensor), w, bias [H][cd] or null -> x_tok [B*H][Tp][Fp], x_chan [B*H][Fp][Tp], padding included.
// grid (Tp / 256, cd, B*H): a thread owns token t of channel c of head h and produces its P patch features: the
// feature-major copy is written coalesced (consecutive threads = consecutive tokens), the token-major one in runs of P
// floats; tokens >= T and (last channel block) features >= cd * P are written as zeros -- no memset passes.
__global__ void __launch_bounds__(256) k_mha_project_fwd(const float* __restrict__ src, const float* __restrict__ w,
                                                         const float* __restrict__ bias, float* __restrict__ x_tok,
                                                         float* __restrict__ x_chan, MhaGeom g, MhaW ws, int cin, int cd,
                                                         int Fp) {
  extern __shared__ float swt[];  // the cin weights of this (h, c)
  const int c = blockIdx.y, bh = blockIdx.z, b = bh / g.H, h = bh % g.H;
  for (int i = threadIdx.x; i < cin; i += 256) swt[i] = __ldg(w + h * ws.sh + c * ws.sc + i * ws.si);
  __syncthreads();
  const int t = blockIdx.x * 256 + threadIdx.x;
  if (t >= g.Tp) return;
  const int P = g.pd * g.ph * g.pw;
  float* xt = x_tok + ((long)bh * g.Tp + t) * Fp + c * P;
  float* xc = x_chan + ((long)bh * Fp + c * P) * g.Tp + t;
  const bool live = t < g.T;
  const float b0 = bias ? __ldg(bias + h * cd + c) : 0.f;
  const MhaTok k = live ? mha_token(g, t) : MhaTok{0, 0, 0};
  const float* sp = src + (long)b * cin * g.M;
  if (P <= kMhaPMax) {
    // the P mode offsets of this token's patch, then cin x P independent loads (the naive nest -- one load, one FMA, next --
    // ran at one L2 round trip per element: 27 us for 3 MB)
    int moff[kMhaPMax];
    mha_patch_offsets(g, k, moff);
    float acc[kMhaPMax];
#pragma unroll
    for (int po = 0; po < kMhaPMax; ++po) acc[po] = live ? b0 : 0.f;
    if (live) {
#pragma unroll 2
      for (int i = 0; i < cin; ++i) {
        const float wv = swt[i];
        const float* zi = sp + (long)i * g.M;
#pragma unroll
        for (int po = 0; po < kMhaPMax; ++po)
          if (po < P) acc[po] = fmaf(wv, __ldg(zi + moff[po]), acc[po]);
      }
    }
#pragma unroll
    for (int po = 0; po < kMhaPMax; ++po)
      if (po < P) {
        xt[po] = acc[po];
        xc[(long)po * g.Tp] = acc[po];
      }
  } else {
    int po = 0;
    for (int id = 0; id < g.pd; ++id)
      for (int ih = 0; ih < g.ph; ++ih)
        for (int iw = 0; iw < g.pw; ++iw, ++po) {
          float acc = 0.f;
          if (live) {
            const long m = ((long)(k.td * g.pd + id) * g.Lh + k.th * g.ph + ih) * g.Lw + k.tw * g.pw + iw;
            acc = b0;
#pragma unroll 4
            for (int i = 0; i < cin; ++i) acc = fmaf(swt[i], __ldg(sp + (long)i * g.M + m), acc);
          }
          xt[po] = acc;
          xc[(long)po * g.Tp] = acc;
        }
  }
  if (c == cd - 1) {  // feature padding
    for (int f = cd * P; f < Fp; ++f) {
      x_tok[((long)bh * g.Tp + t) * Fp + f] = 0.f;
      x_chan[((long)bh * Fp + f) * g.Tp + t] = 0.f;
    }
  }
}

// dz [B][cin][M] (+)= sum_{h, c} w[h][c][i] * dx_tok[b, h][t][c P + po]
__global__ void __launch_bounds__(256) k_mha_project_bwd_z(const float* __restrict__ dx_tok, const float* __restrict__ w,
                                                           float* __restrict__ dz, MhaGeom g, int cin, int cd, int Fp,
                                                           int accumulate) {
  const long idx = blockIdx.x * 256L + threadIdx.x;
  if (idx >= g.M) return;
  const int m = (int)idx;
  const int i = blockIdx.y, b = blockIdx.z;
  int t, po;
  mha_locate(g, m, t, po);
  const int P = g.pd * g.ph * g.pw;
  float acc = 0.f;
  for (int h = 0; h < g.H; ++h) {
    const float* dx = dx_tok + ((long)(b * g.H + h) * g.Tp + t) * Fp + po;
    const float* wp = w + (long)h * cd * cin + i;
#pragma unroll 6
    for (int c = 0; c < cd; ++c) acc = fmaf(__ldg(wp + c * cin), __ldg(dx + c * P), acc);
  }
  float* o = dz + ((long)b * cin + i) * g.M + m;
  *o = accumulate ? *o + acc : acc;
}

// dw(h, c, i) = sum_{b, t, po} dx_tok[b, h][t][c P + po] * src[b][i][m(t, po)]          (dw indexed through MhaW like w)
// grid (cin + extra, cd, H) + bias blocks; one CTA per output, threads over (b, t) pairs (token decode once per pair, patch
// offsets in the inner loops: no per-element index divisions), bounded fp32 runs folded into fp64.
//   bias_mode 1 (projections): block i == cin sums dx_tok over everything            -> dbias[h][c]
//   bias_mode 2 (output projection, src = dy): blocks (c == 0, h == 0, i) of an extra grid row sum src[b][i][:] -> dbias[i]
__global__ void __launch_bounds__(256) k_mha_wgrad(const float* __restrict__ dx_tok, const float* __restrict__ src,
                                                   float* __restrict__ dw, float* __restrict__ dbias, MhaGeom g, MhaW ws,
                                                   int cin, int cd, int Fp, int bias_mode) {
  __shared__ double sred[8];
  const int i = blockIdx.x, c = blockIdx.y, h = blockIdx.z;
  const int P = g.pd * g.ph * g.pw;
  const bool proj_bias = bias_mode == 1 && i == cin;
  const bool out_bias = bias_mode == 2 && h == g.H;  // extra z-slice: only c == 0 does work
  if (out_bias && c != 0) return;
  double acc = 0.0;
  if (out_bias) {
    for (int b = 0; b < g.B; ++b) {
      const float* sp = src + ((long)b * cin + i) * g.M;
      float run = 0.f;
      int n = 0;
      for (long m = threadIdx.x; m < g.M; m += 256) {
        run += __ldg(sp + m);
        if (++n == 32) {
          acc += (double)run;
          run = 0.f;
          n = 0;
        }
      }
      acc += (double)run;
    }
  } else {
    const int pairs = g.B * g.T;
    for (int q = threadIdx.x; q < pairs; q += 256) {
      const int b = q / g.T, t = q - b * g.T;
      const MhaTok k = mha_token(g, t);
      const float* dx = dx_tok + ((long)(b * g.H + h) * g.Tp + t) * Fp + c * P;
      const float* sp = src + ((long)b * cin + (proj_bias ? 0 : i)) * g.M;
      float run = 0.f;
      if (P <= kMhaPMax) {
        int moff[kMhaPMax];
        mha_patch_offsets(g, k, moff);
        float d[kMhaPMax], x[kMhaPMax];
#pragma unroll
        for (int po = 0; po < kMhaPMax; ++po) {  // all loads first: 2 P independent requests per pair
          d[po] = po < P ? __ldg(dx + po) : 0.f;
          x[po] = (po < P && !proj_bias) ? __ldg(sp + moff[po]) : 1.f;
        }
#pragma unroll
        for (int po = 0; po < kMhaPMax; ++po) run = fmaf(d[po], x[po], run);
      } else {
        int po = 0;
        for (int id = 0; id < g.pd; ++id)
          for (int ih = 0; ih < g.ph; ++ih)
            for (int iw = 0; iw < g.pw; ++iw, ++po) {
              const float d = __ldg(dx + po);
              if (proj_bias) {
                run += d;
              } else {
                const long m = ((long)(k.td * g.pd + id) * g.Lh + k.th * g.ph + ih) * g.Lw + k.tw * g.pw + iw;
                run = fmaf(d, __ldg(sp + m), run);
              }
            }
      }
      acc += (double)run;
    }
  }
  acc = warp_sum_d(acc);
  if ((threadIdx.x & 31) == 0) sred[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    double tot = 0.0;
    for (int k = 0; k < 8; ++k) tot += sred[k];
    if (out_bias) dbias[i] = (float)tot;
    else if (proj_bias) dbias[h * cd + c] = (float)tot;
    else dw[h * ws.sh + c * ws.sc + i * ws.si] = (float)tot;
  }
}

// Tiled weight gradient (round 2).  The kernel above runs one CTA per output and re-reads (and re-gathers) both operands for
// every one of the H cd cin outputs: 82 us per call, four calls per attention block.  Here a CTA stages a 16-token tile of
// dx_tok (all heads) and the grouped patches of src ONCE in shared memory and every thread owns a few outputs; one partial
// row per tile, fp64 reduction by k_reduce_partials.  grid (ceil(T / 16), B).
constexpr int kWgTT = 16;
__global__ void __launch_bounds__(256) k_mha_wgrad_tile(const float* __restrict__ dx_tok, const float* __restrict__ src,
                                                        float* __restrict__ partials, MhaGeom g, MhaW ws, int cin, int cd,
                                                        int Fp, int bias_mode) {
  extern __shared__ float smw[];
  const int P = g.pd * g.ph * g.pw;
  const int HF = g.H * cd * P, HFp = HF + 4;  // dx tile pitch
  const int CF = cin * P, CFp = CF + 4;       // patch tile pitch
  float* dxs = smw;                // [kWgTT][HFp]
  float* zs = dxs + kWgTT * HFp;   // [kWgTT][CFp]
  const int b = blockIdx.y, t0 = blockIdx.x * kWgTT;
  const int F4 = (cd * P) >> 2;  // float4 pieces of a (token, head) row; (cd P) % 4 == 0 is checked by the launcher
  for (int idx = threadIdx.x; idx < kWgTT * g.H * F4; idx += 256) {
    const int f4 = idx % F4, r = idx / F4, h = r % g.H, t = r / g.H;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (t0 + t < g.T) v = *reinterpret_cast<const float4*>(dx_tok + ((long)(b * g.H + h) * g.Tp + t0 + t) * Fp + 4 * f4);
    *reinterpret_cast<float4*>(dxs + t * HFp + h * cd * P + 4 * f4) = v;
  }
  for (int idx = threadIdx.x; idx < kWgTT * cin; idx += 256) {
    const int t = idx % kWgTT, i = idx / kWgTT;  // consecutive threads = consecutive tokens = neighbouring patches
    const bool live = t0 + t < g.T;
    const MhaTok k = live ? mha_token(g, t0 + t) : MhaTok{0, 0, 0};
    const float* sp = src + ((long)b * cin + i) * g.M;
    int po = 0;
    for (int id = 0; id < g.pd; ++id)
      for (int ih = 0; ih < g.ph; ++ih)
        for (int iw = 0; iw < g.pw; ++iw, ++po) {
          const long m = ((long)(k.td * g.pd + id) * g.Lh + k.th * g.ph + ih) * g.Lw + k.tw * g.pw + iw;
          zs[t * CFp + i * P + po] = live ? __ldg(sp + m) : 0.f;
        }
  }
  __syncthreads();
  const int NO = g.H * cd * cin;
  const int nbias = bias_mode == 1 ? g.H * cd : (bias_mode == 2 ? cin : 0);
  float* prow = partials + ((long)blockIdx.y * gridDim.x + blockIdx.x) * (NO + nbias);
  for (int idx = threadIdx.x; idx < NO + nbias; idx += 256) {
    float acc = 0.f;
    if (idx < NO) {
      const int i = idx % cin, r = idx / cin, c = r % cd, h = r / cd;
      const float* dp = dxs + (h * cd + c) * P;
      const float* zp = zs + i * P;
      if ((P & 3) == 0) {
        for (int t = 0; t < kWgTT; ++t)
          for (int q = 0; q < P; q += 4) {
            const float4 d = *reinterpret_cast<const float4*>(dp + t * HFp + q);
            const float4 x = *reinterpret_cast<const float4*>(zp + t * CFp + q);
            acc = fmaf(d.x, x.x, fmaf(d.y, x.y, fmaf(d.z, x.z, fmaf(d.w, x.w, acc))));
          }
      } else {
        for (int t = 0; t < kWgTT; ++t)
          for (int q = 0; q < P; ++q) acc = fmaf(dp[t * HFp + q], zp[t * CFp + q], acc);
      }
      prow[h * ws.sh + c * ws.sc + i * ws.si] = acc;
    } else if (bias_mode == 1) {  // projections: dbias[h][c] = sum of dx_tok
      const float* dp = dxs + (idx - NO) * P;
      for (int t = 0; t < kWgTT; ++t)
        for (int q = 0; q < P; ++q) acc += dp[t * HFp + q];
      prow[idx] = acc;
    } else {                      // output projection (src = dy): dbias[i] = sum of src over the modes
      const float* zp = zs + (idx - NO) * P;
      for (int t = 0; t < kWgTT; ++t)
        for (int q = 0; q < P; ++q) acc += zp[t * CFp + q];
      prow[idx] = acc;
    }
  }
}

int reduce_partials(const float* partials, int nrows, int nw, int nb, float* dweight, float* dbias, int accumulate,
                    cudaStream_t st);

// Shared-memory footprint of the tiled weight gradient; 0 = shape not served by it (the per-output kernel takes over).
static size_t wgrad_tile_smem(const MhaGeom& g, int cin, int cd) {
  const int P = g.pd * g.ph * g.pw;
  if ((cd * P) % 4 != 0) return 0;
  const size_t bytes = (size_t)kWgTT * ((size_t)g.H * cd * P + 4 + (size_t)cin * P + 4) * sizeof(float);
  return bytes <= 200 * 1024 ? bytes : 0;
}

size_t mha_wgrad_workspace_bytes(int B, int H, int cin, int cd, int T) {
  return (size_t)ceil_div(T, kWgTT) * B * ((size_t)H * cd * cin + (H * cd > cin ? H * cd : cin)) * sizeof(float) + 256;
}

// dw (indexed through ws like the weight) and the optional bias gradient; `wsp` may be null (per-output kernel)
static int mha_wgrad(const float* dx_tok, const float* src, float* dw, float* dbias, void* wsp, const MhaGeom& g, MhaW ws,
                     int cin, int cd, int Fp, int bias_mode, cudaStream_t st) {
  const size_t smem = wgrad_tile_smem(g, cin, cd);
  static const bool tile_on = !(getenv("HNO_MHA_WGRAD_TILE") && atoi(getenv("HNO_MHA_WGRAD_TILE")) == 0);
  if (wsp != nullptr && smem > 0 && tile_on && g.B <= 65535) {
    const int NO = g.H * cd * cin;
    const int nbias = bias_mode == 1 ? g.H * cd : (bias_mode == 2 ? cin : 0);
    dim3 grid(ceil_div(g.T, kWgTT), g.B);
    HNO_CUDA(cudaFuncSetAttribute(k_mha_wgrad_tile, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k_mha_wgrad_tile<<<grid, 256, smem, st>>>(dx_tok, src, reinterpret_cast<float*>(wsp), g, ws, cin, cd, Fp, bias_mode);
    HNO_LAUNCH_CHECK();
    return reduce_partials(reinterpret_cast<const float*>(wsp), (int)(grid.x * grid.y), NO, nbias, dw, nbias ? dbias : nullptr, 0, st);
  }
  if (bias_mode == 2) {
    dim3 gw(cin, cd, g.H + 1);
    k_mha_wgrad<<<gw, 256, 0, st>>>(dx_tok, src, dw, dbias, g, ws, cin, cd, Fp, 2);
  } else {
    dim3 gw(cin + (bias_mode == 1 ? 1 : 0), cd, g.H);
    k_mha_wgrad<<<gw, 256, 0, st>>>(dx_tok, src, dw, dbias, g, ws, cin, cd, Fp, bias_mode);
  }
  HNO_LAUNCH_CHECK();
  return 0;
}

// o_tok [B*H][Tp][Fp] -> y [B][co][M] = sum_{h, c} wout[o][h cd + c] * o_tok[b, h][t][c P + po] + bias[o]
__global__ void __launch_bounds__(256) k_mha_output_fwd(const float* __restrict__ o_tok, const float* __restrict__ wout,
                                                        const float* __restrict__ bias, float* __restrict__ y, MhaGeom g,
                                                        int co, int cd, int Fp) {
  const long idx = blockIdx.x * 256L + threadIdx.x;
  if (idx >= g.M) return;
  const int m = (int)idx;
  const int o = blockIdx.y, b = blockIdx.z;
  int t, po;
  mha_locate(g, m, t, po);
  const int P = g.pd * g.ph * g.pw;
  float acc = bias ? __ldg(bias + o) : 0.f;
  for (int h = 0; h < g.H; ++h) {
    const float* op = o_tok + ((long)(b * g.H + h) * g.Tp + t) * Fp + po;
    const float* wp = wout + (long)o * g.H * cd + h * cd;
#pragma unroll 6
    for (int c = 0; c < cd; ++c) acc = fmaf(__ldg(wp + c), __ldg(op + c * P), acc);
  }
  y[((long)b * co + o) * g.M + m] = acc;
}

// ---- token-major glue kernels (round 2).  The kernels above re-read the mode tensor once per (head, channel) CTA (48 x for the
// BASELINE attention) or the token rows once per output channel: 27 - 40 us per launch for 3 - 6 MB.  Here a thread owns ONE
// TOKEN (8-mode patches, channel counts that are multiples of 4): it gathers / scatters its patch once and keeps 4 channels x 8
// patch positions of accumulators in registers per pass.
// modes -> tokens: x_tok [B*H][Tp][Fp], x_chan [B*H][Fp][Tp] = per-head 1x1x1 convolution of src [B][cin][M] + grouping.
// grid (Tp / 128, B*H), 128 threads; weights of the head in shared memory as [c][i].
__global__ void __launch_bounds__(128) k_mha_project_tok(const float* __restrict__ src, const float* __restrict__ w,
                                                         const float* __restrict__ bias, float* __restrict__ x_tok,
                                                         float* __restrict__ x_chan, MhaGeom g, MhaW ws, int cin, int cd,
                                                         int Fp) {
  constexpr int P = 8;
  extern __shared__ float swt[];  // [cd][cin]
  const int bh = blockIdx.y, b = bh / g.H, h = bh % g.H;
  for (int idx = threadIdx.x; idx < cd * cin; idx += 128) {
    const int c = idx / cin, i = idx - c * cin;
    swt[idx] = __ldg(w + h * ws.sh + c * ws.sc + i * ws.si);
  }
  __syncthreads();
  const int t = blockIdx.x * 128 + threadIdx.x;
  if (t >= g.Tp) return;
  const bool live = t < g.T;
  float* xt = x_tok + ((long)bh * g.Tp + t) * Fp;
  float* xc = x_chan + (long)bh * Fp * g.Tp + t;
  int moff[kMhaPMax];
  mha_patch_offsets(g, live ? mha_token(g, t) : MhaTok{0, 0, 0}, moff);
  const float* sp = src + (long)b * cin * g.M;
  for (int c0 = 0; c0 < cd; c0 += 4) {
    float acc[4][P];
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      const float b0 = (live && bias) ? __ldg(bias + h * cd + c0 + c) : 0.f;
#pragma unroll
      for (int po = 0; po < P; ++po) acc[c][po] = b0;
    }
    if (live) {
#pragma unroll 2
      for (int i = 0; i < cin; ++i) {
        float zv[P];
#pragma unroll
        for (int po = 0; po < P; ++po) zv[po] = __ldg(sp + (long)i * g.M + moff[po]);
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          const float wv = swt[(c0 + c) * cin + i];
#pragma unroll
          for (int po = 0; po < P; ++po) acc[c][po] = fmaf(wv, zv[po], acc[c][po]);
        }
      }
    }
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      float* o = xt + (c0 + c) * P;
      *reinterpret_cast<float4*>(o) = make_float4(acc[c][0], acc[c][1], acc[c][2], acc[c][3]);
      *reinterpret_cast<float4*>(o + 4) = make_float4(acc[c][4], acc[c][5], acc[c][6], acc[c][7]);
#pragma unroll
      for (int po = 0; po < P; ++po) xc[(long)((c0 + c) * P + po) * g.Tp] = acc[c][po];
    }
  }
  for (int f = cd * P; f < Fp; ++f) {  // feature padding
    xt[f] = 0.f;
    xc[(long)f * g.Tp] = 0.f;
  }
}

// tokens -> modes: out [B][nout][M] (+)= sum_{h, c} W(h, c, o) x_tok[b, h][t][c P + po] (+ bias[o]); W(h, c, o) = w[h sh + c sc + o si].
// grid (ceil(T / 128), nout / 4, B), 128 threads: a thread owns one token and FOUR outputs (one thread for all outputs left 32
// CTAs on the machine: 35 us, 70 us when accumulating); the weights of the four outputs in shared memory as [h][c][4].
__global__ void __launch_bounds__(128) k_mha_tok_to_modes(const float* __restrict__ x_tok, const float* __restrict__ w,
                                                          const float* __restrict__ bias, float* __restrict__ out, MhaGeom g,
                                                          MhaW ws, int nout, int cd, int Fp, int accumulate) {
  constexpr int P = 8;
  extern __shared__ float swt[];  // [H][cd][4]
  const int b = blockIdx.z, o0 = blockIdx.y * 4;
  for (int idx = threadIdx.x; idx < g.H * cd * 4; idx += 128) {
    const int o = idx & 3, r = idx >> 2, c = r % cd, h = r / cd;
    swt[idx] = __ldg(w + h * ws.sh + c * ws.sc + (o0 + o) * ws.si);
  }
  __syncthreads();
  const int t = blockIdx.x * 128 + threadIdx.x;
  if (t >= g.T) return;
  int moff[kMhaPMax];
  mha_patch_offsets(g, mha_token(g, t), moff);
  float* ob = out + (long)b * nout * g.M;
  {
    float acc[4][P];
#pragma unroll
    for (int o = 0; o < 4; ++o) {
      const float b0 = bias ? __ldg(bias + o0 + o) : 0.f;
#pragma unroll
      for (int po = 0; po < P; ++po) acc[o][po] = b0;
    }
    for (int h = 0; h < g.H; ++h) {
      const float* xr = x_tok + ((long)(b * g.H + h) * g.Tp + t) * Fp;
#pragma unroll 2
      for (int c = 0; c < cd; ++c) {
        const float4 v0 = *reinterpret_cast<const float4*>(xr + c * P), v1 = *reinterpret_cast<const float4*>(xr + c * P + 4);
        const float xv[P] = {v0.x, v0.y, v0.z, v0.w, v1.x, v1.y, v1.z, v1.w};
        const float* wp = swt + (h * cd + c) * 4;
#pragma unroll
        for (int o = 0; o < 4; ++o) {
          const float wv = wp[o];
#pragma unroll
          for (int po = 0; po < P; ++po) acc[o][po] = fmaf(wv, xv[po], acc[o][po]);
        }
      }
    }
#pragma unroll
    for (int o = 0; o < 4; ++o) {
      float* dst = ob + (long)(o0 + o) * g.M;
#pragma unroll
      for (int po = 0; po < P; ++po) dst[moff[po]] = accumulate ? dst[moff[po]] + acc[o][po] : acc[o][po];
    }
  }
}

// the token-major kernels serve 8-mode patches and channel counts that are multiples of 4 (HNO_MHA_TOK=0: never)
static bool mha_tok_path(const MhaGeom& g, int cin, int cd) {
  static const bool on = !(getenv("HNO_MHA_TOK") && atoi(getenv("HNO_MHA_TOK")) == 0);
  return on && g.pd * g.ph * g.pw == 8 && cin % 4 == 0 && cd % 4 == 0 && (size_t)g.H * cd * cin * sizeof(float) <= 48 * 1024;
}

// ------------------------------------------------------------------------------------------------ host side
static int make_geom(MhaGeom* g, int B, int H, int Ld, int Lh, int Lw, int pd, int ph, int pw, int Tp) {
  HNO_CHECK(B >= 1 && H >= 1 && Ld >= 1 && Lh >= 1 && Lw >= 1, "mha: bad sizes");
  HNO_CHECK(pd >= 1 && ph >= 1 && pw >= 1 && Ld % pd == 0 && Lh % ph == 0 && Lw % pw == 0,
            "mha: the retained modes (%d, %d, %d) must be divisible by the patch size (%d, %d, %d)", Ld, Lh, Lw, pd, ph, pw);
  g->B = B, g->H = H, g->Ld = Ld, g->Lh = Lh, g->Lw = Lw, g->pd = pd, g->ph = ph, g->pw = pw;
  g->T = (Ld / pd) * (Lh / ph) * (Lw / pw);
  g->Tp = Tp;
  g->M = (long)Ld * Lh * Lw;
  HNO_CHECK(Tp >= g->T && Tp % 128 == 0, "mha: the padded token count must be a multiple of 128 and >= %d", g->T);
  HNO_CHECK((long)B * H <= 65535 && g->M < (1L << 30), "mha: problem too large");
  return 0;
}

int mha_project_forward(const float* z, const float* w, const float* bias, float* x_tok, float* x_chan, int B, int H,
                        int cin, int cd, int Ld, int Lh, int Lw, int pd, int ph, int pw, int Tp, int Fp, cudaStream_t st) {
  MhaGeom g;
  if (make_geom(&g, B, H, Ld, Lh, Lw, pd, ph, pw, Tp)) return -1;
  HNO_CHECK(z && w && x_tok && x_chan, "mha_project_forward: null pointer");
  HNO_CHECK(Fp >= cd * pd * ph * pw && Fp % 32 == 0, "mha_project_forward: feature pitch %d too small / not a multiple of 32", Fp);
  HNO_CHECK(cd <= 65535 && cin * sizeof(float) <= 48 * 1024, "mha_project_forward: too many channels");
  const MhaW ws{(long)cd * cin, (long)cin, 1};
  if (mha_tok_path(g, cin, cd)) {
    dim3 grid(Tp / 128, B * H);
    k_mha_project_tok<<<grid, 128, (size_t)cd * cin * sizeof(float), st>>>(z, w, bias, x_tok, x_chan, g, ws, cin, cd, Fp);
  } else {
    dim3 grid(ceil_div(Tp, 256), cd, B * H);
    k_mha_project_fwd<<<grid, 256, cin * sizeof(float), st>>>(z, w, bias, x_tok, x_chan, g, ws, cin, cd, Fp);
  }
  HNO_LAUNCH_CHECK();
  return 0;
}

int mha_project_backward(const float* dx_tok, const float* z, const float* w, float* dz, float* dw, float* dbias, void* wsp,
                         int B, int H, int cin, int cd, int Ld, int Lh, int Lw, int pd, int ph, int pw, int Tp, int Fp,
                         int accumulate_dz, cudaStream_t st) {
  MhaGeom g;
  if (make_geom(&g, B, H, Ld, Lh, Lw, pd, ph, pw, Tp)) return -1;
  HNO_CHECK(dx_tok && z && w, "mha_project_backward: null pointer");
  HNO_CHECK(cin <= 65535 && B <= 65535, "mha_project_backward: too many channels");
  if (dz) {
    if (mha_tok_path(g, cin, cd)) {
      const MhaW wsz{(long)cd * cin, (long)cin, 1};  // W(h, c, i) = w[h][c][i]
      dim3 grid(ceil_div(g.T, 128), cin / 4, B);
      k_mha_tok_to_modes<<<grid, 128, (size_t)H * cd * 4 * sizeof(float), st>>>(dx_tok, w, nullptr, dz, g, wsz, cin, cd, Fp,
                                                                              accumulate_dz);
    } else {
      dim3 grid(ceil_div(g.M, 256), cin, B);
      k_mha_project_bwd_z<<<grid, 256, 0, st>>>(dx_tok, w, dz, g, cin, cd, Fp, accumulate_dz);
    }
    HNO_LAUNCH_CHECK();
  }
  if (dw) {
    const MhaW ws{(long)cd * cin, (long)cin, 1};
    return mha_wgrad(dx_tok, z, dw, dbias, wsp, g, ws, cin, cd, Fp, dbias ? 1 : 0, st);
  }
  return 0;
}

// P [BH][Tp][Tp], PT = its transpose (kept for the backward), O_tok [BH][Tp][Fvp].  activation: 1 SELU, 0 none.
int mha_attention_forward(const float* q_tok, const float* k_tok, const float* v_chan, float* P, float* PT, float* o_tok,
                          int BH, int Tp, int Fqp, int Fvp, float scale, int activation, cudaStream_t st) {
  HNO_CHECK(q_tok && k_tok && v_chan && P && o_tok, "mha_attention_forward: null pointer");
  HNO_CHECK(activation == 0 || activation == 1, "mha_attention_forward: activation must be 0 (none) or 1 (SELU)");
  GemmArgs g1 = {};
  g1.a = q_tok, g1.lda = Fqp, g1.sa = (long)Tp * Fqp;
  g1.b = k_tok, g1.ldb = Fqp, g1.sb = (long)Tp * Fqp;
  g1.c = P, g1.ldc = Tp, g1.sc = (long)Tp * Tp;
  g1.ct = PT, g1.ldct = Tp, g1.sct = (long)Tp * Tp;
  g1.batch = BH, g1.M = Tp, g1.N = Tp, g1.K = Fqp;
  g1.alpha = scale, g1.epi = activation;
  if (int rc = gemm_tn(g1, st)) return rc;
  GemmArgs g2 = {};
  g2.a = P, g2.lda = Tp, g2.sa = (long)Tp * Tp;
  g2.b = v_chan, g2.ldb = Tp, g2.sb = (long)Fvp * Tp;
  g2.c = o_tok, g2.ldc = Fvp, g2.sc = (long)Tp * Fvp;
  g2.batch = BH, g2.M = Tp, g2.N = Fvp, g2.K = Tp;
  g2.alpha = 1.f, g2.epi = 0;
  return gemm_tn(g2, st);
}

// dS, dST: [BH][Tp][Tp] scratch.  dq_tok, dk_tok [BH][Tp][Fqp], dv_tok [BH][Tp][Fvp].
int mha_attention_backward(const float* do_tok, const float* do_chan, const float* q_chan, const float* k_chan,
                           const float* v_tok, const float* P, const float* PT, float* dS, float* dST, float* dq_tok,
                           float* dk_tok, float* dv_tok, int BH, int Tp, int Fqp, int Fvp, float scale, int activation,
                           cudaStream_t st) {
  HNO_CHECK(do_tok && do_chan && q_chan && k_chan && v_tok && P && PT && dS && dST && dq_tok && dk_tok && dv_tok,
            "mha_attention_backward: null pointer");
  const long tt = (long)Tp * Tp;
  GemmArgs g = {};
  // dP = dO V^T;  dS = dP * selu'(S / sqrt(F)) / sqrt(F), the derivative taken from the SELU output P
  g.a = do_tok, g.lda = Fvp, g.sa = (long)Tp * Fvp;
  g.b = v_tok, g.ldb = Fvp, g.sb = (long)Tp * Fvp;
  g.c = dS, g.ldc = Tp, g.sc = tt;
  g.ct = dST, g.ldct = Tp, g.sct = tt;
  g.e = activation == 1 ? P : nullptr, g.lde = Tp, g.se = tt;
  g.batch = BH, g.M = Tp, g.N = Tp, g.K = Fvp;
  g.alpha = scale, g.epi = activation == 1 ? 2 : 0;
  if (int rc = gemm_tn(g, st)) return rc;
  // dQ[q][f] = sum_k dS[q][k] K[k][f]
  g = GemmArgs{};
  g.a = dS, g.lda = Tp, g.sa = tt;
  g.b = k_chan, g.ldb = Tp, g.sb = (long)Fqp * Tp;
  g.c = dq_tok, g.ldc = Fqp, g.sc = (long)Tp * Fqp;
  g.batch = BH, g.M = Tp, g.N = Fqp, g.K = Tp;
  g.alpha = 1.f;
  if (int rc = gemm_tn(g, st)) return rc;
  // dK[k][f] = sum_q dS[q][k] Q[q][f]
  g.a = dST;
  g.b = q_chan;
  g.c = dk_tok;
  if (int rc = gemm_tn(g, st)) return rc;
  // dV[k][f] = sum_q P[q][k] dO[q][f]
  g.a = PT;
  g.b = do_chan, g.sb = (long)Fvp * Tp;
  g.c = dv_tok, g.ldc = Fvp, g.sc = (long)Tp * Fvp;
  g.N = Fvp;
  return gemm_tn(g, st);
}

int mha_output_forward(const float* o_tok, const float* wout, const float* bias, float* y, int B, int H, int co, int cd,
                       int Ld, int Lh, int Lw, int pd, int ph, int pw, int Tp, int Fp, cudaStream_t st) {
  MhaGeom g;
  if (make_geom(&g, B, H, Ld, Lh, Lw, pd, ph, pw, Tp)) return -1;
  HNO_CHECK(o_tok && wout && y, "mha_output_forward: null pointer");
  HNO_CHECK(co <= 65535 && B <= 65535, "mha_output_forward: too many channels");
  if (mha_tok_path(g, co, cd)) {
    const MhaW wso{(long)cd, 1, (long)H * cd};  // W(h, c, o) = weight_out[o][h cd + c]
    dim3 grid(ceil_div(g.T, 128), co / 4, B);
    k_mha_tok_to_modes<<<grid, 128, (size_t)H * cd * 4 * sizeof(float), st>>>(o_tok, wout, bias, y, g, wso, co, cd, Fp, 0);
  } else {
    dim3 grid(ceil_div(g.M, 256), co, B);
    k_mha_output_fwd<<<grid, 256, 0, st>>>(o_tok, wout, bias, y, g, co, cd, Fp);
  }
  HNO_LAUNCH_CHECK();
  return 0;
}

int mha_output_backward(const float* dy, const float* o_tok, const float* wout, float* do_tok, float* do_chan, float* dwout,
                        float* dbias, void* wsp, int B, int H, int co, int cd, int Ld, int Lh, int Lw, int pd, int ph, int pw,
                        int Tp, int Fp, cudaStream_t st) {
  MhaGeom g;
  if (make_geom(&g, B, H, Ld, Lh, Lw, pd, ph, pw, Tp)) return -1;
  HNO_CHECK(dy && o_tok && wout && do_tok && do_chan && dwout, "mha_output_backward: null pointer");
  HNO_CHECK(Fp >= cd * pd * ph * pw && Fp % 32 == 0, "mha_output_backward: bad feature pitch %d", Fp);
  HNO_CHECK(cd <= 65535 && co * sizeof(float) <= 48 * 1024, "mha_output_backward: too many channels");
  // dO = W_out^T dy in both attention layouts: the projection kernel with the weight read transposed
  const MhaW ws{(long)cd, 1, (long)H * cd};
  if (mha_tok_path(g, co, cd)) {
    dim3 grid(Tp / 128, B * H);
    k_mha_project_tok<<<grid, 128, (size_t)cd * co * sizeof(float), st>>>(dy, wout, nullptr, do_tok, do_chan, g, ws, co, cd, Fp);
  } else {
    dim3 grid(ceil_div(Tp, 256), cd, B * H);
    k_mha_project_fwd<<<grid, 256, co * sizeof(float), st>>>(dy, wout, nullptr, do_tok, do_chan, g, ws, co, cd, Fp);
  }
  HNO_LAUNCH_CHECK();
  // dW_out[o][h cd + c] = sum dy[b][o][m] O[b, h][t][f]; bias: sum of dy over the modes
  return mha_wgrad(o_tok, dy, dwout, dbias, wsp, g, ws, co, cd, Fp, dbias ? 2 : 0, st);
}

}  // namespace hno
