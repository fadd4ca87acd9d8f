// Deep-supervision convolution: out = act( sum_i W_i in_i + b ) over a LIST of activations, without ever forming their
// concatenation.  Replaces `torch.cat(tensors, 1)` + `conv_ds` of the reference (nets/architectures.py:295-311, 330-343:
// every block output -- 17 x 12 channels for HartleyMHASeg, 9 x 24 for HNOSeg-XS -- is concatenated and reduced to
// out_channels by a 1x1x1 ConvNormAct; nets/hnosegxs.py:110-125, 154-172 does the same for HNOSeg-XS).
// Both directions are single passes over the sources: HBM-bound streaming kernels (bytes = all sources once + the output).
#include "common.cuh"

namespace hno {

constexpr int kDsMaxSrc = 40;   // sources (blocks + 1)
constexpr int kDsMaxCo = 8;     // output channels
constexpr int kDsIts = 4;       // float4 groups per thread in the backward (16 voxels)

struct DsSrc {
  const float* in[kDsMaxSrc];
  float* din[kDsMaxSrc];
  int ch[kDsMaxSrc];
  int n;
  int ctot;
};

// grid (chunks, B); every thread owns 4 consecutive voxels.  w [CO][ctot] staged in shared memory.
// w2 [CO][CO] (optional): the bias-free conv_out that follows the deep-supervision head (nets/architectures.py:311-313,
// nets/hnosegxs.py:133-134 with in_channels = out_channels); out2 = w2 * out is written next to out in the same pass.
template <int CO>
__global__ void __launch_bounds__(256) k_dsconv_fwd(const DsSrc s, const float* __restrict__ w, const float* __restrict__ bias,
                                                    float* __restrict__ out, const float* __restrict__ w2,
                                                    float* __restrict__ out2, long S, int act) {
  extern __shared__ float sw[];
  for (int i = threadIdx.x; i < CO * s.ctot; i += 256) sw[i] = __ldg(w + i);
  __syncthreads();
  const int b = blockIdx.y;
  for (long v = (blockIdx.x * 256L + threadIdx.x) * 4; v < S; v += (long)gridDim.x * 1024) {
    float4 acc[CO];
#pragma unroll
    for (int o = 0; o < CO; ++o) {
      const float bo = bias ? __ldg(bias + o) : 0.f;
      acc[o] = make_float4(bo, bo, bo, bo);
    }
    int k = 0;
    for (int i = 0; i < s.n; ++i) {
      const float* p = s.in[i] + (long)b * s.ch[i] * S + v;
      for (int c = 0; c < s.ch[i]; ++c, ++k) {
        const float4 x = __ldg(reinterpret_cast<const float4*>(p + (long)c * S));
#pragma unroll
        for (int o = 0; o < CO; ++o) {
          const float wk = sw[o * s.ctot + k];
          acc[o].x = fmaf(wk, x.x, acc[o].x);
          acc[o].y = fmaf(wk, x.y, acc[o].y);
          acc[o].z = fmaf(wk, x.z, acc[o].z);
          acc[o].w = fmaf(wk, x.w, acc[o].w);
        }
      }
    }
#pragma unroll
    for (int o = 0; o < CO; ++o) {
      float4 r = acc[o];
      if (act == 1) r = make_float4(selu_f(r.x), selu_f(r.y), selu_f(r.z), selu_f(r.w));
      acc[o] = r;
      *reinterpret_cast<float4*>(out + ((long)b * CO + o) * S + v) = r;
    }
    if (w2) {
#pragma unroll
      for (int o = 0; o < CO; ++o) {
        float4 r = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int c = 0; c < CO; ++c) {
          const float wk = __ldg(w2 + o * CO + c);
          r.x = fmaf(wk, acc[c].x, r.x), r.y = fmaf(wk, acc[c].y, r.y);
          r.z = fmaf(wk, acc[c].z, r.z), r.w = fmaf(wk, acc[c].w, r.w);
        }
        *reinterpret_cast<float4*>(out2 + ((long)b * CO + o) * S + v) = r;
      }
    }
  }
}

// Backward: d(pre) = dy * selu'(y);  din_i = W_i^T d(pre);  partial dW / db per CTA -> partials [grid.x * B][CO * ctot + CO].
// `valid` = number of real voxels per plane row group: voxels with (v % P) >= HW are layout padding and carry no gradient.
// With w2: dy is the gradient of out2 = w2 * out; d(out) = w2^T dy and dw2 [CO][CO] = sum dy out^T joins the partials.
template <int CO>
__global__ void __launch_bounds__(256) k_dsconv_bwd(const DsSrc s, const float* __restrict__ w, const float* __restrict__ dy,
                                                    const float* __restrict__ y, const float* __restrict__ w2,
                                                    float* __restrict__ partials, long S, long P, long HW, int act) {
  extern __shared__ float sm[];
  float* sw = sm;                       // [CO][ctot]
  float* sacc = sm + CO * s.ctot;       // [CO][ctot] + [CO] + [CO][CO]   partial sums of this CTA
  const int nacc = CO * s.ctot + CO + CO * CO;
  for (int i = threadIdx.x; i < CO * s.ctot; i += 256) sw[i] = __ldg(w + i);
  for (int i = threadIdx.x; i < nacc; i += 256) sacc[i] = 0.f;
  __syncthreads();
  const int b = blockIdx.y;
  const int lane = threadIdx.x & 31;
  for (long v0 = (long)blockIdx.x * (1024 * kDsIts); v0 < S; v0 += (long)gridDim.x * (1024 * kDsIts)) {
    float4 dp[kDsIts][CO];
    bool live[kDsIts];
#pragma unroll
    for (int it = 0; it < kDsIts; ++it) {
      const long v = v0 + it * 1024 + threadIdx.x * 4;
      live[it] = v < S;
      float4 gy[CO], yy[CO];
      const long r = live[it] ? v % P : 0;
#pragma unroll
      for (int o = 0; o < CO; ++o) {
        gy[o] = make_float4(0.f, 0.f, 0.f, 0.f);
        yy[o] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (live[it]) {
          gy[o] = __ldg(reinterpret_cast<const float4*>(dy + ((long)b * CO + o) * S + v));
          if (act == 1 || w2) yy[o] = __ldg(reinterpret_cast<const float4*>(y + ((long)b * CO + o) * S + v));
          if (P != HW) {  // padding columns of the planar layout: no gradient
            if (r + 0 >= HW) gy[o].x = 0.f;
            if (r + 1 >= HW) gy[o].y = 0.f;
            if (r + 2 >= HW) gy[o].z = 0.f;
            if (r + 3 >= HW) gy[o].w = 0.f;
          }
        }
      }
      if (w2) {
        // dw2[o][c] += dy[o] . out[c]   (warp-reduced per tile: CO * CO values)
#pragma unroll
        for (int o = 0; o < CO; ++o)
#pragma unroll
          for (int c = 0; c < CO; ++c) {
            float t = fmaf(gy[o].x, yy[c].x, fmaf(gy[o].y, yy[c].y, fmaf(gy[o].z, yy[c].z, gy[o].w * yy[c].w)));
            t = warp_sum(t);
            if (lane == 0) atomicAdd(&sacc[CO * s.ctot + CO + o * CO + c], t);
          }
      }
#pragma unroll
      for (int o = 0; o < CO; ++o) {
        float4 g = gy[o];
        if (w2) {  // d(out)[o] = sum_c w2[c][o] dy[c]
          g = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
          for (int c = 0; c < CO; ++c) {
            const float wk = __ldg(w2 + c * CO + o);
            g.x = fmaf(wk, gy[c].x, g.x), g.y = fmaf(wk, gy[c].y, g.y);
            g.z = fmaf(wk, gy[c].z, g.z), g.w = fmaf(wk, gy[c].w, g.w);
          }
        }
        if (act == 1) {
          g.x *= selu_grad_from_out(yy[o].x), g.y *= selu_grad_from_out(yy[o].y);
          g.z *= selu_grad_from_out(yy[o].z), g.w *= selu_grad_from_out(yy[o].w);
        }
        dp[it][o] = g;
      }
    }
    // bias gradient
#pragma unroll
    for (int o = 0; o < CO; ++o) {
      float t = 0.f;
#pragma unroll
      for (int it = 0; it < kDsIts; ++it) t += (dp[it][o].x + dp[it][o].y) + (dp[it][o].z + dp[it][o].w);
      t = warp_sum(t);
      if (lane == 0) atomicAdd(&sacc[CO * s.ctot + o], t);
    }
    int k = 0;
    for (int i = 0; i < s.n; ++i) {
      const float* p = s.in[i] + (long)b * s.ch[i] * S;
      float* q = s.din[i] ? s.din[i] + (long)b * s.ch[i] * S : nullptr;
      for (int c = 0; c < s.ch[i]; ++c, ++k) {
        float part[CO];
#pragma unroll
        for (int o = 0; o < CO; ++o) part[o] = 0.f;
#pragma unroll
        for (int it = 0; it < kDsIts; ++it) {
          if (!live[it]) continue;
          const long v = v0 + it * 1024 + threadIdx.x * 4;
          const float4 x = __ldg(reinterpret_cast<const float4*>(p + (long)c * S + v));
          float4 gin = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
          for (int o = 0; o < CO; ++o) {
            const float4 g = dp[it][o];
            part[o] = fmaf(g.x, x.x, fmaf(g.y, x.y, fmaf(g.z, x.z, fmaf(g.w, x.w, part[o]))));
            const float wk = sw[o * s.ctot + k];
            gin.x = fmaf(wk, g.x, gin.x), gin.y = fmaf(wk, g.y, gin.y);
            gin.z = fmaf(wk, g.z, gin.z), gin.w = fmaf(wk, g.w, gin.w);
          }
          if (q) *reinterpret_cast<float4*>(q + (long)c * S + v) = gin;
        }
#pragma unroll
        for (int o = 0; o < CO; ++o) {
          const float t = warp_sum(part[o]);
          if (lane == 0) atomicAdd(&sacc[o * s.ctot + k], t);
        }
      }
    }
  }
  __syncthreads();
  float* dst = partials + ((long)blockIdx.y * gridDim.x + blockIdx.x) * nacc;
  for (int i = threadIdx.x; i < nacc; i += 256) dst[i] = sacc[i];
}

// dw [CO][ctot], db [CO], dw2 [CO][CO] = sum over the CTA partials (fp64)
__global__ void __launch_bounds__(128) k_dsconv_reduce(const float* __restrict__ partials, int nparts, int n, int nw, int nb,
                                                       float* __restrict__ dw, float* __restrict__ db,
                                                       float* __restrict__ dw2) {
  const int i = blockIdx.x * 128 + threadIdx.x;
  if (i >= n) return;
  double t = 0.0;
  for (int p = 0; p < nparts; ++p) t += (double)partials[(long)p * n + i];
  if (i < nw) dw[i] = (float)t;
  else if (i < nw + nb) {
    if (db) db[i - nw] = (float)t;
  } else if (dw2) dw2[i - nw - nb] = (float)t;
}

static int ds_pack(DsSrc* s, const float* const* in, float* const* din, const int* ch, int n) {
  HNO_CHECK(n >= 1 && n <= kDsMaxSrc, "dsconv: %d sources (supported: 1..%d)", n, kDsMaxSrc);
  s->n = n;
  s->ctot = 0;
  for (int i = 0; i < n; ++i) {
    HNO_CHECK(in[i] && ch[i] >= 1, "dsconv: null source / bad channel count");
    HNO_CHECK(reinterpret_cast<uintptr_t>(in[i]) % 16 == 0, "dsconv: sources must be 16-byte aligned");
    s->in[i] = in[i];
    s->din[i] = din ? din[i] : nullptr;
    s->ch[i] = ch[i];
    s->ctot += ch[i];
  }
  for (int i = n; i < kDsMaxSrc; ++i) s->in[i] = nullptr, s->din[i] = nullptr, s->ch[i] = 0;
  return 0;
}

static int ds_grid(long S, int per_thread) {
  long g = (S + 256L * per_thread - 1) / (256L * per_thread);
  const long cap = (long)sm_count() * 8;
  return (int)(g < cap ? g : cap);
}

#define HNO_DS_CO_SWITCH(CO, ...)                                          \
  switch (CO) {                                                            \
    case 1: { constexpr int kCO = 1; __VA_ARGS__ } break;                  \
    case 2: { constexpr int kCO = 2; __VA_ARGS__ } break;                  \
    case 3: { constexpr int kCO = 3; __VA_ARGS__ } break;                  \
    case 4: { constexpr int kCO = 4; __VA_ARGS__ } break;                  \
    case 5: { constexpr int kCO = 5; __VA_ARGS__ } break;                  \
    case 6: { constexpr int kCO = 6; __VA_ARGS__ } break;                  \
    case 7: { constexpr int kCO = 7; __VA_ARGS__ } break;                  \
    case 8: { constexpr int kCO = 8; __VA_ARGS__ } break;                  \
    default: set_error("dsconv: %d output channels (supported: 1..%d)", CO, kDsMaxCo); return -1; \
  }

int dsconv_forward(const float* const* in, const int* ch, int n, const float* w, const float* bias, float* out,
                   const float* w2, float* out2, int B, int CO, long S, int act, cudaStream_t st) {
  DsSrc s;
  if (ds_pack(&s, in, nullptr, ch, n)) return -1;
  HNO_CHECK(w && out && B >= 1 && B <= 65535 && S >= 4 && S % 4 == 0, "dsconv_forward: bad arguments (S must be a multiple of 4)");
  HNO_CHECK(reinterpret_cast<uintptr_t>(out) % 16 == 0, "dsconv_forward: out must be 16-byte aligned");
  const size_t smem = (size_t)CO * s.ctot * sizeof(float);
  HNO_CHECK(smem <= 48 * 1024, "dsconv_forward: %d x %d weights do not fit shared memory", CO, s.ctot);
  dim3 grid(ds_grid(S, 4), B);
  HNO_CHECK((w2 == nullptr) == (out2 == nullptr), "dsconv_forward: w2 and out2 go together");
  HNO_CHECK(!out2 || reinterpret_cast<uintptr_t>(out2) % 16 == 0, "dsconv_forward: out2 must be 16-byte aligned");
  HNO_DS_CO_SWITCH(CO, { k_dsconv_fwd<kCO><<<grid, 256, smem, st>>>(s, w, bias, out, w2, out2, S, act); })
  HNO_LAUNCH_CHECK();
  return 0;
}

size_t dsconv_backward_workspace_bytes(int ctot, int CO, int B, long S) {
  const long parts = (long)ds_grid(S, 4 * kDsIts) * B;
  return (size_t)parts * ((size_t)CO * ctot + CO + CO * CO) * sizeof(float) + 256;
}

int dsconv_backward(const float* const* in, float* const* din, const int* ch, int n, const float* w, const float* dy,
                    const float* y, const float* w2, float* dw, float* db, float* dw2, void* ws, int B, int CO, long S,
                    long P, long HW, int act, cudaStream_t st) {
  DsSrc s;
  if (ds_pack(&s, in, din, ch, n)) return -1;
  HNO_CHECK(w && dy && dw && ws && ((act == 0 && !w2) || y), "dsconv_backward: null pointer");
  HNO_CHECK((w2 == nullptr) == (dw2 == nullptr), "dsconv_backward: w2 and dw2 go together");
  HNO_CHECK(B >= 1 && B <= 65535 && S >= 4 && S % 4 == 0 && P >= HW && P >= 1 && S % P == 0,
            "dsconv_backward: bad sizes (S must be a multiple of 4 and of the plane pitch)");
  const size_t smem = ((size_t)2 * CO * s.ctot + CO + CO * CO) * sizeof(float);
  HNO_CHECK(smem <= 96 * 1024, "dsconv_backward: %d x %d weights do not fit shared memory", CO, s.ctot);
  dim3 grid(ds_grid(S, 4 * kDsIts), B);
  float* partials = reinterpret_cast<float*>(ws);
  HNO_DS_CO_SWITCH(CO, {
    auto kern = k_dsconv_bwd<kCO>;
    HNO_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kern<<<grid, 256, smem, st>>>(s, w, dy, y, w2, partials, S, P, HW, act);
  })
  HNO_LAUNCH_CHECK();
  const int nw = CO * s.ctot, nall = nw + CO + CO * CO;
  k_dsconv_reduce<<<ceil_div(nall, 128), 128, 0, st>>>(partials, (int)(grid.x * grid.y), nall, nw, CO, dw, db, dw2);
  HNO_LAUNCH_CHECK();
  return 0;
}

}  // namespace hno
