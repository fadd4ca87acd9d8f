// Per-mode ("individual" weights) Hartley mixing and the fused Adamax step, sm_100a.
//
// Replaces nets/hartley_operator.py:293-299 (the weights_type == 'individual' branch of
// _call3d_notransform), :302-317 (hartley_conv) and :320-333 (get_reverse = flip + roll by one, i.e.
// index j -> (n - j) mod n on each of the three mode axes of the CROPPED block):
//     out(k) = 1/2 [ W(k) (X(k) + X(~k)) + W(~k) (X(k) - X(~k)) ]        summed over input channels
// The op is bound by the weight read (CO*CI*M floats, 36 MB at 24x24x15,680 modes): every thread owns
// one mode k, streams W(k) and W(~k) once (both coalesced: ~k runs backwards through memory) and
// keeps the B x CO outputs in registers.
#include "common.cuh"
#include "hno_b200.h"

namespace hno {

__device__ __forceinline__ long rev_index(long k, int n0, int n1, int n2) {
  const int k2 = (int)(k % n2);
  const long r = k / n2;
  const int k1 = (int)(r % n1);
  const int k0 = (int)(r / n1);
  const int r0 = k0 == 0 ? 0 : n0 - k0, r1 = k1 == 0 ? 0 : n1 - k1, r2 = k2 == 0 ? 0 : n2 - k2;
  return ((long)r0 * n1 + r1) * n2 + r2;
}

constexpr int kHcBatch = 2;  // samples per register tile

template <int CO>
__global__ void __launch_bounds__(128) k_hartley_conv_fwd(const float* __restrict__ x, const float* __restrict__ w,
                                                          float* __restrict__ out, int B, int CI, int n0, int n1,
                                                          int n2, int residual_selu) {
  const long M = (long)n0 * n1 * n2;
  const long k = blockIdx.x * 128L + threadIdx.x;
  if (k >= M) return;
  const long kr = rev_index(k, n0, n1, n2);
  for (int b0 = 0; b0 < B; b0 += kHcBatch) {
    float acc[kHcBatch][CO];
#pragma unroll
    for (int bb = 0; bb < kHcBatch; ++bb)
#pragma unroll
      for (int o = 0; o < CO; ++o) acc[bb][o] = 0.f;
    for (int i = 0; i < CI; ++i) {
      float e[kHcBatch], od[kHcBatch];
#pragma unroll
      for (int bb = 0; bb < kHcBatch; ++bb) {
        const int b = min(b0 + bb, B - 1);
        const float a = __ldg(x + ((long)b * CI + i) * M + k), c = __ldg(x + ((long)b * CI + i) * M + kr);
        e[bb] = a + c;
        od[bb] = a - c;
      }
#pragma unroll
      for (int o = 0; o < CO; ++o) {
        const float wk = __ldg(w + ((long)o * CI + i) * M + k), wr = __ldg(w + ((long)o * CI + i) * M + kr);
#pragma unroll
        for (int bb = 0; bb < kHcBatch; ++bb) acc[bb][o] = fmaf(wk, e[bb], fmaf(wr, od[bb], acc[bb][o]));
      }
    }
#pragma unroll
    for (int bb = 0; bb < kHcBatch; ++bb)
      if (b0 + bb < B) {
#pragma unroll
        for (int o = 0; o < CO; ++o) {
          float v = 0.5f * acc[bb][o];
          if (residual_selu) v = selu_f(v + __ldg(x + ((long)(b0 + bb) * CI + o) * M + k));  // CI == CO here
          out[((long)(b0 + bb) * CO + o) * M + k] = v;
        }
      }
  }
}

// dX(k) = 1/2 sum_o [ W(k) (g(k) - g(~k)) + W(~k) (g(k) + g(~k)) ]
template <int CI>
__global__ void __launch_bounds__(128) k_hartley_conv_bwd_x(const float* __restrict__ g, const float* __restrict__ y,
                                                            const float* __restrict__ w, float* __restrict__ dx, int B,
                                                            int CO, int n0, int n1, int n2) {
  const long M = (long)n0 * n1 * n2;
  const long k = blockIdx.x * 128L + threadIdx.x;
  if (k >= M) return;
  const long kr = rev_index(k, n0, n1, n2);
  for (int b0 = 0; b0 < B; b0 += kHcBatch) {
    float acc[kHcBatch][CI];
#pragma unroll
    for (int bb = 0; bb < kHcBatch; ++bb)
#pragma unroll
      for (int i = 0; i < CI; ++i) acc[bb][i] = 0.f;
    for (int o = 0; o < CO; ++o) {
      float dm[kHcBatch], dp[kHcBatch];
#pragma unroll
      for (int bb = 0; bb < kHcBatch; ++bb) {
        const int b = min(b0 + bb, B - 1);
        float a = __ldg(g + ((long)b * CO + o) * M + k), c = __ldg(g + ((long)b * CO + o) * M + kr);
        if (y != nullptr) {  // fused selu(mix + x): g is the gradient of the activated output
          a *= selu_grad_from_out(__ldg(y + ((long)b * CO + o) * M + k));
          c *= selu_grad_from_out(__ldg(y + ((long)b * CO + o) * M + kr));
        }
        dm[bb] = a - c;
        dp[bb] = a + c;
      }
#pragma unroll
      for (int i = 0; i < CI; ++i) {
        const float wk = __ldg(w + ((long)o * CI + i) * M + k), wr = __ldg(w + ((long)o * CI + i) * M + kr);
#pragma unroll
        for (int bb = 0; bb < kHcBatch; ++bb) acc[bb][i] = fmaf(wk, dm[bb], fmaf(wr, dp[bb], acc[bb][i]));
      }
    }
#pragma unroll
    for (int bb = 0; bb < kHcBatch; ++bb)
      if (b0 + bb < B) {
#pragma unroll
        for (int i = 0; i < CI; ++i) {
          float v = 0.5f * acc[bb][i];
          if (y != nullptr) {  // residual path (CI == CO)
            const long off = ((long)(b0 + bb) * CO + i) * M + k;
            v = fmaf(__ldg(g + off), selu_grad_from_out(__ldg(y + off)), v);
          }
          dx[((long)(b0 + bb) * CI + i) * M + k] = v;
        }
      }
  }
}

// dW(k)[o][i] = 1/2 sum_b [ g_o(k) (X_i(k) + X_i(~k)) + g_o(~k) (X_i(~k) - X_i(k)) ]
__global__ void __launch_bounds__(128) k_hartley_conv_bwd_w(const float* __restrict__ g, const float* __restrict__ y,
                                                            const float* __restrict__ x, float* __restrict__ dw, int B,
                                                            int CI, int CO, int n0, int n1, int n2, int accumulate) {
  const long M = (long)n0 * n1 * n2;
  const long k = blockIdx.x * 128L + threadIdx.x;
  if (k >= M) return;
  const int o = blockIdx.y;
  const long kr = rev_index(k, n0, n1, n2);
  for (int i = 0; i < CI; ++i) {
    float acc = 0.f;
    for (int b = 0; b < B; ++b) {
      float gk = __ldg(g + ((long)b * CO + o) * M + k), gr = __ldg(g + ((long)b * CO + o) * M + kr);
      if (y != nullptr) {
        gk *= selu_grad_from_out(__ldg(y + ((long)b * CO + o) * M + k));
        gr *= selu_grad_from_out(__ldg(y + ((long)b * CO + o) * M + kr));
      }
      const float xk = __ldg(x + ((long)b * CI + i) * M + k), xr = __ldg(x + ((long)b * CI + i) * M + kr);
      acc = fmaf(gk, xk + xr, fmaf(gr, xr - xk, acc));
    }
    float* dst = dw + ((long)o * CI + i) * M + k;
    *dst = accumulate ? *dst + 0.5f * acc : 0.5f * acc;
  }
}

#define HNO_HC_CHANNELS(X) X(8) X(16) X(24) X(32)

int hartley_conv_forward(const float* x, const float* w, float* out, int B, int ci, int co, int n0, int n1, int n2,
                         int residual_selu, cudaStream_t st) {
  HNO_CHECK(x && w && out, "hartley_conv_forward: null pointer");
  HNO_CHECK(!residual_selu || ci == co, "hartley_conv_forward: the fused residual needs in_channels == out_channels");
  const long M = (long)n0 * n1 * n2;
  const int grid = ceil_div(M, 128);
#define X(C)                                                                   \
  if (co == C) {                                                               \
    k_hartley_conv_fwd<C><<<grid, 128, 0, st>>>(x, w, out, B, ci, n0, n1, n2, residual_selu); \
    HNO_LAUNCH_CHECK();                                                        \
    return 0;                                                                  \
  }
  HNO_HC_CHANNELS(X)
#undef X
  set_error("hartley_conv_forward: out_channels %d not in {8,16,24,32}", co);
  return -1;
}

int hartley_conv_backward(const float* dout, const float* y, const float* x, const float* w, float* dx, float* dw,
                          int B, int ci, int co, int n0, int n1, int n2, int accumulate_dw, cudaStream_t st) {
  HNO_CHECK(dout && x && w, "hartley_conv_backward: null pointer");
  HNO_CHECK(y == nullptr || ci == co, "hartley_conv_backward: the fused residual needs in_channels == out_channels");
  const long M = (long)n0 * n1 * n2;
  const int grid = ceil_div(M, 128);
  if (dx) {
    bool done = false;
#define X(C)                                                                        \
  if (ci == C) {                                                                    \
    k_hartley_conv_bwd_x<C><<<grid, 128, 0, st>>>(dout, y, w, dx, B, co, n0, n1, n2);  \
    done = true;                                                                    \
  }
    HNO_HC_CHANNELS(X)
#undef X
    HNO_CHECK(done, "hartley_conv_backward: in_channels %d not in {8,16,24,32}", ci);
    HNO_LAUNCH_CHECK();
  }
  if (dw) {
    dim3 g2(grid, co);
    k_hartley_conv_bwd_w<<<g2, 128, 0, st>>>(dout, y, x, dw, B, ci, co, n0, n1, n2, accumulate_dw);
    HNO_LAUNCH_CHECK();
  }
  return 0;
}

// ------------------------------------------------------------------------------------------ full-spectrum reversal
// HartleyOperator(use_transform=True, weights_type='individual') (nets/hartley_operator.py:196-241): the reversal partner
// of X is taken in the FULL N-point spectrum, x_reverse = get_reverse(dht3(x)) (:199), and only afterwards cropped, while
// the weight is reversed inside its own 2m-sized block (:200).  The partner of the retained frequency n - m + t is m - t;
// for t = 0 that is +m, which lies OUTSIDE the retained corner set -- so the truncated transform here also produces
// frequency +m per axis ("extended" mode tensor x_ext [B][CI][E0][E1][E2], E = 2m + 1, or 2m when n == 2m) and
//     out(j) = act( 1/2 [ W(j) (X(j) + X(r(j))) + W(~j) (X(j) - X(r(j))) ] ),
// j over the retained block [n0][n1][n2] (position j of the block = position j of x_ext), r(j) = position of the partner in
// x_ext, per axis from `rtab` = [r0 (n0 ints) | r1 (n1) | r2 (n2)], ~j = (2m - j) mod 2m.  Same thread layout as above.
struct HcFull {
  int n0, n1, n2, e0, e1, e2;
  const int* r0;
  const int* r1;
  const int* r2;
};

__device__ __forceinline__ void hc_full_locate(const HcFull& t, long k, long& kx, long& kxr, long& kwr) {
  const int k2 = (int)(k % t.n2);
  const long r = k / t.n2;
  const int k1 = (int)(r % t.n1);
  const int k0 = (int)(r / t.n1);
  kx = ((long)k0 * t.e1 + k1) * t.e2 + k2;
  kxr = ((long)__ldg(t.r0 + k0) * t.e1 + __ldg(t.r1 + k1)) * t.e2 + __ldg(t.r2 + k2);
  const int w0 = k0 == 0 ? 0 : t.n0 - k0, w1 = k1 == 0 ? 0 : t.n1 - k1, w2 = k2 == 0 ? 0 : t.n2 - k2;
  kwr = ((long)w0 * t.n1 + w1) * t.n2 + w2;
}

template <int CO>
__global__ void __launch_bounds__(128) k_hconv_full_fwd(const float* __restrict__ x, const float* __restrict__ w,
                                                        float* __restrict__ out, int B, int CI, HcFull t, int act) {
  const long M = (long)t.n0 * t.n1 * t.n2, E = (long)t.e0 * t.e1 * t.e2;
  const long k = blockIdx.x * 128L + threadIdx.x;
  if (k >= M) return;
  long kx, kxr, kwr;
  hc_full_locate(t, k, kx, kxr, kwr);
  for (int b = 0; b < B; ++b) {
    float acc[CO];
#pragma unroll
    for (int o = 0; o < CO; ++o) acc[o] = 0.f;
    for (int i = 0; i < CI; ++i) {
      const float a = __ldg(x + ((long)b * CI + i) * E + kx), c = __ldg(x + ((long)b * CI + i) * E + kxr);
      const float e = a + c, od = a - c;
#pragma unroll
      for (int o = 0; o < CO; ++o)
        acc[o] = fmaf(__ldg(w + ((long)o * CI + i) * M + k), e, fmaf(__ldg(w + ((long)o * CI + i) * M + kwr), od, acc[o]));
    }
#pragma unroll
    for (int o = 0; o < CO; ++o) {
      const float v = 0.5f * acc[o];
      out[((long)b * CO + o) * M + k] = act ? selu_f(v) : v;
    }
  }
}

// dX scatter: thread j adds 1/2 (W(j) + W(~j))^T d(j) at position j and 1/2 (W(j) - W(~j))^T d(j) at position r(j) of the
// zero-initialised dx_ext.  Every element receives at most two contributions (its own and its partner's), so the
// floating-point sum does not depend on the order of the atomics.
template <int CI>
__global__ void __launch_bounds__(128) k_hconv_full_bwd_x(const float* __restrict__ g, const float* __restrict__ y,
                                                          const float* __restrict__ w, float* __restrict__ dx, int B,
                                                          int CO, HcFull t) {
  const long M = (long)t.n0 * t.n1 * t.n2, E = (long)t.e0 * t.e1 * t.e2;
  const long k = blockIdx.x * 128L + threadIdx.x;
  if (k >= M) return;
  long kx, kxr, kwr;
  hc_full_locate(t, k, kx, kxr, kwr);
  for (int b = 0; b < B; ++b) {
    float own[CI], par[CI];
#pragma unroll
    for (int i = 0; i < CI; ++i) own[i] = par[i] = 0.f;
    for (int o = 0; o < CO; ++o) {
      float d = __ldg(g + ((long)b * CO + o) * M + k);
      if (y != nullptr) d *= selu_grad_from_out(__ldg(y + ((long)b * CO + o) * M + k));
#pragma unroll
      for (int i = 0; i < CI; ++i) {
        const float wk = __ldg(w + ((long)o * CI + i) * M + k), wr = __ldg(w + ((long)o * CI + i) * M + kwr);
        own[i] = fmaf(wk + wr, d, own[i]);
        par[i] = fmaf(wk - wr, d, par[i]);
      }
    }
#pragma unroll
    for (int i = 0; i < CI; ++i) {
      atomicAdd(dx + ((long)b * CI + i) * E + kx, 0.5f * own[i]);
      atomicAdd(dx + ((long)b * CI + i) * E + kxr, 0.5f * par[i]);
    }
  }
}

// dW(u)[o][i] = 1/2 sum_b [ d_o(u) (X_i(u) + X_i(r(u))) + d_o(~u) (X_i(~u) - X_i(r(~u))) ]
__global__ void __launch_bounds__(128) k_hconv_full_bwd_w(const float* __restrict__ g, const float* __restrict__ y,
                                                          const float* __restrict__ x, float* __restrict__ dw, int B,
                                                          int CI, int CO, HcFull t) {
  const long M = (long)t.n0 * t.n1 * t.n2, E = (long)t.e0 * t.e1 * t.e2;
  const long k = blockIdx.x * 128L + threadIdx.x;
  if (k >= M) return;
  const int o = blockIdx.y;
  long kx, kxr, kwr, vx, vxr, vwr;
  hc_full_locate(t, k, kx, kxr, kwr);
  hc_full_locate(t, kwr, vx, vxr, vwr);  // the mode whose ~ is k
  for (int i = 0; i < CI; ++i) {
    float acc = 0.f;
    for (int b = 0; b < B; ++b) {
      float dk = __ldg(g + ((long)b * CO + o) * M + k), dv = __ldg(g + ((long)b * CO + o) * M + kwr);
      if (y != nullptr) {
        dk *= selu_grad_from_out(__ldg(y + ((long)b * CO + o) * M + k));
        dv *= selu_grad_from_out(__ldg(y + ((long)b * CO + o) * M + kwr));
      }
      const float* xb = x + ((long)b * CI + i) * E;
      acc = fmaf(dk, __ldg(xb + kx) + __ldg(xb + kxr), fmaf(dv, __ldg(xb + vx) - __ldg(xb + vxr), acc));
    }
    dw[((long)o * CI + i) * M + k] = 0.5f * acc;
  }
}

static int hc_full_make(HcFull* t, const int* rtab, int n0, int n1, int n2, int e0, int e1, int e2) {
  HNO_CHECK(rtab, "hartley_conv_full: null partner table");
  HNO_CHECK(n0 >= 1 && n1 >= 1 && n2 >= 1 && e0 >= n0 && e1 >= n1 && e2 >= n2, "hartley_conv_full: bad sizes");
  t->n0 = n0, t->n1 = n1, t->n2 = n2, t->e0 = e0, t->e1 = e1, t->e2 = e2;
  t->r0 = rtab, t->r1 = rtab + n0, t->r2 = rtab + n0 + n1;
  return 0;
}

int hartley_conv_full_forward(const float* x_ext, const float* w, const int* rtab, float* out, int B, int ci, int co, int n0,
                              int n1, int n2, int e0, int e1, int e2, int act, cudaStream_t st) {
  HcFull t;
  if (hc_full_make(&t, rtab, n0, n1, n2, e0, e1, e2)) return -1;
  HNO_CHECK(x_ext && w && out, "hartley_conv_full_forward: null pointer");
  const int grid = ceil_div((long)n0 * n1 * n2, 128);
#define X(C)                                                                  \
  if (co == C) {                                                              \
    k_hconv_full_fwd<C><<<grid, 128, 0, st>>>(x_ext, w, out, B, ci, t, act);  \
    HNO_LAUNCH_CHECK();                                                       \
    return 0;                                                                 \
  }
  HNO_HC_CHANNELS(X)
#undef X
  set_error("hartley_conv_full_forward: out_channels %d not in {8,16,24,32}", co);
  return -1;
}

int hartley_conv_full_backward(const float* dout, const float* y, const float* x_ext, const float* w, const int* rtab,
                               float* dx_ext, float* dw, int B, int ci, int co, int n0, int n1, int n2, int e0, int e1, int e2,
                               cudaStream_t st) {
  HcFull t;
  if (hc_full_make(&t, rtab, n0, n1, n2, e0, e1, e2)) return -1;
  HNO_CHECK(dout && x_ext && w, "hartley_conv_full_backward: null pointer");
  const int grid = ceil_div((long)n0 * n1 * n2, 128);
  if (dx_ext) {
    HNO_CUDA(cudaMemsetAsync(dx_ext, 0, (size_t)B * ci * e0 * e1 * e2 * sizeof(float), st));
    bool done = false;
#define X(C)                                                                       \
  if (ci == C) {                                                                   \
    k_hconv_full_bwd_x<C><<<grid, 128, 0, st>>>(dout, y, w, dx_ext, B, co, t);     \
    done = true;                                                                   \
  }
    HNO_HC_CHANNELS(X)
#undef X
    HNO_CHECK(done, "hartley_conv_full_backward: in_channels %d not in {8,16,24,32}", ci);
    HNO_LAUNCH_CHECK();
  }
  if (dw) {
    dim3 g2(grid, co);
    k_hconv_full_bwd_w<<<g2, 128, 0, st>>>(dout, y, x_ext, dw, B, ci, co, t);
    HNO_LAUNCH_CHECK();
  }
  return 0;
}

// ------------------------------------------------------------------------------------------ complex per-mode mixing
// Replaces the weights_type == 'individual' branch of FourierOperator._call3d (nets/fourier_operator.py:165-187, four
// corner einsums 'oidhw,bidhw->bodhw' with weight = complex(weight_real, weight_imag)) on the retained half-spectrum
// held as separate real / imaginary mode tensors [B][C][M] (M = 2 m0 * 2 m1 * m2, the weights' own index order):
//     a + i b = (wr + i wi)(re + i im)   per mode, summed over input channels.
// Bound by the weight read (2 CO CI M floats): a thread owns one (o, k) pair, k fastest, so wr / wi stream through once,
// coalesced; the activations (a few hundred KB) are re-read from L1 / L2.  Samples in register tiles of kCmBatch.
constexpr int kCmBatch = 4;

__global__ void __launch_bounds__(128) k_cmix_fwd(const float* __restrict__ re, const float* __restrict__ im,
                                                  const float* __restrict__ wr, const float* __restrict__ wi,
                                                  float* __restrict__ a, float* __restrict__ b, int B, int CI, int CO,
                                                  long M) {
  const long k = blockIdx.x * 128L + threadIdx.x;
  const int o = blockIdx.y;
  if (k >= M) return;
  for (int b0 = 0; b0 < B; b0 += kCmBatch) {
    float aa[kCmBatch], ab[kCmBatch];
#pragma unroll
    for (int bb = 0; bb < kCmBatch; ++bb) aa[bb] = ab[bb] = 0.f;
    for (int i = 0; i < CI; ++i) {
      const float r = __ldg(wr + ((long)o * CI + i) * M + k), q = __ldg(wi + ((long)o * CI + i) * M + k);
#pragma unroll
      for (int bb = 0; bb < kCmBatch; ++bb) {
        if (b0 + bb < B) {
          const float xr = __ldg(re + ((long)(b0 + bb) * CI + i) * M + k);
          const float xi = __ldg(im + ((long)(b0 + bb) * CI + i) * M + k);
          aa[bb] = fmaf(r, xr, fmaf(-q, xi, aa[bb]));
          ab[bb] = fmaf(q, xr, fmaf(r, xi, ab[bb]));
        }
      }
    }
#pragma unroll
    for (int bb = 0; bb < kCmBatch; ++bb) {
      if (b0 + bb < B) {
        a[((long)(b0 + bb) * CO + o) * M + k] = aa[bb];
        b[((long)(b0 + bb) * CO + o) * M + k] = ab[bb];
      }
    }
  }
}

// d re = sum_o (wr da + wi db),  d im = sum_o (-wi da + wr db): a thread owns one (i, k) pair
__global__ void __launch_bounds__(128) k_cmix_bwd_x(const float* __restrict__ da, const float* __restrict__ db,
                                                    const float* __restrict__ wr, const float* __restrict__ wi,
                                                    float* __restrict__ dre, float* __restrict__ dim_, int B, int CI,
                                                    int CO, long M) {
  const long k = blockIdx.x * 128L + threadIdx.x;
  const int i = blockIdx.y;
  if (k >= M) return;
  for (int b0 = 0; b0 < B; b0 += kCmBatch) {
    float gr[kCmBatch], gi[kCmBatch];
#pragma unroll
    for (int bb = 0; bb < kCmBatch; ++bb) gr[bb] = gi[bb] = 0.f;
    for (int o = 0; o < CO; ++o) {
      const float r = __ldg(wr + ((long)o * CI + i) * M + k), q = __ldg(wi + ((long)o * CI + i) * M + k);
#pragma unroll
      for (int bb = 0; bb < kCmBatch; ++bb) {
        if (b0 + bb < B) {
          const float ga = __ldg(da + ((long)(b0 + bb) * CO + o) * M + k);
          const float gb = __ldg(db + ((long)(b0 + bb) * CO + o) * M + k);
          gr[bb] = fmaf(r, ga, fmaf(q, gb, gr[bb]));
          gi[bb] = fmaf(-q, ga, fmaf(r, gb, gi[bb]));
        }
      }
    }
#pragma unroll
    for (int bb = 0; bb < kCmBatch; ++bb) {
      if (b0 + bb < B) {
        dre[((long)(b0 + bb) * CI + i) * M + k] = gr[bb];
        dim_[((long)(b0 + bb) * CI + i) * M + k] = gi[bb];
      }
    }
  }
}

// d wr = sum_b (da re + db im),  d wi = sum_b (-da im + db re): a thread owns one (o, i, k) triple; grid (M/128, CI, CO)
__global__ void __launch_bounds__(128) k_cmix_bwd_w(const float* __restrict__ da, const float* __restrict__ db,
                                                    const float* __restrict__ re, const float* __restrict__ im,
                                                    float* __restrict__ dwr, float* __restrict__ dwi, int B, int CI,
                                                    int CO, long M, int accumulate) {
  const long k = blockIdx.x * 128L + threadIdx.x;
  const int i = blockIdx.y, o = blockIdx.z;
  if (k >= M) return;
  float gr = 0.f, gi = 0.f;
  for (int bb = 0; bb < B; ++bb) {
    const float ga = __ldg(da + ((long)bb * CO + o) * M + k), gb = __ldg(db + ((long)bb * CO + o) * M + k);
    const float xr = __ldg(re + ((long)bb * CI + i) * M + k), xi = __ldg(im + ((long)bb * CI + i) * M + k);
    gr = fmaf(ga, xr, fmaf(gb, xi, gr));
    gi = fmaf(-ga, xi, fmaf(gb, xr, gi));
  }
  const long idx = ((long)o * CI + i) * M + k;
  dwr[idx] = accumulate ? dwr[idx] + gr : gr;
  dwi[idx] = accumulate ? dwi[idx] + gi : gi;
}

int complex_modemix_forward(const float* re, const float* im, const float* wr, const float* wi, float* a, float* b,
                            int B, int ci, int co, long M, cudaStream_t st) {
  HNO_CHECK(re && im && wr && wi && a && b, "complex_modemix_forward: null pointer");
  HNO_CHECK(B >= 1 && ci >= 1 && co >= 1 && co <= 65535 && M >= 1, "complex_modemix_forward: bad sizes");
  dim3 grid(ceil_div(M, 128), co);
  k_cmix_fwd<<<grid, 128, 0, st>>>(re, im, wr, wi, a, b, B, ci, co, M);
  HNO_LAUNCH_CHECK();
  return 0;
}

int complex_modemix_backward(const float* da, const float* db, const float* re, const float* im, const float* wr,
                             const float* wi, float* dre, float* dim_, float* dwr, float* dwi, int B, int ci, int co,
                             long M, int accumulate_dw, cudaStream_t st) {
  HNO_CHECK(da && db && re && im && wr && wi, "complex_modemix_backward: null pointer");
  HNO_CHECK((dre == nullptr) == (dim_ == nullptr) && (dwr == nullptr) == (dwi == nullptr),
            "complex_modemix_backward: gradients come in (real, imaginary) pairs");
  HNO_CHECK(B >= 1 && ci >= 1 && co >= 1 && ci <= 65535 && co <= 65535 && M >= 1,
            "complex_modemix_backward: bad sizes");
  if (dre) {
    dim3 grid(ceil_div(M, 128), ci);
    k_cmix_bwd_x<<<grid, 128, 0, st>>>(da, db, wr, wi, dre, dim_, B, ci, co, M);
    HNO_LAUNCH_CHECK();
  }
  if (dwr) {
    dim3 grid(ceil_div(M, 128), ci, co);
    k_cmix_bwd_w<<<grid, 128, 0, st>>>(da, db, re, im, dwr, dwi, B, ci, co, M, accumulate_dw);
    HNO_LAUNCH_CHECK();
  }
  return 0;
}

// torch.optim.Adamax (single-tensor semantics) on a flat vector:
//   g = grad*grad_scale + wd*p;  m = lerp(m, g, 1-b1);  u = max(b2*u, |g| + eps);  p -= lr/(1-b1^t) * m/u
__global__ void __launch_bounds__(256) k_adamax(float* __restrict__ p, const float* __restrict__ grad,
                                                float* __restrict__ m, float* __restrict__ u, long n, float clr,
                                                float b1, float b2, float eps, float wd, float gs) {
  const long i = blockIdx.x * 256L + threadIdx.x;
  if (i >= n) return;
  float g = grad[i] * gs;
  const float pv = p[i];
  if (wd != 0.f) g = fmaf(wd, pv, g);
  const float mv = m[i] + (g - m[i]) * (1.f - b1);
  const float uv = fmaxf(u[i] * b2, fabsf(g) + eps);
  m[i] = mv;
  u[i] = uv;
  p[i] = pv - clr * (mv / uv);
}

int adamax_step(float* param, const float* grad, float* exp_avg, float* exp_inf, long n, float lr, float beta1,
                float beta2, float eps, float weight_decay, int step, float grad_scale, cudaStream_t st) {
  HNO_CHECK(param && grad && exp_avg && exp_inf && n >= 0 && step >= 1, "adamax_step: bad arguments");
  if (n == 0) return 0;
  const double bc = 1.0 - pow((double)beta1, (double)step);
  const float clr = (float)((double)lr / bc);
  k_adamax<<<ceil_div(n, 256), 256, 0, st>>>(param, grad, exp_avg, exp_inf, n, clr, beta1, beta2, eps, weight_decay,
                                              grad_scale);
  HNO_LAUNCH_CHECK();
  return 0;
}

}  // namespace hno
