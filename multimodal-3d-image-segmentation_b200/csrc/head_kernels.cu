// Output head and losses of HNOSeg-XS for sm_100a.
//
// Replaces nets/hnosegxs.py:174-180 (trilinear F.interpolate -> 1x1x1 conv_out -> softmax) and
// nets/custom_losses.py:17-111 (PCCLoss, DiceLoss), plus experiments/utils.py:74-97 (to_categorical)
// in the fused variant.  conv_out has no bias and trilinear weights sum to one, so the conv commutes
// with the interpolation exactly: it is applied at LOW resolution (hno_pwconv_forward, 24 -> classes)
// and only `classes` channels are up-sampled; the 24-channel full-resolution tensor (857 MB/sample at
// 240x240x155) never exists.
//
// Interpolation follows ATen's upsample_trilinear3d with align_corners=False and scales=None
// (ratio = in/out in fp32, src = max(ratio*(dst+0.5)-0.5, 0), i0 = min(floor(src), in-1),
// i1 = i0 + (i0 < in-1), lambda1 = clamp(src - i0, 0, 1)); the per-axis tables are built on the host.
//
// Backward of the interpolation is a deterministic gather in three separable passes (W, then H, then
// D): for every low-resolution index the contiguous range of high-resolution indices that touch it is
// tabulated, so no atomics are needed.
#include "common.cuh"
#include "hno_b200.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>

namespace hno {

constexpr int kInterpMagic = 0x484E4F49;  // 'HNOI'
constexpr int kMaxClasses = 8;

struct InterpHeader {
  int magic;
  int lo[3];      // D, H, W   (low resolution)
  int hi[3];      // Dx, Hx, Wx
  int off_i0[3];  // int [hi]
  int off_i1[3];  // int [hi]
  int off_l1[3];  // float [hi]  weight of i1 (weight of i0 is 1 - l1)
  int off_s[3];   // int [lo]  first hi index touching lo
  int off_e[3];   // int [lo]  one past the last
  int total_words;
  int pad[5];
};

size_t interp_tables_bytes(int D, int H, int W, int Dx, int Hx, int Wx) {
  size_t words = sizeof(InterpHeader) / 4 + 3 * ((size_t)Dx + Hx + Wx) + 2 * ((size_t)D + H + W) + 64;
  return words * 4;
}

int interp_tables_fill(void* buf, size_t bytes, int D, int H, int W, int Dx, int Hx, int Wx) {
  HNO_CHECK(buf, "interp_tables_fill: null buffer");
  HNO_CHECK(bytes >= interp_tables_bytes(D, H, W, Dx, Hx, Wx), "interp_tables_fill: buffer too small");
  HNO_CHECK(D >= 1 && H >= 1 && W >= 1 && Dx >= 1 && Hx >= 1 && Wx >= 1, "interp_tables_fill: bad sizes");
  memset(buf, 0, bytes);
  auto* hdr = reinterpret_cast<InterpHeader*>(buf);
  int* iw = reinterpret_cast<int*>(buf);
  float* fw = reinterpret_cast<float*>(buf);
  hdr->magic = kInterpMagic;
  const int lo[3] = {D, H, W}, hi[3] = {Dx, Hx, Wx};
  size_t cur = sizeof(InterpHeader) / 4;
  for (int a = 0; a < 3; ++a) {
    hdr->lo[a] = lo[a];
    hdr->hi[a] = hi[a];
    hdr->off_i0[a] = (int)cur; cur += hi[a];
    hdr->off_i1[a] = (int)cur; cur += hi[a];
    hdr->off_l1[a] = (int)cur; cur += hi[a];
    hdr->off_s[a] = (int)cur; cur += lo[a];
    hdr->off_e[a] = (int)cur; cur += lo[a];
    int* i0 = iw + hdr->off_i0[a];
    int* i1 = iw + hdr->off_i1[a];
    float* l1 = fw + hdr->off_l1[a];
    int* s = iw + hdr->off_s[a];
    int* e = iw + hdr->off_e[a];
    for (int t = 0; t < lo[a]; ++t) { s[t] = hi[a]; e[t] = 0; }
    const float ratio = (float)lo[a] / (float)hi[a];
    for (int o = 0; o < hi[a]; ++o) {
      if (hi[a] == lo[a]) {
        i0[o] = i1[o] = o;
        l1[o] = 0.f;
      } else {
        // one fused multiply-add: bit-identical to ATen's CPU (AVX2/AVX512 builds) and CUDA kernels,
        // verified against F.interpolate on ramps in tests/test_plan_cpu.py
        float src = fmaf(ratio, (float)o + 0.5f, -0.5f);
        if (src < 0.f) src = 0.f;
        int f = (int)floorf(src);
        if (f > lo[a] - 1) f = lo[a] - 1;
        float lam = src - (float)f;
        lam = lam < 0.f ? 0.f : (lam > 1.f ? 1.f : lam);
        i0[o] = f;
        i1[o] = f + (f < lo[a] - 1 ? 1 : 0);
        l1[o] = lam;
      }
      for (int t : {i0[o], i1[o]}) {
        if (o < s[t]) s[t] = o;
        if (o + 1 > e[t]) e[t] = o + 1;
      }
    }
    for (int t = 0; t < lo[a]; ++t)
      if (e[t] < s[t]) s[t] = e[t] = 0;  // untouched low index (down-sampling case): empty range
  }
  hdr->total_words = (int)cur;
  return 0;
}

struct InterpDev {
  int lo[3], hi[3];
  const int* i0[3];
  const int* i1[3];
  const float* l1[3];
  const int* s[3];
  const int* e[3];
};

static int make_dev(const void* th, const void* td, InterpDev* d) {
  HNO_CHECK(th && td, "head: null interpolation tables");
  const auto* h = reinterpret_cast<const InterpHeader*>(th);
  HNO_CHECK(h->magic == kInterpMagic, "head: bad interpolation table blob");
  const int* iw = reinterpret_cast<const int*>(td);
  const float* fw = reinterpret_cast<const float*>(td);
  for (int a = 0; a < 3; ++a) {
    d->lo[a] = h->lo[a];
    d->hi[a] = h->hi[a];
    d->i0[a] = iw + h->off_i0[a];
    d->i1[a] = iw + h->off_i1[a];
    d->l1[a] = fw + h->off_l1[a];
    d->s[a] = iw + h->off_s[a];
    d->e[a] = iw + h->off_e[a];
  }
  return 0;
}

// logits at one high-resolution voxel, same nesting as ATen: W innermost, then H, then D.
template <int C>
__device__ __forceinline__ void interp_logits(const float* __restrict__ ll, long S, long P, int W, const InterpDev& t,
                                              int zd, int zh, int zw, float (&out)[C]) {
  const int d0 = t.i0[0][zd], d1 = t.i1[0][zd];
  const int h0 = t.i0[1][zh], h1 = t.i1[1][zh];
  const int w0 = t.i0[2][zw], w1 = t.i1[2][zw];
  const float ld1 = t.l1[0][zd], lh1 = t.l1[1][zh], lw1 = t.l1[2][zw];
  const float ld0 = 1.f - ld1, lh0 = 1.f - lh1, lw0 = 1.f - lw1;
  const long o00 = (long)d0 * P + (long)h0 * W, o01 = (long)d0 * P + (long)h1 * W;
  const long o10 = (long)d1 * P + (long)h0 * W, o11 = (long)d1 * P + (long)h1 * W;
#pragma unroll
  for (int c = 0; c < C; ++c) {
    const float* p = ll + (long)c * S;
    const float a00 = lw0 * __ldg(p + o00 + w0) + lw1 * __ldg(p + o00 + w1);
    const float a01 = lw0 * __ldg(p + o01 + w0) + lw1 * __ldg(p + o01 + w1);
    const float a10 = lw0 * __ldg(p + o10 + w0) + lw1 * __ldg(p + o10 + w1);
    const float a11 = lw0 * __ldg(p + o11 + w0) + lw1 * __ldg(p + o11 + w1);
    out[c] = ld0 * (lh0 * a00 + lh1 * a01) + ld1 * (lh0 * a10 + lh1 * a11);
  }
}

template <int C>
__device__ __forceinline__ void softmax_inplace(float (&v)[C]) {
  float m = v[0];
#pragma unroll
  for (int c = 1; c < C; ++c) m = fmaxf(m, v[c]);
  float s = 0.f;
#pragma unroll
  for (int c = 0; c < C; ++c) {
    v[c] = expf(v[c] - m);
    s += v[c];
  }
  const float inv = 1.f / s;
#pragma unroll
  for (int c = 0; c < C; ++c) v[c] *= inv;
}

// Softmax of the fused training kernels: ex2.approx on a log2(e)-scaled argument and rcp.approx (each ~1 ulp; the
// probabilities differ from expf / IEEE division by <= ~1e-6 relative, far inside the parity tolerance, and the row
// kernels drop from ~70 to ~25 instructions per voxel for it).  The drop-in head keeps softmax_inplace.
template <int C>
__device__ __forceinline__ void softmax_fast(float (&v)[C]) {
  float m = v[0];
#pragma unroll
  for (int c = 1; c < C; ++c) m = fmaxf(m, v[c]);
  const float ml = m * kLog2e;
  float s = 0.f;
#pragma unroll
  for (int c = 0; c < C; ++c) {
    v[c] = ex2_approx(fmaf(v[c], kLog2e, -ml));
    s += v[c];
  }
  float inv;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(inv) : "f"(s));
#pragma unroll
  for (int c = 0; c < C; ++c) v[c] *= inv;
}

// ------------------------------------------------------------------------------------------ head forward
template <int C, int ACT>
__global__ void __launch_bounds__(256) k_head_fwd(const float* __restrict__ ll, float* __restrict__ probs,
                                                  InterpDev t, long P) {
  const long Nx = (long)t.hi[0] * t.hi[1] * t.hi[2];
  const long v = blockIdx.x * 256L + threadIdx.x;
  if (v >= Nx) return;
  const int b = blockIdx.y;
  const int zw = (int)(v % t.hi[2]);
  const long r = v / t.hi[2];
  const int zh = (int)(r % t.hi[1]);
  const int zd = (int)(r / t.hi[1]);
  const long S = (long)t.lo[0] * P;
  float lg[C];
  interp_logits<C>(ll + (long)b * C * S, S, P, t.lo[2], t, zd, zh, zw, lg);
  if (ACT == 1) softmax_inplace<C>(lg);
  float* po = probs + (long)b * C * Nx + v;
#pragma unroll
  for (int c = 0; c < C; ++c) po[(long)c * Nx] = lg[c];
}

// Inference head: the label map itself.  argmax over the classes of the interpolated logits (softmax is monotone, so
// this is the label the reference gets from np.argmax(probs, 1) on the host, experiments/train_test.py:402-408; ties
// go to the lowest class index like numpy).  One byte per voxel leaves the device instead of 4 * C.
template <int C>
__global__ void __launch_bounds__(256) k_head_argmax(const float* __restrict__ ll, uint8_t* __restrict__ labels,
                                                     InterpDev t, long P) {
  const long Nx = (long)t.hi[0] * t.hi[1] * t.hi[2];
  const long v = blockIdx.x * 256L + threadIdx.x;
  if (v >= Nx) return;
  const int b = blockIdx.y;
  const int zw = (int)(v % t.hi[2]);
  const long r = v / t.hi[2];
  const int zh = (int)(r % t.hi[1]);
  const int zd = (int)(r / t.hi[1]);
  const long S = (long)t.lo[0] * P;
  float lg[C];
  interp_logits<C>(ll + (long)b * C * S, S, P, t.lo[2], t, zd, zh, zw, lg);
  int best = 0;
  float m = lg[0];
#pragma unroll
  for (int c = 1; c < C; ++c)
    if (lg[c] > m) {
      m = lg[c];
      best = c;
    }
  labels[(long)b * Nx + v] = (uint8_t)best;
}

// ------------------------------------------------------------------------------------------ head backward
// MODE 0: dprobs / probs tensors are given (drop-in autograd path).
// MODE 1: probabilities are recomputed from the low-resolution logits and dL/dprobs comes from the
//         loss coefficients and the integer labels (fused training step).
template <int C, int ACT, int MODE>
__global__ void __launch_bounds__(256) k_head_bwd_w(const float* __restrict__ dprobs, const float* __restrict__ probs,
                                                    const float* __restrict__ ll, const uint8_t* __restrict__ labels,
                                                    const float* __restrict__ coef,
                                                    const float* __restrict__ grad_loss, float* __restrict__ g1,
                                                    InterpDev t, long P) {
  // thread per (zd, zh, w_lo); output g1[b][c][zd][zh][w_lo]
  const int W = t.lo[2];
  const long rows = (long)t.hi[0] * t.hi[1];
  const long idx = blockIdx.x * 256L + threadIdx.x;
  if (idx >= rows * W) return;
  const int b = blockIdx.y;
  const int wl = (int)(idx % W);
  const long row = idx / W;
  const int zh = (int)(row % t.hi[1]);
  const int zd = (int)(row / t.hi[1]);
  const long Nx = rows * t.hi[2];
  const long S = (long)t.lo[0] * P;
  float acc[C];
#pragma unroll
  for (int c = 0; c < C; ++c) acc[c] = 0.f;
  float ca[C], cb[C], cg[C];
  if (MODE == 1) {
    const float gl = grad_loss ? __ldg(grad_loss) : 1.f;
#pragma unroll
    for (int c = 0; c < C; ++c) {
      ca[c] = gl * __ldg(coef + ((long)b * C + c) * 3 + 0);
      cb[c] = gl * __ldg(coef + ((long)b * C + c) * 3 + 1);
      cg[c] = gl * __ldg(coef + ((long)b * C + c) * 3 + 2);
    }
  }
  const int ws = t.s[2][wl], we = t.e[2][wl];
  for (int zw = ws; zw < we; ++zw) {
    const float l1 = t.l1[2][zw];
    const float wgt = (t.i0[2][zw] == wl ? 1.f - l1 : 0.f) + (t.i1[2][zw] == wl ? l1 : 0.f);
    if (wgt == 0.f) continue;
    const long v = row * t.hi[2] + zw;
    float p[C], g[C];
    if (MODE == 0) {
#pragma unroll
      for (int c = 0; c < C; ++c) {
        g[c] = __ldg(dprobs + ((long)b * C + c) * Nx + v);
        if (ACT == 1) p[c] = __ldg(probs + ((long)b * C + c) * Nx + v);
      }
    } else {
      interp_logits<C>(ll + (long)b * C * S, S, P, W, t, zd, zh, zw, p);
      if (ACT == 1) softmax_inplace<C>(p);
      const int lab = labels[(long)b * Nx + v];
#pragma unroll
      for (int c = 0; c < C; ++c) g[c] = ca[c] + (lab == c ? cb[c] : 0.f) + cg[c] * p[c];
    }
    if (ACT == 1) {
      float dot = 0.f;
#pragma unroll
      for (int c = 0; c < C; ++c) dot = fmaf(g[c], p[c], dot);
#pragma unroll
      for (int c = 0; c < C; ++c) acc[c] = fmaf(wgt, p[c] * (g[c] - dot), acc[c]);
    } else {
#pragma unroll
      for (int c = 0; c < C; ++c) acc[c] = fmaf(wgt, g[c], acc[c]);
    }
  }
  const long plane = rows * W;
#pragma unroll
  for (int c = 0; c < C; ++c) g1[((long)b * C + c) * plane + idx] = acc[c];
}

// g1[bc][zd][zh][w] -> g2[bc][zd][h][w]; grid (ceil(H*W / 256), nbc * Dx): 32-bit index arithmetic only
__global__ void __launch_bounds__(256) k_head_bwd_h(const float* __restrict__ g1, float* __restrict__ g2, InterpDev t,
                                                    long nbc) {
  const int W = t.lo[2], H = t.lo[1], Hx = t.hi[1];
  const int hw = blockIdx.x * 256 + threadIdx.x;
  if (hw >= H * W) return;
  const long r = blockIdx.y;  // bc * Dx + zd
  const int h = hw / W, w = hw - h * W;
  const float* src = g1 + r * (long)Hx * W + w;
  float acc = 0.f;
  for (int zh = t.s[1][h]; zh < t.e[1][h]; ++zh) {
    const float l1 = t.l1[1][zh];
    const float wgt = (t.i0[1][zh] == h ? 1.f - l1 : 0.f) + (t.i1[1][zh] == h ? l1 : 0.f);
    acc = fmaf(wgt, __ldg(src + zh * W), acc);
  }
  g2[r * (long)H * W + hw] = acc;
}

// g2[bc][zd][h][w] -> dll[bc][d][P]  (padding columns zeroed); grid (ceil(P / 256), D, nbc)
__global__ void __launch_bounds__(256) k_head_bwd_d(const float* __restrict__ g2, float* __restrict__ dll, InterpDev t,
                                                    long nbc, long P) {
  const int W = t.lo[2], H = t.lo[1], D = t.lo[0], Dx = t.hi[0];
  const int p = blockIdx.x * 256 + threadIdx.x;
  if (p >= P) return;
  const int d = blockIdx.y;
  const long bc = blockIdx.z;
  float* dst = dll + (bc * D + d) * P + p;
  if (p >= H * W) {
    *dst = 0.f;
    return;
  }
  const int HW = H * W;
  const float* src = g2 + bc * (long)Dx * HW + p;
  float acc = 0.f;
  for (int zd = t.s[0][d]; zd < t.e[0][d]; ++zd) {
    const float l1 = t.l1[0][zd];
    const float wgt = (t.i0[0][zd] == d ? 1.f - l1 : 0.f) + (t.i1[0][zd] == d ? l1 : 0.f);
    acc = fmaf(wgt, __ldg(src + (long)zd * HW), acc);
  }
  *dst = acc;
}

// Gather-table variants of the two passes above (round 2).  The per-thread loops over [s, e) with three table loads and two
// compares per tap made both passes instruction bound (190 us for 324 MB).  Here the (at most kGatherK) taps of every
// low-resolution index are turned into a fixed-length weight row ONCE per CTA (shared memory: first tap + kGatherK weights,
// zero beyond the range), a thread reuses its weights for several slices, and the inner loop is K loads + K FMAs with no
// branches.  Same taps, same order, same products as the loops above (zero-weight terms add exactly 0).
constexpr int kGatherMax = 6;  // instantiated for 4 taps (2x up-sampling) and 6

template <int AX, int kGatherK>
__device__ __forceinline__ void head_build_gather(const InterpDev& t, int* ss, float* sw) {
  for (int i = threadIdx.x; i < t.lo[AX]; i += blockDim.x) {
    const int s = t.s[AX][i], e = t.e[AX][i];
    ss[i] = s;
#pragma unroll
    for (int k = 0; k < kGatherK; ++k) {
      const int z = s + k;
      float w = 0.f;
      if (z < e) {
        const float l1 = t.l1[AX][z];
        w = (t.i0[AX][z] == i ? 1.f - l1 : 0.f) + (t.i1[AX][z] == i ? l1 : 0.f);
      }
      sw[i * kGatherK + k] = w;
    }
  }
}

// g1[r][zh][w] -> g2[r][h][w], r = bc * Dx + zd; grid (ceil(H*W / 256), ceil(R / 4))
template <int kGatherK>
__global__ void __launch_bounds__(256) k_head_bwd_h_gather(const float* __restrict__ g1, float* __restrict__ g2, InterpDev t,
                                                           long R) {
  extern __shared__ float sgw[];
  const int W = t.lo[2], H = t.lo[1], Hx = t.hi[1];
  int* ss = reinterpret_cast<int*>(sgw);
  float* sw = sgw + H;
  head_build_gather<1, kGatherK>(t, ss, sw);
  __syncthreads();
  const int hw = blockIdx.x * 256 + threadIdx.x;
  if (hw >= H * W) return;
  const int h = hw / W, w = hw - h * W;
  float wk[kGatherK];
  int off[kGatherK];
  const int s0 = ss[h];
#pragma unroll
  for (int k = 0; k < kGatherK; ++k) {
    wk[k] = sw[h * kGatherK + k];
    off[k] = min(s0 + k, Hx - 1) * W + w;  // taps beyond the range have weight 0: any valid address will do
  }
  const long r0 = (long)blockIdx.y * 4;
  float v[4][kGatherK];  // all loads of the four slices first (one slice at a time left the kernel latency bound: ncu
                         // long_scoreboard 12 stalls per issue, 2.5 TB/s)
#pragma unroll
  for (int rb = 0; rb < 4; ++rb) {
    const float* src = g1 + min(r0 + rb, R - 1) * (long)Hx * W;
#pragma unroll
    for (int k = 0; k < kGatherK; ++k) v[rb][k] = __ldg(src + off[k]);
  }
#pragma unroll
  for (int rb = 0; rb < 4; ++rb) {
    const long r = r0 + rb;
    float acc = 0.f;
#pragma unroll
    for (int k = 0; k < kGatherK; ++k) acc = fmaf(wk[k], v[rb][k], acc);
    if (r < R) g2[r * (long)H * W + hw] = acc;
  }
}

// g2[bc][zd][h*w] -> dll[bc][d][P] (padding columns zeroed); grid (ceil(P / 256), D, nbc)
template <int kGatherK>
__global__ void __launch_bounds__(256) k_head_bwd_d_gather(const float* __restrict__ g2, float* __restrict__ dll, InterpDev t,
                                                           long P, long nbc) {
  __shared__ float swk[kGatherK];
  __shared__ int ss0;
  const int W = t.lo[2], H = t.lo[1], D = t.lo[0], Dx = t.hi[0];
  const int d = blockIdx.y;
  if (threadIdx.x < kGatherK) {
    const int k = threadIdx.x, s = t.s[0][d], e = t.e[0][d], z = s + k;
    float w = 0.f;
    if (z < e) {
      const float l1 = t.l1[0][z];
      w = (t.i0[0][z] == d ? 1.f - l1 : 0.f) + (t.i1[0][z] == d ? l1 : 0.f);
    }
    swk[k] = w;
    if (k == 0) ss0 = s;
  }
  __syncthreads();
  const int p = blockIdx.x * 256 + threadIdx.x;
  if (p >= P) return;
  const long bc0 = (long)blockIdx.z * 4;  // four (b, c) slices per thread: 4 x K loads in flight, one weight prologue
  const int HW = H * W;
  if (p >= HW) {
#pragma unroll
    for (int j = 0; j < 4; ++j)
      if (bc0 + j < nbc) dll[((bc0 + j) * D + d) * P + p] = 0.f;
    return;
  }
  float v[4][kGatherK];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const float* src = g2 + min(bc0 + j, nbc - 1) * (long)Dx * HW + p;
#pragma unroll
    for (int k = 0; k < kGatherK; ++k) v[j][k] = __ldg(src + (long)min(ss0 + k, Dx - 1) * HW);
  }
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    float acc = 0.f;
#pragma unroll
    for (int k = 0; k < kGatherK; ++k) acc = fmaf(swk[k], v[j][k], acc);
    if (bc0 + j < nbc) dll[((bc0 + j) * D + d) * P + p] = acc;
  }
}

// ------------------------------------------------------------------------------------------ losses
// Five moments per (b, c): sum p, sum t, sum p*t, sum p*p, sum t*t.  Block partials in fp64.
constexpr int kMoments = 5;
constexpr int kLossChunks = 296;  // partial rows per (b, c) / per b

__device__ __forceinline__ void block_reduce_store(double (&m)[kMoments], double* dst) {
  __shared__ double sred[8][kMoments];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int k = 0; k < kMoments; ++k) m[k] = warp_sum_d(m[k]);
  __syncthreads();
  if (lane == 0) {
#pragma unroll
    for (int k = 0; k < kMoments; ++k) sred[warp][k] = m[k];
  }
  __syncthreads();
  if (threadIdx.x < kMoments) {
    double s = 0.0;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) s += sred[w][threadIdx.x];
    dst[threadIdx.x] = s;
  }
}

// grid (kLossChunks, B*C); y_pred / y_true [B*C][N].  V = 4: 16-byte loads, two pairs in flight per thread (N % 4 == 0 and
// 16-byte aligned bases; the launcher checks), V = 1: any N.
template <int V>
__global__ void __launch_bounds__(256) k_loss_moments(const float* __restrict__ yp, const float* __restrict__ yt,
                                                      double* __restrict__ partials, long N) {
  const long bc = blockIdx.y;
  const float* p = yp + bc * N;
  const float* t = yt + bc * N;
  float m[kMoments] = {0.f, 0.f, 0.f, 0.f, 0.f};
  double md[kMoments] = {0.0, 0.0, 0.0, 0.0, 0.0};
  int cnt = 0;
  auto add = [&](float a, float b) {
    m[0] += a;
    m[1] += b;
    m[2] = fmaf(a, b, m[2]);
    m[3] = fmaf(a, a, m[3]);
    m[4] = fmaf(b, b, m[4]);
  };
  auto spill = [&]() {  // bounded fp32 run length (64 elements), then into fp64
#pragma unroll
    for (int k = 0; k < kMoments; ++k) {
      md[k] += (double)m[k];
      m[k] = 0.f;
    }
    cnt = 0;
  };
  if (V == 4) {
    const float4* p4 = reinterpret_cast<const float4*>(p);
    const float4* t4 = reinterpret_cast<const float4*>(t);
    const long n4 = N >> 2, step = (long)gridDim.x * 256L;
    long i = blockIdx.x * 256L + threadIdx.x;
    for (; i + step < n4; i += 2 * step) {
      const float4 a0 = __ldg(p4 + i), b0 = __ldg(t4 + i), a1 = __ldg(p4 + i + step), b1 = __ldg(t4 + i + step);
      add(a0.x, b0.x); add(a0.y, b0.y); add(a0.z, b0.z); add(a0.w, b0.w);
      add(a1.x, b1.x); add(a1.y, b1.y); add(a1.z, b1.z); add(a1.w, b1.w);
      if (++cnt == 8) spill();
    }
    if (i < n4) {
      const float4 a0 = __ldg(p4 + i), b0 = __ldg(t4 + i);
      add(a0.x, b0.x); add(a0.y, b0.y); add(a0.z, b0.z); add(a0.w, b0.w);
    }
  } else {
    for (long i = blockIdx.x * 256L + threadIdx.x; i < N; i += (long)gridDim.x * 256L) {
      add(__ldg(p + i), __ldg(t + i));
      if (++cnt == 64) spill();
    }
  }
#pragma unroll
  for (int k = 0; k < kMoments; ++k) md[k] += (double)m[k];
  block_reduce_store(md, partials + (bc * gridDim.x + blockIdx.x) * kMoments);
}

// fused head + moments on integer labels: grid (kLossChunks, B); partials [B][C][chunks][5]
template <int C, int ACT>
__global__ void __launch_bounds__(256) k_head_loss_moments(const float* __restrict__ ll,
                                                           const uint8_t* __restrict__ labels,
                                                           double* __restrict__ partials, InterpDev t, long P) {
  const long Nx = (long)t.hi[0] * t.hi[1] * t.hi[2];
  const int b = blockIdx.y;
  const long S = (long)t.lo[0] * P;
  float m[C][4];
  double md[C][4];
#pragma unroll
  for (int c = 0; c < C; ++c)
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      m[c][k] = 0.f;
      md[c][k] = 0.0;
    }
  int cnt = 0;
  for (long v = blockIdx.x * 256L + threadIdx.x; v < Nx; v += (long)gridDim.x * 256L) {
    const int zw = (int)(v % t.hi[2]);
    const long r = v / t.hi[2];
    const int zh = (int)(r % t.hi[1]);
    const int zd = (int)(r / t.hi[1]);
    float p[C];
    interp_logits<C>(ll + (long)b * C * S, S, P, t.lo[2], t, zd, zh, zw, p);
    if (ACT == 1) softmax_inplace<C>(p);
    const int lab = labels[(long)b * Nx + v];
#pragma unroll
    for (int c = 0; c < C; ++c) {
      const float tt = lab == c ? 1.f : 0.f;
      m[c][0] += p[c];
      m[c][1] += tt;
      m[c][2] = fmaf(p[c], tt, m[c][2]);
      m[c][3] = fmaf(p[c], p[c], m[c][3]);
    }
    if (++cnt == 64) {
#pragma unroll
      for (int c = 0; c < C; ++c)
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          md[c][k] += (double)m[c][k];
          m[c][k] = 0.f;
        }
      cnt = 0;
    }
  }
#pragma unroll
  for (int c = 0; c < C; ++c) {
    double mm[kMoments];
#pragma unroll
    for (int k = 0; k < 4; ++k) mm[k] = md[c][k] + (double)m[c][k];
    mm[4] = mm[1];  // t*t == t for one-hot labels
    block_reduce_store(mm, partials + (((long)b * C + c) * gridDim.x + blockIdx.x) * kMoments);
  }
}

// ---- row kernels of the fused head + loss (the training hot path) --------------------------------------------------
// One CTA walks whole high-resolution rows (b, zd, zh); a warp owns 32 consecutive voxels of the row (forward) or a
// group of low-resolution taps and every voxel that touches them (backward).  The warp first blends the four
// low-resolution rows (d0|d1 x h0|h1) of every class into ONE row segment in shared memory with coalesced loads, then
// each lane interpolates along W from that segment: 2 shared loads per class and voxel instead of 8 scattered global
// ones, and every voxel's softmax is evaluated exactly once (the gather of the W pass goes through shared memory).
// (Trilinear interpolation is linear, so blending D and H before W is the same map as ATen's W-first nesting up to
// fp32 rounding of the intermediate sums, ~1e-7 relative.)
constexpr int kRowWarps = 6;

// per-row constants of the D / H blend (identical for every lane: computed once per row)
struct RowBlend {
  int o00, o01, o10, o11;  // offsets of the four low-resolution rows inside one class volume (fit 32 bits)
  float c00, c01, c10, c11;
};
__device__ __forceinline__ RowBlend make_row_blend(const InterpDev& t, long P, int W, int zd, int zh) {
  const int d0 = t.i0[0][zd], d1 = t.i1[0][zd];
  const int h0 = t.i0[1][zh], h1 = t.i1[1][zh];
  const float ld1 = t.l1[0][zd], lh1 = t.l1[1][zh];
  const float ld0 = 1.f - ld1, lh0 = 1.f - lh1;
  RowBlend r;
  r.o00 = d0 * (int)P + h0 * W;
  r.o01 = d0 * (int)P + h1 * W;
  r.o10 = d1 * (int)P + h0 * W;
  r.o11 = d1 * (int)P + h1 * W;
  r.c00 = lh0;
  r.c01 = lh1;
  r.c10 = ld0;
  r.c11 = ld1;
  return r;
}
template <int C>
__device__ __forceinline__ void blend_rows(const float* __restrict__ llb, int S, const RowBlend& rb, int w,
                                           float (&r)[C]) {
  float a[C][4];
#pragma unroll
  for (int c = 0; c < C; ++c) {  // all loads first: 4 C independent requests in flight per lane
    const float* p = llb + c * S + w;
    a[c][0] = __ldg(p + rb.o00);
    a[c][1] = __ldg(p + rb.o01);
    a[c][2] = __ldg(p + rb.o10);
    a[c][3] = __ldg(p + rb.o11);
  }
#pragma unroll
  for (int c = 0; c < C; ++c)
    r[c] = rb.c10 * (rb.c00 * a[c][0] + rb.c01 * a[c][1]) + rb.c11 * (rb.c00 * a[c][2] + rb.c01 * a[c][3]);
}

// forward: moments per (b, c); grid (chunks, B); partials [B][C][chunks][5]
template <int C, int ACT>
__global__ void __launch_bounds__(32 * kRowWarps, 4) k_head_loss_rows(const float* __restrict__ ll,
                                                                      const uint8_t* __restrict__ labels,
                                                                      double* __restrict__ partials, InterpDev t,
                                                                      long P) {
  __shared__ float sR[kRowWarps][C][32];
  __shared__ double sacc[kRowWarps][C][4];  // per-warp fp64 running sums (the lanes' fp32 sums are folded in every 64 rows)
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int b = blockIdx.y;
  const int Wx = t.hi[2], Hx = t.hi[1], W = t.lo[2];
  const int rows = t.hi[0] * Hx;
  const int S = t.lo[0] * (int)P;
  const float* llb = ll + (long)b * C * S;
  const uint8_t* lb = labels + (long)b * rows * Wx;
  float m[C][4];
#pragma unroll
  for (int c = 0; c < C; ++c)
#pragma unroll
    for (int k = 0; k < 4; ++k) m[c][k] = 0.f;
  if (lane < C * 4) sacc[warp][lane >> 2][lane & 3] = 0.0;
  __syncwarp();
  auto fold = [&]() {
#pragma unroll
    for (int c = 0; c < C; ++c)
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const float v = warp_sum(m[c][k]);
        if (lane == 0) sacc[warp][c][k] += (double)v;
        m[c][k] = 0.f;
      }
  };
  int cnt = 0;
  const int nseg = (Wx + 31) >> 5;
  for (int row = blockIdx.x; row < rows; row += gridDim.x) {
    const int zd = row / Hx, zh = row - zd * Hx;
    const RowBlend rb = make_row_blend(t, P, W, zd, zh);
    const uint8_t* lrow = lb + (long)row * Wx;
    for (int seg = warp; seg < nseg; seg += kRowWarps) {
      const int zw0 = seg << 5;
      const int zw = zw0 + lane;
      const int lab = zw < Wx ? (int)lrow[zw] : 0;  // requested before the blend: its latency hides behind it
      const int wb = t.i0[2][zw0];
      const int ntap = t.i1[2][min(zw0 + 31, Wx - 1)] - wb + 1;
      if (lane < ntap) {
        float r[C];
        blend_rows<C>(llb, S, rb, wb + lane, r);
#pragma unroll
        for (int c = 0; c < C; ++c) sR[warp][c][lane] = r[c];
      }
      __syncwarp();
      if (zw < Wx) {
        const int w0 = t.i0[2][zw] - wb, w1 = t.i1[2][zw] - wb;
        const float lw1 = t.l1[2][zw], lw0 = 1.f - lw1;
        float p[C];
#pragma unroll
        for (int c = 0; c < C; ++c) p[c] = lw0 * sR[warp][c][w0] + lw1 * sR[warp][c][w1];
        if (ACT == 1) softmax_fast<C>(p);
#pragma unroll
        for (int c = 0; c < C; ++c) {
          const float tt = lab == c ? 1.f : 0.f;
          m[c][0] += p[c];
          m[c][1] += tt;
          m[c][2] = fmaf(p[c], tt, m[c][2]);
          m[c][3] = fmaf(p[c], p[c], m[c][3]);
        }
      }
      __syncwarp();
    }
    if (++cnt == 64) {  // bounded fp32 run length, then into fp64
      fold();
      cnt = 0;
    }
  }
  fold();
  __syncthreads();
  for (int i = threadIdx.x; i < C * kMoments; i += blockDim.x) {
    const int c = i / kMoments, k = i - c * kMoments;
    const int kk = k == 4 ? 1 : k;  // t*t == t for one-hot labels
    double sum = 0.0;
    for (int w = 0; w < kRowWarps; ++w) sum += sacc[w][c][kk];
    partials[(((long)b * C + c) * gridDim.x + blockIdx.x) * kMoments + k] = sum;
  }
}

// backward, W pass: g1[b][c][zd][zh][w_lo] = sum over the voxels of the row that touch tap w_lo of weight * dlogit.
// A warp owns TG consecutive taps and the <= 32 voxels [s(first tap), e(last tap)) that touch them.
template <int C, int ACT>
__global__ void __launch_bounds__(32 * kRowWarps, 4) k_head_bwd_w_rows(const float* __restrict__ ll,
                                                                    const uint8_t* __restrict__ labels,
                                                                    const float* __restrict__ coef,
                                                                    const float* __restrict__ grad_loss,
                                                                    float* __restrict__ g1, InterpDev t, long P,
                                                                    int TG) {
  __shared__ float sR[kRowWarps][C][32];
  __shared__ float sD[kRowWarps][C][32];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int b = blockIdx.y;
  const int Wx = t.hi[2], Hx = t.hi[1], W = t.lo[2];
  const int rows = t.hi[0] * Hx;
  const int S = t.lo[0] * (int)P;
  const float* llb = ll + (long)b * C * S;
  const uint8_t* lb = labels + (long)b * rows * Wx;
  float ca[C], cb[C], cg[C];
  {
    const float gl = grad_loss ? __ldg(grad_loss) : 1.f;
#pragma unroll
    for (int c = 0; c < C; ++c) {
      ca[c] = gl * __ldg(coef + ((long)b * C + c) * 3 + 0);
      cb[c] = gl * __ldg(coef + ((long)b * C + c) * 3 + 1);
      cg[c] = gl * __ldg(coef + ((long)b * C + c) * 3 + 2);
    }
  }
  const int ngroup = (W + TG - 1) / TG;
  for (int row = blockIdx.x; row < rows; row += gridDim.x) {
    const int zd = row / Hx, zh = row - zd * Hx;
    const RowBlend rb = make_row_blend(t, P, W, zd, zh);
    const uint8_t* lrow = lb + (long)row * Wx;
    for (int q = warp; q < ngroup; q += kRowWarps) {
      const int t0 = q * TG, t1 = min(t0 + TG, W);  // taps [t0, t1)
      const int zs = t.s[2][t0], ze = t.e[2][t1 - 1];
      const int zw = zs + lane;
      const int lab = zw < ze ? (int)lrow[zw] : 0;  // requested before the blend: its latency hides behind it
      const int wb = t.i0[2][zs];
      const int ntap = t.i1[2][ze - 1] - wb + 1;
      if (lane < ntap) {
        float r[C];
        blend_rows<C>(llb, S, rb, wb + lane, r);
#pragma unroll
        for (int c = 0; c < C; ++c) sR[warp][c][lane] = r[c];
      }
      __syncwarp();
      if (zw < ze) {
        const int w0 = t.i0[2][zw] - wb, w1 = t.i1[2][zw] - wb;
        const float lw1 = t.l1[2][zw], lw0 = 1.f - lw1;
        float p[C], g[C];
#pragma unroll
        for (int c = 0; c < C; ++c) p[c] = lw0 * sR[warp][c][w0] + lw1 * sR[warp][c][w1];
        if (ACT == 1) softmax_fast<C>(p);
#pragma unroll
        for (int c = 0; c < C; ++c) g[c] = ca[c] + (lab == c ? cb[c] : 0.f) + cg[c] * p[c];
        if (ACT == 1) {
          float dot = 0.f;
#pragma unroll
          for (int c = 0; c < C; ++c) dot = fmaf(g[c], p[c], dot);
#pragma unroll
          for (int c = 0; c < C; ++c) sD[warp][c][lane] = p[c] * (g[c] - dot);
        } else {
#pragma unroll
          for (int c = 0; c < C; ++c) sD[warp][c][lane] = g[c];
        }
      }
      __syncwarp();
      const int wl = t0 + lane;
      if (wl < t1) {
        float acc[C];
#pragma unroll
        for (int c = 0; c < C; ++c) acc[c] = 0.f;
        for (int z2 = t.s[2][wl]; z2 < t.e[2][wl]; ++z2) {
          const float l1 = t.l1[2][z2];
          const float wgt = (t.i0[2][z2] == wl ? 1.f - l1 : 0.f) + (t.i1[2][z2] == wl ? l1 : 0.f);
#pragma unroll
          for (int c = 0; c < C; ++c) acc[c] = fmaf(wgt, sD[warp][c][z2 - zs], acc[c]);
        }
#pragma unroll
        for (int c = 0; c < C; ++c) g1[(((long)b * C + c) * (long)rows + row) * W + wl] = acc[c];
      }
      __syncwarp();
    }
  }
}

// ---- whole-row kernels (round 2) ---------------------------------------------------------------------------------
// The segment kernels above spend most of their issue slots outside the arithmetic: a warp blends ~17 taps on 17 of its
// 32 lanes for every 32-voxel segment (segments overlap by a tap), reloads the W-axis tables for every voxel, and with
// five segments per 155-voxel row the sixth warp of the CTA idles (ncu, round 1: 196 M warp instructions for 17.9 M
// voxels, issue-bound at 2 % of HBM).  Here ONE WARP OWNS A WHOLE ROW: it blends the four low-resolution rows of every
// class over the full low-resolution width into its own shared-memory row (coalesced, all lanes busy, done once per row
// instead of once per segment), and every lane then walks its voxels lane, lane + 32, ... with the W-axis taps and weights
// of those voxels held in registers for the whole kernel (they are the same for every row).
constexpr int kWRowWarps = 8;

template <int C, int ACT, int KMAX>
__global__ void __launch_bounds__(32 * kWRowWarps, 2) k_head_loss_wrow(const float* __restrict__ ll,
                                                                       const uint8_t* __restrict__ labels,
                                                                       double* __restrict__ partials, InterpDev t, long P,
                                                                       int WP) {
  extern __shared__ float sRow[];                      // [kWRowWarps][C][WP]
  __shared__ double sacc[kWRowWarps][C][4];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int b = blockIdx.y;
  const int Wx = t.hi[2], Hx = t.hi[1], W = t.lo[2];
  const int rows = t.hi[0] * Hx;
  const int S = t.lo[0] * (int)P;
  const float* llb = ll + (long)b * C * S;
  const uint8_t* lb = labels + (long)b * rows * Wx;
  float* sR = sRow + warp * C * WP;
  int w0[KMAX], w1[KMAX];
  float lw1[KMAX];
#pragma unroll
  for (int k = 0; k < KMAX; ++k) {
    const int zw = lane + 32 * k;
    const bool ok = zw < Wx;
    w0[k] = ok ? t.i0[2][zw] : 0;
    w1[k] = ok ? t.i1[2][zw] : 0;
    lw1[k] = ok ? t.l1[2][zw] : 0.f;
  }
  float m[C][4];
#pragma unroll
  for (int c = 0; c < C; ++c)
#pragma unroll
    for (int k = 0; k < 4; ++k) m[c][k] = 0.f;
  if (lane < C * 4) sacc[warp][lane >> 2][lane & 3] = 0.0;
  __syncwarp();
  auto fold = [&]() {
#pragma unroll
    for (int c = 0; c < C; ++c)
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const float v = warp_sum(m[c][k]);
        if (lane == 0) sacc[warp][c][k] += (double)v;
        m[c][k] = 0.f;
      }
  };
  int cnt = 0;
  for (int row = blockIdx.x * kWRowWarps + warp; row < rows; row += gridDim.x * kWRowWarps) {
    const int zd = row / Hx, zh = row - zd * Hx;
    const RowBlend rb = make_row_blend(t, P, W, zd, zh);
    const uint8_t* lrow = lb + (long)row * Wx;
    int lab[KMAX];
#pragma unroll
    for (int k = 0; k < KMAX; ++k) lab[k] = lane + 32 * k < Wx ? (int)__ldg(lrow + lane + 32 * k) : -1;
    for (int w = lane; w < W; w += 32) {
      float r[C];
      blend_rows<C>(llb, S, rb, w, r);
#pragma unroll
      for (int c = 0; c < C; ++c) sR[c * WP + w] = r[c];
    }
    __syncwarp();
#pragma unroll
    for (int k = 0; k < KMAX; ++k) {
      if (lane + 32 * k < Wx) {
        const float l1 = lw1[k], l0 = 1.f - l1;
        float p[C];
#pragma unroll
        for (int c = 0; c < C; ++c) p[c] = l0 * sR[c * WP + w0[k]] + l1 * sR[c * WP + w1[k]];
        if (ACT == 1) softmax_fast<C>(p);
#pragma unroll
        for (int c = 0; c < C; ++c) {
          const float tt = lab[k] == c ? 1.f : 0.f;
          m[c][0] += p[c];
          m[c][1] += tt;
          m[c][2] = fmaf(p[c], tt, m[c][2]);
          m[c][3] = fmaf(p[c], p[c], m[c][3]);
        }
      }
    }
    __syncwarp();
    if (++cnt == 16) {  // bounded fp32 run length (16 rows x KMAX voxels per lane), then into fp64
      fold();
      cnt = 0;
    }
  }
  fold();
  __syncthreads();
  for (int i = threadIdx.x; i < C * kMoments; i += blockDim.x) {
    const int c = i / kMoments, k = i - c * kMoments;
    const int kk = k == 4 ? 1 : k;  // t*t == t for one-hot labels
    double sum = 0.0;
    for (int w = 0; w < kWRowWarps; ++w) sum += sacc[w][c][kk];
    partials[(((long)b * C + c) * gridDim.x + blockIdx.x) * kMoments + k] = sum;
  }
}

// backward, W pass with one warp per row: the d(logit) of the whole row goes to the warp's shared-memory row, then lane
// w_lo gathers the (two to five) voxels that touch low-resolution tap w_lo.  g1 as k_head_bwd_w_rows.
// KG > 0: the gather runs on per-CTA weight rows (first voxel + KG weights per low-resolution tap, built once in shared memory)
// instead of walking [s, e) with three table loads and two compares per voxel -- that loop was 43 % of the kernel's
// instructions (ncu source view, 182 M warp instructions for 17.9 M voxels).
template <int C, int ACT, int KMAX, int KG>
__global__ void __launch_bounds__(32 * kWRowWarps, 2) k_head_bwd_wrow(const float* __restrict__ ll,
                                                                      const uint8_t* __restrict__ labels,
                                                                      const float* __restrict__ coef,
                                                                      const float* __restrict__ grad_loss,
                                                                      float* __restrict__ g1, InterpDev t, long P, int WP,
                                                                      int WXP) {
  extern __shared__ float sRow[];                      // [kWRowWarps][C][WP] blended rows, then [kWRowWarps][C][WXP] d(logit)
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int b = blockIdx.y;
  const int Wx = t.hi[2], Hx = t.hi[1], W = t.lo[2];
  const int rows = t.hi[0] * Hx;
  const int S = t.lo[0] * (int)P;
  const float* llb = ll + (long)b * C * S;
  const uint8_t* lb = labels + (long)b * rows * Wx;
  float* sR = sRow + warp * C * WP;
  float* sD = sRow + kWRowWarps * C * WP + warp * C * WXP;
  int* sGs = reinterpret_cast<int*>(sRow + kWRowWarps * C * (WP + WXP));  // [W] first voxel of a tap
  float* sGw = reinterpret_cast<float*>(sGs + W);                          // [W][KG]
  if (KG > 0) {
    head_build_gather<2, (KG > 0 ? KG : 1)>(t, sGs, sGw);
    __syncthreads();
  }
  int w0[KMAX], w1[KMAX];
  float lw1[KMAX];
#pragma unroll
  for (int k = 0; k < KMAX; ++k) {
    const int zw = lane + 32 * k;
    const bool ok = zw < Wx;
    w0[k] = ok ? t.i0[2][zw] : 0;
    w1[k] = ok ? t.i1[2][zw] : 0;
    lw1[k] = ok ? t.l1[2][zw] : 0.f;
  }
  float ca[C], cb[C], cg[C];
  {
    const float gl = grad_loss ? __ldg(grad_loss) : 1.f;
#pragma unroll
    for (int c = 0; c < C; ++c) {
      ca[c] = gl * __ldg(coef + ((long)b * C + c) * 3 + 0);
      cb[c] = gl * __ldg(coef + ((long)b * C + c) * 3 + 1);
      cg[c] = gl * __ldg(coef + ((long)b * C + c) * 3 + 2);
    }
  }
  for (int row = blockIdx.x * kWRowWarps + warp; row < rows; row += gridDim.x * kWRowWarps) {
    const int zd = row / Hx, zh = row - zd * Hx;
    const RowBlend rb = make_row_blend(t, P, W, zd, zh);
    const uint8_t* lrow = lb + (long)row * Wx;
    int lab[KMAX];
#pragma unroll
    for (int k = 0; k < KMAX; ++k) lab[k] = lane + 32 * k < Wx ? (int)__ldg(lrow + lane + 32 * k) : -1;
    for (int w = lane; w < W; w += 32) {
      float r[C];
      blend_rows<C>(llb, S, rb, w, r);
#pragma unroll
      for (int c = 0; c < C; ++c) sR[c * WP + w] = r[c];
    }
    __syncwarp();
#pragma unroll
    for (int k = 0; k < KMAX; ++k) {
      const int zw = lane + 32 * k;
      if (zw < Wx) {
        const float l1 = lw1[k], l0 = 1.f - l1;
        float p[C], g[C];
#pragma unroll
        for (int c = 0; c < C; ++c) p[c] = l0 * sR[c * WP + w0[k]] + l1 * sR[c * WP + w1[k]];
        if (ACT == 1) softmax_fast<C>(p);
#pragma unroll
        for (int c = 0; c < C; ++c) g[c] = ca[c] + (lab[k] == c ? cb[c] : 0.f) + cg[c] * p[c];
        if (ACT == 1) {
          float dot = 0.f;
#pragma unroll
          for (int c = 0; c < C; ++c) dot = fmaf(g[c], p[c], dot);
#pragma unroll
          for (int c = 0; c < C; ++c) sD[c * WXP + zw] = p[c] * (g[c] - dot);
        } else {
#pragma unroll
          for (int c = 0; c < C; ++c) sD[c * WXP + zw] = g[c];
        }
      }
    }
    __syncwarp();
    for (int wl = lane; wl < W; wl += 32) {
      float acc[C];
#pragma unroll
      for (int c = 0; c < C; ++c) acc[c] = 0.f;
      if (KG > 0) {
        const int z0 = sGs[wl];
#pragma unroll
        for (int k = 0; k < KG; ++k) {
          const float wgt = sGw[wl * KG + k];
          const int z2 = min(z0 + k, Wx - 1);  // beyond the range the weight is 0
#pragma unroll
          for (int c = 0; c < C; ++c) acc[c] = fmaf(wgt, sD[c * WXP + z2], acc[c]);
        }
      } else {
        const int z1 = t.e[2][wl];
        for (int z2 = t.s[2][wl]; z2 < z1; ++z2) {
          const float l1 = t.l1[2][z2];
          const float wgt = (t.i0[2][z2] == wl ? 1.f - l1 : 0.f) + (t.i1[2][z2] == wl ? l1 : 0.f);
#pragma unroll
          for (int c = 0; c < C; ++c) acc[c] = fmaf(wgt, sD[c * WXP + z2], acc[c]);
        }
      }
#pragma unroll
      for (int c = 0; c < C; ++c) g1[(((long)b * C + c) * (long)rows + row) * W + wl] = acc[c];
    }
    __syncwarp();
  }
}

// 0: the whole-row kernels do not apply (too wide a row for the register-resident tables / shared-memory rows)
static int wrow_kmax(const InterpDev& t) {
  static const bool on = !(getenv("HNO_HEAD_WROW") && atoi(getenv("HNO_HEAD_WROW")) == 0);
  if (!on || t.hi[2] < t.lo[2] || t.lo[2] > 256) return 0;
  if (t.hi[2] <= 32 * 5) return 5;
  if (t.hi[2] <= 32 * 10) return 10;
  return 0;
}

// Largest tap-group size (<= 16) for which every group's voxels and taps fit one warp; 0 = use the per-voxel kernels.
static int row_kernels_tap_group(const void* th) {
  const auto* h = reinterpret_cast<const InterpHeader*>(th);
  const int* iw = reinterpret_cast<const int*>(th);
  const int W = h->lo[2], Wx = h->hi[2];
  if (Wx < W || W < 1) return 0;
  const int* i0 = iw + h->off_i0[2];
  const int* i1 = iw + h->off_i1[2];
  const int* s = iw + h->off_s[2];
  const int* e = iw + h->off_e[2];
  for (int w = 0; w < W; ++w)
    if (e[w] <= s[w]) return 0;  // untouched tap
  for (int seg = 0; seg * 32 < Wx; ++seg) {  // forward segments
    const int zl = seg * 32 + 31 < Wx - 1 ? seg * 32 + 31 : Wx - 1;
    if (i1[zl] - i0[seg * 32] + 1 > 32) return 0;
  }
  for (int tg = 16; tg >= 1; --tg) {
    bool ok = true;
    for (int t0 = 0; t0 < W && ok; t0 += tg) {
      const int t1 = t0 + tg < W ? t0 + tg : W;
      const int zs = s[t0], ze = e[t1 - 1];
      if (ze - zs > 32 || i1[ze - 1] - i0[zs] + 1 > 32) ok = false;
    }
    if (ok) return tg;
  }
  return 0;
}
static bool row_kernels_enabled() {
  static const bool on = !(getenv("HNO_HEAD_ROWS") && atoi(getenv("HNO_HEAD_ROWS")) == 0);
  return on;
}

// one block: reduce chunk partials, evaluate the loss and the backward coefficients
__global__ void __launch_bounds__(256) k_loss_finalize(const double* __restrict__ partials, int nchunks, int BC,
                                                       double N, int kind, double param,
                                                       float* __restrict__ loss, float* __restrict__ coef) {
  __shared__ double sterm[256];
  double term = 0.0;
  const int fwarp = threadIdx.x >> 5, flane = threadIdx.x & 31, fnw = blockDim.x >> 5;
  for (int bc = fwarp; bc < BC; bc += fnw) {  // one warp per (b, c): lanes stride over the chunk partials
    double m[kMoments] = {0, 0, 0, 0, 0};
    for (int ch = flane; ch < nchunks; ch += 32)
#pragma unroll
      for (int k = 0; k < kMoments; ++k) m[k] += partials[((long)bc * nchunks + ch) * kMoments + k];
#pragma unroll
    for (int k = 0; k < kMoments; ++k) m[k] = warp_sum_d(m[k]);
    if (flane != 0) continue;
    const double sp = m[0], st = m[1], spt = m[2], spp = m[3], stt = m[4];
    double a, b, g;
    if (kind == 0) {  // nets/custom_losses.py:73-111: dice = 2 I / (sum(t + p) + 1e-7); loss = mean(1 - dice)
      const double U = st + sp + 1e-7;
      const double dice = 2.0 * spt / U;
      term += 1.0 - dice;
      a = 2.0 * spt / (U * U) / BC;
      b = -2.0 / U / BC;
      g = 0.0;
    } else if (kind == 2) {  // nets/custom_losses.py:114-133: mean((-log(clamp(dice, 1e-7, 1 - 1e-7)))^exp)
      const double U = st + sp + 1e-7;
      const double dice = 2.0 * spt / U;
      const double lo = 1e-7, hi = 1.0 - 1e-7;
      const double dc = fmin(fmax(dice, lo), hi);
      const double nl = -log(dc);
      term += pow(nl, param);
      // d term / d dice (zero where the clamp is active), then d dice / d p = 2 t / U - 2 I / U^2
      const double sl = (dice >= lo && dice <= hi) ? -param * pow(nl, param - 1.0) / dc : 0.0;
      a = sl * (-2.0 * spt / (U * U)) / BC;
      b = sl * (2.0 / U) / BC;
      g = 0.0;
    } else {  // nets/custom_losses.py:17-70: r = tp / sqrt(tt*pp + 1e-7) on centred data; loss = mean(1 - (r+1)/2)
      const double mt = st / N, mp = sp / N;
      const double tp = spt - N * mt * mp;
      const double tt = stt - N * mt * mt;
      const double pp = spp - N * mp * mp;
      const double q2 = tt * pp + 1e-7;
      const double q = sqrt(q2);
      const double r = tp / q;
      term += 1.0 - (r + 1.0) * 0.5;
      const double k = -0.5 / BC;
      const double w = tp * tt / (q2 * q);
      a = k * (-mt / q + w * mp);
      b = k / q;
      g = -k * w;
    }
    coef[bc * 3 + 0] = (float)a;
    coef[bc * 3 + 1] = (float)b;
    coef[bc * 3 + 2] = (float)g;
  }
  sterm[threadIdx.x] = term;
  __syncthreads();
  if (threadIdx.x == 0) {
    double s = 0.0;
    for (int i = 0; i < (int)blockDim.x; ++i) s += sterm[i];
    loss[0] = (float)(s / BC);
  }
}

// dy_pred = grad_loss * (alpha + beta * y_true + gamma * y_pred)
__global__ void __launch_bounds__(256) k_loss_bwd(const float* __restrict__ yp, const float* __restrict__ yt,
                                                  const float* __restrict__ coef, const float* __restrict__ grad_loss,
                                                  float* __restrict__ dyp, long N) {
  const long bc = blockIdx.y;
  const float gl = grad_loss ? __ldg(grad_loss) : 1.f;
  const float a = gl * coef[bc * 3 + 0], b = gl * coef[bc * 3 + 1], g = gl * coef[bc * 3 + 2];
  for (long i = blockIdx.x * 256L + threadIdx.x; i < N; i += (long)gridDim.x * 256L)
    dyp[bc * N + i] = a + b * __ldg(yt + bc * N + i) + g * __ldg(yp + bc * N + i);
}

// ------------------------------------------------------------------------------------------ cross entropy
// torch.nn.CrossEntropyLoss() as the reference configures it (experiments/run.py:105-110, called as
// loss_fn(y_pred, y_true) in train_test.py:159-160): the "logits" are the network's softmax OUTPUT p, the target is
// class probabilities t (one-hot floats), reduction 'mean' over batch x voxels:
//   loss = 1/(B N) sum_{b,v} [ (sum_c t_c) lse(p) - sum_c t_c p_c ],   d loss / d p_c = (softmax(p)_c sum t - t_c) / (B N)
// One pass over p and t (or uint8 labels, LAB = 1: t = one-hot(label)); y [B][C][N]; grid (chunks, B).
// V voxels per thread and iteration: V = 4 uses 16-byte loads (uchar4 for labels); needs N % 4 == 0 and aligned bases.
constexpr int kCeChunks = 592;  // blocks per sample: 4 per SM

template <int V>
__device__ __forceinline__ void ce_ldf(const float* __restrict__ base, long i, float (&out)[V]) {
  if constexpr (V == 4) {
    const float4 q = __ldg(reinterpret_cast<const float4*>(base) + i);
    out[0] = q.x, out[1] = q.y, out[2] = q.z, out[3] = q.w;
  } else {
    out[0] = __ldg(base + i);
  }
}

template <int C, int LAB, int V>
__device__ __forceinline__ void ce_load(const float* __restrict__ p, const float* __restrict__ t,
                                        const uint8_t* __restrict__ lab, long N, long i, float (&pv)[C][V],
                                        float (&tv)[C][V]) {
#pragma unroll
  for (int c = 0; c < C; ++c) ce_ldf<V>(p + c * N, i, pv[c]);
  if constexpr (LAB) {
    int l[V];
    if constexpr (V == 4) {
      const uchar4 q = __ldg(reinterpret_cast<const uchar4*>(lab) + i);
      l[0] = q.x, l[1] = q.y, l[2] = q.z, l[3] = q.w;
    } else {
      l[0] = (int)__ldg(lab + i);
    }
#pragma unroll
    for (int c = 0; c < C; ++c)
#pragma unroll
      for (int v = 0; v < V; ++v) tv[c][v] = l[v] == c ? 1.f : 0.f;
  } else {
#pragma unroll
    for (int c = 0; c < C; ++c) ce_ldf<V>(t + c * N, i, tv[c]);
  }
}

template <int C, int LAB, int V>
__global__ void __launch_bounds__(256) k_ce_fwd(const float* __restrict__ yp, const float* __restrict__ yt,
                                                const uint8_t* __restrict__ labels, double* __restrict__ partials,
                                                long N) {
  __shared__ double sred[8];
  const long b = blockIdx.y;
  const float* p = yp + b * C * N;
  const float* t = LAB ? nullptr : yt + b * C * N;
  const uint8_t* lab = LAB ? labels + b * N : nullptr;
  float acc = 0.f;
  double accd = 0.0;
  int cnt = 0;
  for (long i = blockIdx.x * 256L + threadIdx.x; i < N / V; i += (long)gridDim.x * 256L) {
    float pv[C][V], tv[C][V];
    ce_load<C, LAB, V>(p, t, lab, N, i, pv, tv);
#pragma unroll
    for (int v = 0; v < V; ++v) {
      float mx = pv[0][v];
#pragma unroll
      for (int c = 1; c < C; ++c) mx = fmaxf(mx, pv[c][v]);
      float se = 0.f, dot = 0.f, ts = 0.f;
#pragma unroll
      for (int c = 0; c < C; ++c) {
        se += expf(pv[c][v] - mx);
        dot = fmaf(tv[c][v], pv[c][v], dot);
        ts += tv[c][v];
      }
      acc += fmaf(ts, mx + logf(se), -dot);
    }
    if (++cnt == 64 / V) {  // bounded fp32 run length, then spill into fp64
      accd += (double)acc;
      acc = 0.f;
      cnt = 0;
    }
  }
  accd = warp_sum_d(accd + (double)acc);
  if ((threadIdx.x & 31) == 0) sred[threadIdx.x >> 5] = accd;
  __syncthreads();
  if (threadIdx.x == 0) {
    double s = 0.0;
    for (int w = 0; w < 8; ++w) s += sred[w];
    partials[b * gridDim.x + blockIdx.x] = s;
  }
}

__global__ void __launch_bounds__(256) k_ce_finalize(const double* __restrict__ partials, int n, double inv_count,
                                                     float* __restrict__ loss) {
  __shared__ double sred[8];
  double s = 0.0;
  for (int i = threadIdx.x; i < n; i += 256) s += partials[i];
  s = warp_sum_d(s);
  if ((threadIdx.x & 31) == 0) sred[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    double tot = 0.0;
    for (int w = 0; w < 8; ++w) tot += sred[w];
    loss[0] = (float)(tot * inv_count);
  }
}

template <int C, int LAB, int V>
__global__ void __launch_bounds__(256) k_ce_bwd(const float* __restrict__ yp, const float* __restrict__ yt,
                                                const uint8_t* __restrict__ labels,
                                                const float* __restrict__ grad_loss, float* __restrict__ dyp, long N,
                                                float inv_count) {
  const long b = blockIdx.y;
  const float* p = yp + b * C * N;
  const float* t = LAB ? nullptr : yt + b * C * N;
  const uint8_t* lab = LAB ? labels + b * N : nullptr;
  float* d = dyp + b * C * N;
  const float k = (grad_loss ? __ldg(grad_loss) : 1.f) * inv_count;
  for (long i = blockIdx.x * 256L + threadIdx.x; i < N / V; i += (long)gridDim.x * 256L) {
    float pv[C][V], tv[C][V];
    ce_load<C, LAB, V>(p, t, lab, N, i, pv, tv);
#pragma unroll
    for (int v = 0; v < V; ++v) {
      float mx = pv[0][v];
#pragma unroll
      for (int c = 1; c < C; ++c) mx = fmaxf(mx, pv[c][v]);
      float se = 0.f, ts = 0.f;
#pragma unroll
      for (int c = 0; c < C; ++c) {
        pv[c][v] = expf(pv[c][v] - mx);
        se += pv[c][v];
        ts += tv[c][v];
      }
      const float r = ts / se;
#pragma unroll
      for (int c = 0; c < C; ++c) pv[c][v] = k * fmaf(pv[c][v], r, -tv[c][v]);
    }
#pragma unroll
    for (int c = 0; c < C; ++c) {
      if constexpr (V == 4) {
        reinterpret_cast<float4*>(d + c * N)[i] = make_float4(pv[c][0], pv[c][1], pv[c][2], pv[c][3]);
      } else {
        d[c * N + i] = pv[c][0];
      }
    }
  }
}

// ------------------------------------------------------------------------------------------ launchers
#define HNO_CLASS_SWITCH(C, ...) \
  switch (C) {                    \
    case 1: { constexpr int kC = 1; __VA_ARGS__ } break; \
    case 2: { constexpr int kC = 2; __VA_ARGS__ } break; \
    case 3: { constexpr int kC = 3; __VA_ARGS__ } break; \
    case 4: { constexpr int kC = 4; __VA_ARGS__ } break; \
    case 5: { constexpr int kC = 5; __VA_ARGS__ } break; \
    case 6: { constexpr int kC = 6; __VA_ARGS__ } break; \
    case 7: { constexpr int kC = 7; __VA_ARGS__ } break; \
    case 8: { constexpr int kC = 8; __VA_ARGS__ } break; \
    default: set_error("head: number of classes %d not in [1, %d]", C, kMaxClasses); return -1; \
  }

int head_forward(const void* th, const void* td, const float* ll, float* probs, int B, int C, long P, int activation,
                 cudaStream_t st) {
  InterpDev t;
  if (make_dev(th, td, &t)) return -1;
  HNO_CHECK(ll && probs, "head_forward: null pointer");
  HNO_CHECK(P >= (long)t.lo[1] * t.lo[2], "head_forward: plane pitch too small");
  const long Nx = (long)t.hi[0] * t.hi[1] * t.hi[2];
  dim3 grid(ceil_div(Nx, 256), B);
  HNO_CLASS_SWITCH(C, {
    if (activation == 1) k_head_fwd<kC, 1><<<grid, 256, 0, st>>>(ll, probs, t, P);
    else k_head_fwd<kC, 0><<<grid, 256, 0, st>>>(ll, probs, t, P);
  })
  HNO_LAUNCH_CHECK();
  return 0;
}

int head_argmax(const void* th, const void* td, const float* ll, uint8_t* labels, int B, int C, long P,
                cudaStream_t st) {
  InterpDev t;
  if (make_dev(th, td, &t)) return -1;
  HNO_CHECK(ll && labels, "head_argmax: null pointer");
  HNO_CHECK(P >= (long)t.lo[1] * t.lo[2], "head_argmax: plane pitch too small");
  const long Nx = (long)t.hi[0] * t.hi[1] * t.hi[2];
  dim3 grid(ceil_div(Nx, 256), B);
  HNO_CLASS_SWITCH(C, { k_head_argmax<kC><<<grid, 256, 0, st>>>(ll, labels, t, P); })
  HNO_LAUNCH_CHECK();
  return 0;
}

size_t head_backward_workspace_bytes(const void* th, int B, int C) {
  if (!th) return 0;
  const auto* h = reinterpret_cast<const InterpHeader*>(th);
  const size_t g1 = (size_t)B * C * h->hi[0] * h->hi[1] * h->lo[2];
  const size_t g2 = (size_t)B * C * h->hi[0] * h->lo[1] * h->lo[2];
  const size_t mom = (size_t)B * C * kLossChunks * kMoments * sizeof(double);
  return (g1 + g2) * sizeof(float) + mom + 512;
}

// longest tap range of a low-resolution index along `axis` (from the host copy of the tables)
static int max_tap_range(const void* th, int axis) {
  const auto* h = reinterpret_cast<const InterpHeader*>(th);
  const int* iw = reinterpret_cast<const int*>(th);
  int m = 0;
  for (int i = 0; i < h->lo[axis]; ++i) {
    const int n = iw[h->off_e[axis] + i] - iw[h->off_s[axis] + i];
    if (n > m) m = n;
  }
  return m;
}

static int head_backward_passes(const void* th, const InterpDev& t, float* g1, float* g2, float* dll, int B, int C, long P,
                                cudaStream_t st) {
  const long nbc = (long)B * C;
  HNO_CHECK(nbc * t.hi[0] <= 65535 && t.lo[0] <= 65535 && nbc <= 65535, "head backward: grid too large");
  static const bool gather_on = !(getenv("HNO_HEAD_GATHER") && atoi(getenv("HNO_HEAD_GATHER")) == 0);
  const int kh = max_tap_range(th, 1), kd = max_tap_range(th, 0);
  const bool gh = gather_on && kh <= kGatherMax && (size_t)t.lo[1] * (kGatherMax + 1) * 4 <= 40 * 1024;
  const bool gd = gather_on && kd <= kGatherMax;
  if (gh) {
    const long R = nbc * t.hi[0];
    dim3 grid(ceil_div((long)t.lo[1] * t.lo[2], 256), (unsigned)ceil_div(R, 4));
    if (kh <= 4)
      k_head_bwd_h_gather<4><<<grid, 256, (size_t)t.lo[1] * 5 * 4, st>>>(g1, g2, t, R);
    else
      k_head_bwd_h_gather<6><<<grid, 256, (size_t)t.lo[1] * 7 * 4, st>>>(g1, g2, t, R);
    HNO_LAUNCH_CHECK();
  } else {
    dim3 grid(ceil_div((long)t.lo[1] * t.lo[2], 256), (unsigned)(nbc * t.hi[0]));
    k_head_bwd_h<<<grid, 256, 0, st>>>(g1, g2, t, nbc);
    HNO_LAUNCH_CHECK();
  }
  if (gd) {
    dim3 grid(ceil_div(P, 256), t.lo[0], (unsigned)ceil_div(nbc, 4));
    if (kd <= 4)
      k_head_bwd_d_gather<4><<<grid, 256, 0, st>>>(g2, dll, t, P, nbc);
    else
      k_head_bwd_d_gather<6><<<grid, 256, 0, st>>>(g2, dll, t, P, nbc);
    HNO_LAUNCH_CHECK();
  } else {
    dim3 grid(ceil_div(P, 256), t.lo[0], (unsigned)nbc);
    k_head_bwd_d<<<grid, 256, 0, st>>>(g2, dll, t, nbc, P);
    HNO_LAUNCH_CHECK();
  }
  return 0;
}

int head_backward(const void* th, const void* td, const float* dprobs, const float* probs, float* dll, void* ws, int B,
                  int C, long P, int activation, cudaStream_t st) {
  InterpDev t;
  if (make_dev(th, td, &t)) return -1;
  HNO_CHECK(dprobs && dll && ws && (activation == 0 || probs), "head_backward: null pointer");
  float* g1 = reinterpret_cast<float*>(ws);
  float* g2 = g1 + (size_t)B * C * t.hi[0] * t.hi[1] * t.lo[2];
  dim3 grid(ceil_div((long)t.hi[0] * t.hi[1] * t.lo[2], 256), B);
  HNO_CLASS_SWITCH(C, {
    if (activation == 1)
      k_head_bwd_w<kC, 1, 0><<<grid, 256, 0, st>>>(dprobs, probs, nullptr, nullptr, nullptr, nullptr, g1, t, P);
    else
      k_head_bwd_w<kC, 0, 0><<<grid, 256, 0, st>>>(dprobs, probs, nullptr, nullptr, nullptr, nullptr, g1, t, P);
  })
  HNO_LAUNCH_CHECK();
  return head_backward_passes(th, t, g1, g2, dll, B, C, P, st);
}

size_t loss_workspace_bytes(int B, int C) { return (size_t)B * C * kLossChunks * kMoments * sizeof(double) + 256; }

int loss_forward(const float* yp, const float* yt, float* loss, float* coef, void* ws, int B, int C, long N, int kind,
                 float param, cudaStream_t st) {
  HNO_CHECK(yp && yt && loss && coef && ws, "loss_forward: null pointer");
  HNO_CHECK(kind >= 0 && kind <= 2, "loss_forward: kind must be 0 (Dice), 1 (PCC) or 2 (ExpDice)");
  HNO_CHECK(kind != 2 || param > 0.f, "loss_forward: ExpDice needs a positive exponent");
  HNO_CHECK((long)B * C <= 65535, "loss_forward: too many (batch, label) pairs");
  double* partials = reinterpret_cast<double*>(ws);
  int chunks = (int)((N + 255) / 256 < kLossChunks ? (N + 255) / 256 : kLossChunks);
  dim3 grid(chunks, B * C);
  if (N % 4 == 0 && ((reinterpret_cast<uintptr_t>(yp) | reinterpret_cast<uintptr_t>(yt)) & 15) == 0)
    k_loss_moments<4><<<grid, 256, 0, st>>>(yp, yt, partials, N);
  else
    k_loss_moments<1><<<grid, 256, 0, st>>>(yp, yt, partials, N);
  HNO_LAUNCH_CHECK();
  k_loss_finalize<<<1, 256, 0, st>>>(partials, chunks, B * C, (double)N, kind, (double)param, loss, coef);
  HNO_LAUNCH_CHECK();
  return 0;
}

int loss_backward(const float* yp, const float* yt, const float* coef, const float* grad_loss, float* dyp, int B, int C,
                  long N, cudaStream_t st) {
  HNO_CHECK(yp && yt && coef && dyp, "loss_backward: null pointer");
  int chunks = (int)((N + 255) / 256 < 1184 ? (N + 255) / 256 : 1184);
  dim3 grid(chunks, B * C);
  k_loss_bwd<<<grid, 256, 0, st>>>(yp, yt, coef, grad_loss, dyp, N);
  HNO_LAUNCH_CHECK();
  return 0;
}

size_t ce_loss_workspace_bytes(int B) { return (size_t)B * kCeChunks * sizeof(double) + 256; }

// 16-byte path: four voxels per thread; needs every channel row (pitch N floats) and label row (N bytes) aligned
static bool ce_vec4(const void* a, const void* b, const void* c, const void* d, long N, int C) {
  const uintptr_t bits = reinterpret_cast<uintptr_t>(a) | reinterpret_cast<uintptr_t>(b) |
                         reinterpret_cast<uintptr_t>(c) | reinterpret_cast<uintptr_t>(d);
  return N % 4 == 0 && (bits & 15) == 0 && C <= 4;
}

int ce_loss_forward(const float* yp, const float* yt, const uint8_t* labels, float* loss, void* ws, int B, int C,
                    long N, cudaStream_t st) {
  HNO_CHECK(yp && loss && ws, "ce_loss_forward: null pointer");
  HNO_CHECK((yt != nullptr) != (labels != nullptr), "ce_loss_forward: pass exactly one of y_true and labels");
  HNO_CHECK(B >= 1 && B <= 65535 && N >= 1, "ce_loss_forward: bad sizes");
  double* partials = reinterpret_cast<double*>(ws);
  const bool v4 = ce_vec4(yp, yt, labels, nullptr, N, C);
  const long items = v4 ? N / 4 : N;
  const int chunks = (int)((items + 255) / 256 < kCeChunks ? (items + 255) / 256 : kCeChunks);
  dim3 grid(chunks, B);
  HNO_CLASS_SWITCH(C, {
    if constexpr (kC <= 4) {
      if (v4 && labels) k_ce_fwd<kC, 1, 4><<<grid, 256, 0, st>>>(yp, nullptr, labels, partials, N);
      else if (v4) k_ce_fwd<kC, 0, 4><<<grid, 256, 0, st>>>(yp, yt, nullptr, partials, N);
    }
    if (!v4 && labels) k_ce_fwd<kC, 1, 1><<<grid, 256, 0, st>>>(yp, nullptr, labels, partials, N);
    else if (!v4) k_ce_fwd<kC, 0, 1><<<grid, 256, 0, st>>>(yp, yt, nullptr, partials, N);
  })
  HNO_LAUNCH_CHECK();
  k_ce_finalize<<<1, 256, 0, st>>>(partials, chunks * B, 1.0 / ((double)B * (double)N), loss);
  HNO_LAUNCH_CHECK();
  return 0;
}

int ce_loss_backward(const float* yp, const float* yt, const uint8_t* labels, const float* grad_loss, float* dyp, int B,
                     int C, long N, cudaStream_t st) {
  HNO_CHECK(yp && dyp, "ce_loss_backward: null pointer");
  HNO_CHECK((yt != nullptr) != (labels != nullptr), "ce_loss_backward: pass exactly one of y_true and labels");
  HNO_CHECK(B >= 1 && B <= 65535 && N >= 1, "ce_loss_backward: bad sizes");
  const bool v4 = ce_vec4(yp, yt, labels, dyp, N, C);
  const long items = v4 ? N / 4 : N;
  const int chunks = (int)((items + 255) / 256 < 1184 ? (items + 255) / 256 : 1184);
  dim3 grid(chunks, B);
  const float inv = (float)(1.0 / ((double)B * (double)N));
  HNO_CLASS_SWITCH(C, {
    if constexpr (kC <= 4) {
      if (v4 && labels) k_ce_bwd<kC, 1, 4><<<grid, 256, 0, st>>>(yp, nullptr, labels, grad_loss, dyp, N, inv);
      else if (v4) k_ce_bwd<kC, 0, 4><<<grid, 256, 0, st>>>(yp, yt, nullptr, grad_loss, dyp, N, inv);
    }
    if (!v4 && labels) k_ce_bwd<kC, 1, 1><<<grid, 256, 0, st>>>(yp, nullptr, labels, grad_loss, dyp, N, inv);
    else if (!v4) k_ce_bwd<kC, 0, 1><<<grid, 256, 0, st>>>(yp, yt, nullptr, grad_loss, dyp, N, inv);
  })
  HNO_LAUNCH_CHECK();
  return 0;
}

int head_loss_forward(const void* th, const void* td, const float* ll, const uint8_t* labels, float* loss, float* coef,
                      void* ws, int B, int C, long P, int kind, float param, cudaStream_t st) {
  InterpDev t;
  if (make_dev(th, td, &t)) return -1;
  HNO_CHECK(ll && labels && loss && coef && ws, "head_loss_forward: null pointer");
  HNO_CHECK(kind >= 0 && kind <= 2, "head_loss_forward: kind must be 0 (Dice), 1 (PCC) or 2 (ExpDice)");
  HNO_CHECK(kind != 2 || param > 0.f, "head_loss_forward: ExpDice needs a positive exponent");
  const long Nx = (long)t.hi[0] * t.hi[1] * t.hi[2];
  // moments live at the END of the head-backward workspace so that forward and backward can share it
  const size_t g12 = ((size_t)B * C * t.hi[0] * t.hi[1] * t.lo[2] + (size_t)B * C * t.hi[0] * t.lo[1] * t.lo[2]);
  double* partials = reinterpret_cast<double*>(reinterpret_cast<char*>(ws) + ((g12 * sizeof(float) + 255) & ~(size_t)255));
  int chunks = (int)((Nx + 255) / 256 < kLossChunks ? (Nx + 255) / 256 : kLossChunks);
  dim3 grid(chunks, B);
  if (const int kmax = wrow_kmax(t)) {
    const long rows = (long)t.hi[0] * t.hi[1];
    chunks = (int)((rows + kWRowWarps - 1) / kWRowWarps < kLossChunks ? (rows + kWRowWarps - 1) / kWRowWarps : kLossChunks);
    grid = dim3(chunks, B);
    const int WP = (t.lo[2] + 3) & ~3;
    const size_t smem = (size_t)kWRowWarps * C * WP * sizeof(float);
    HNO_CLASS_SWITCH(C, {
      if (kmax == 5) k_head_loss_wrow<kC, 1, 5><<<grid, 32 * kWRowWarps, smem, st>>>(ll, labels, partials, t, P, WP);
      else k_head_loss_wrow<kC, 1, 10><<<grid, 32 * kWRowWarps, smem, st>>>(ll, labels, partials, t, P, WP);
    })
  } else if (row_kernels_enabled() && row_kernels_tap_group(th) > 0) {
    const long rows = (long)t.hi[0] * t.hi[1];
    chunks = (int)(rows < kLossChunks ? rows : kLossChunks);
    grid = dim3(chunks, B);
    HNO_CLASS_SWITCH(C, { k_head_loss_rows<kC, 1><<<grid, 32 * kRowWarps, 0, st>>>(ll, labels, partials, t, P); })
  } else {
    HNO_CLASS_SWITCH(C, { k_head_loss_moments<kC, 1><<<grid, 256, 0, st>>>(ll, labels, partials, t, P); })
  }
  HNO_LAUNCH_CHECK();
  k_loss_finalize<<<1, 256, 0, st>>>(partials, chunks, B * C, (double)Nx, kind, (double)param, loss, coef);
  HNO_LAUNCH_CHECK();
  return 0;
}

int head_loss_backward(const void* th, const void* td, const float* ll, const uint8_t* labels, const float* coef,
                       const float* grad_loss, float* dll, void* ws, int B, int C, long P, cudaStream_t st) {
  InterpDev t;
  if (make_dev(th, td, &t)) return -1;
  HNO_CHECK(ll && labels && coef && dll && ws, "head_loss_backward: null pointer");
  float* g1 = reinterpret_cast<float*>(ws);
  float* g2 = g1 + (size_t)B * C * t.hi[0] * t.hi[1] * t.lo[2];
  const int tg = row_kernels_enabled() ? row_kernels_tap_group(th) : 0;
  if (const int kmax = wrow_kmax(t)) {
    const long rows = (long)t.hi[0] * t.hi[1];
    const long want = (rows + kWRowWarps - 1) / kWRowWarps;
    dim3 grid((int)(want < 4 * 296 ? want : 4 * 296), B);
    const int WP = (t.lo[2] + 3) & ~3, WXP = (t.hi[2] + 3) & ~3;
    static const bool gather_on = !(getenv("HNO_HEAD_GATHER") && atoi(getenv("HNO_HEAD_GATHER")) == 0);
    const int kw = max_tap_range(th, 2);
    const int kg = !gather_on || kw > kGatherMax ? 0 : (kw <= 4 ? 4 : 6);
    const size_t smem = (size_t)kWRowWarps * C * (WP + WXP) * sizeof(float) + (size_t)t.lo[2] * (kg + 1) * sizeof(float);
#define HNO_BWD_WROW(KMAX_, KG_)                                                                             \
  {                                                                                                          \
    auto kern = k_head_bwd_wrow<kC, 1, KMAX_, KG_>;                                                          \
    HNO_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));            \
    kern<<<grid, 32 * kWRowWarps, smem, st>>>(ll, labels, coef, grad_loss, g1, t, P, WP, WXP);               \
  }
    HNO_CLASS_SWITCH(C, {
      if (kmax == 5) {
        if (kg == 4) HNO_BWD_WROW(5, 4) else if (kg == 6) HNO_BWD_WROW(5, 6) else HNO_BWD_WROW(5, 0)
      } else {
        if (kg == 4) HNO_BWD_WROW(10, 4) else if (kg == 6) HNO_BWD_WROW(10, 6) else HNO_BWD_WROW(10, 0)
      }
    })
#undef HNO_BWD_WROW
  } else if (tg > 0) {
    const long rows = (long)t.hi[0] * t.hi[1];
    dim3 grid((int)(rows < 4 * 296 ? rows : 4 * 296), B);
    HNO_CLASS_SWITCH(C, {
      k_head_bwd_w_rows<kC, 1><<<grid, 32 * kRowWarps, 0, st>>>(ll, labels, coef, grad_loss, g1, t, P, tg);
    })
  } else {
    dim3 grid(ceil_div((long)t.hi[0] * t.hi[1] * t.lo[2], 256), B);
    HNO_CLASS_SWITCH(C, {
      k_head_bwd_w<kC, 1, 1><<<grid, 256, 0, st>>>(nullptr, nullptr, ll, labels, coef, grad_loss, g1, t, P);
    })
  }
  HNO_LAUNCH_CHECK();
  return head_backward_passes(th, t, g1, g2, dll, B, C, P, st);
}

// ------------------------------------------------------------------------------------------ full-resolution head
// use_resize=False (nets/hnosegxs.py:102-109, 150, 174-180; nets/architectures.py:286-289, 345-351): the network runs at
// the image resolution, there is no interpolation, and the head is conv_out (hno_pwconv_forward) followed by the output
// activation.  These kernels are that activation between the planar logits [B][C][D][P] and the dense probabilities
// [B][C][D][H][W]: one thread per planar position, all classes in registers, coalesced along the plane.
template <int C, int ACT>
__global__ void __launch_bounds__(256) k_head_direct_fwd(const float* __restrict__ ll, float* __restrict__ probs,
                                                         uint8_t* __restrict__ labels, long DP, long P, long HW) {
  // grid (plane chunks, D, B): no integer division in the index arithmetic (218 -> ~60 warp instructions per voxel)
  const long col = blockIdx.x * 256L + threadIdx.x;
  if (col >= HW) return;
  const long d = blockIdx.y;
  const int b = blockIdx.z;
  const long j = d * P + col;
  const long v = d * HW + col, N = (long)gridDim.y * HW;
  float lg[C];
#pragma unroll
  for (int c = 0; c < C; ++c) lg[c] = __ldg(ll + ((long)b * C + c) * DP + j);
  if (labels != nullptr) {  // np.argmax: the first maximum
    int best = 0;
#pragma unroll
    for (int c = 1; c < C; ++c)
      if (lg[c] > lg[best]) best = c;
    labels[(long)b * N + v] = (uint8_t)best;
    return;
  }
  if (ACT == 1) softmax_inplace<C>(lg);
#pragma unroll
  for (int c = 0; c < C; ++c) probs[((long)b * C + c) * N + v] = lg[c];
}

// d logits_c = p_c (dp_c - sum_k dp_k p_k) for the softmax, dp_c without activation; padding columns of the planes get 0
template <int C, int ACT>
__global__ void __launch_bounds__(256) k_head_direct_bwd(const float* __restrict__ dprobs,
                                                         const float* __restrict__ probs, float* __restrict__ dll,
                                                         long DP, long P, long HW) {
  const long col = blockIdx.x * 256L + threadIdx.x;
  if (col >= P) return;
  const long d = blockIdx.y;
  const int b = blockIdx.z;
  const long j = d * P + col;
  float g[C];
  if (col < HW) {
    const long v = d * HW + col, N = (long)gridDim.y * HW;
    float dot = 0.f, p[C];
#pragma unroll
    for (int c = 0; c < C; ++c) {
      g[c] = __ldg(dprobs + ((long)b * C + c) * N + v);
      if (ACT == 1) {
        p[c] = __ldg(probs + ((long)b * C + c) * N + v);
        dot = fmaf(g[c], p[c], dot);
      }
    }
    if (ACT == 1) {
#pragma unroll
      for (int c = 0; c < C; ++c) g[c] = p[c] * (g[c] - dot);
    }
  } else {
#pragma unroll
    for (int c = 0; c < C; ++c) g[c] = 0.f;
  }
#pragma unroll
  for (int c = 0; c < C; ++c) dll[((long)b * C + c) * DP + j] = g[c];
}

static int head_direct_check(const char* who, int B, int D, int H, int W, long P) {
  HNO_CHECK(B >= 1 && B <= 65535 && D >= 1 && D <= 65535 && H >= 1 && W >= 1, "%s: bad sizes", who);
  HNO_CHECK(P >= (long)H * W, "%s: plane pitch too small", who);
  return 0;
}

int head_direct_forward(const float* ll, float* probs, uint8_t* labels, int B, int C, int D, int H, int W, long P,
                        int activation, cudaStream_t st) {
  HNO_CHECK(ll && (probs != nullptr) != (labels != nullptr), "head_direct_forward: pass logits and exactly one output");
  HNO_CHECK(activation == 0 || activation == 1, "head_direct_forward: activation must be 0 (none) or 1 (softmax)");
  if (head_direct_check("head_direct_forward", B, D, H, W, P)) return -1;
  const long DP = (long)D * P, HW = (long)H * W;
  dim3 grid(ceil_div(HW, 256), D, B);
  HNO_CLASS_SWITCH(C, {
    if (activation == 1) k_head_direct_fwd<kC, 1><<<grid, 256, 0, st>>>(ll, probs, labels, DP, P, HW);
    else k_head_direct_fwd<kC, 0><<<grid, 256, 0, st>>>(ll, probs, labels, DP, P, HW);
  })
  HNO_LAUNCH_CHECK();
  return 0;
}

int head_direct_backward(const float* dprobs, const float* probs, float* dll, int B, int C, int D, int H, int W,
                         long P, int activation, cudaStream_t st) {
  HNO_CHECK(dprobs && dll && (activation == 0 || probs), "head_direct_backward: null pointer");
  HNO_CHECK(activation == 0 || activation == 1, "head_direct_backward: activation must be 0 (none) or 1 (softmax)");
  if (head_direct_check("head_direct_backward", B, D, H, W, P)) return -1;
  const long DP = (long)D * P, HW = (long)H * W;
  dim3 grid(ceil_div(P, 256), D, B);
  HNO_CLASS_SWITCH(C, {
    if (activation == 1) k_head_direct_bwd<kC, 1><<<grid, 256, 0, st>>>(dprobs, probs, dll, DP, P, HW);
    else k_head_direct_bwd<kC, 0><<<grid, 256, 0, st>>>(dprobs, probs, dll, DP, P, HW);
  })
  HNO_LAUNCH_CHECK();
  return 0;
}

}  // namespace hno
