// The frequency-domain core of an HNO-XS block as ONE kernel (sm_100a):
//
//     T2 (output of the H analysis)  ->  W analysis + cas recombination  ->  z_0
//         ->  n_XS x [ z_l = selu(W_l z_{l-1} + z_{l-1}) ]                            nets/hnosegxs.py:261-262, 307-329
//         ->  recombination^T + W synthesis  ->  T2 (input of the H synthesis)
//
// replacing k_dht_tail_fwd -> k_modechain_fwd -> k_dht_tail_adj (and, in the backward pass, k_dht_tail_fwd -> k_modechain_bwd
// -> k_dht_tail_adj), three dependent launches of 19 + 18 + 21 us (44 us for the backward chain) on a few MB of L2-resident
// data -- 1.9 ms of a 10.3 ms training step went into these and the two H stages around them.
//
// Why one kernel is possible without a grid-wide barrier: the retained modes decouple.  A retained frequency k_d on the D
// axis only touches the cos / sin rows of ITS |u_d| (dht_plan.h), likewise on H, and the shared-weight mix acts per mode.  So
// the set { (b, c, k_d in +-u_d, k_h in +-u_h, every k_w) } for one (b, |u_d|, |u_h|) is closed under all three steps: a CTA
// owns it for ALL channels, reads the <= 4 rows (cos/sin of u_d) x (cos/sin of u_h) of T2 per channel and writes the same
// rows back (in place).  Grid = JC_h x JC_d x B (15 x 11 x 2 = 330 CTAs at the BASELINE config), 50 KB (forward) / 72 KB
// (backward) of shared memory and 80 registers: three CTAs per SM, one wave.
#include "common.cuh"
#include "dht_plan.h"

#include <stdio.h>
#include <stdlib.h>

namespace hno {

constexpr int kCoreThreads = 256;
constexpr int kCoreMaxLayers = 8;

struct CoreGeom {
  int W, Wp, Jw, Jwp, Ld, Lh, Lw, JCw;
  long rowlen, slablen;             // T2 [slab][jh][jd][Wp]
  int off_fullT_w, off_fullP_w;     // [Wp][Jwp] and [Jw][Wp] copies of the W rows
  int off_kdesc[3], off_jdesc[3];
  int sin_d[32], sin_h[32];         // sin row of the same |u| for every cos row (or -1)
  int nq, nqp;                      // 4 * Lw local modes, rounded up to a multiple of 4
};

struct CorePtrs {
  const float* w[kCoreMaxLayers];
};

// zall: [L + 1][B][C][Ld][Lh][Lw] (z_0 .. z_L): written by the forward when non-null, read by the backward.
__device__ long long g_core_prof[16];
#define CORE_STAMP(i) if (blockIdx.x == 3 && blockIdx.y == 3 && blockIdx.z == 0 && threadIdx.x == 0) g_core_prof[i] = clock64();
template <int C, bool BWD>
__global__ void __launch_bounds__(kCoreThreads, 3) k_spectral_core(float* __restrict__ T2, float* __restrict__ zall,
                                                               float* __restrict__ partials, const float* __restrict__ pf,
                                                               const int* __restrict__ pi, const CoreGeom g,
                                                               const CorePtrs P, int L, int B, float scale_in,
                                                               float scale_out) {
  // programmatic dependent launch: the H-synthesis kernel behind this one may run its prologue (basis image, TMEM, barriers)
  // now; this kernel's own prologue (tables, weights) runs while the H analysis in front of it drains
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  CORE_STAMP(0)
  extern __shared__ float4 smem4[];
  float* s = reinterpret_cast<float*>(smem4);
  float* fwP = s;                               // [Jw][Wp]
  float* fwT = fwP + g.Jw * g.Wp;               // [Wp][Jwp]     fwT | Tw are dead while the mixes run: the backward pass
  float* Tw = fwT + g.Wp * g.Jwp;               // [C][4][Jwp]   parks a staging tile there (>= C (nqp + 4) floats, checked on the host)
  float* zq = Tw + C * 4 * g.Jwp;               // [C][nqp]
  float* wts = zq + C * g.nqp;                  // [L][C][C]   forward: transposed [i][o]; backward: as stored [o][i]
  float* sdp = wts + L * C * C;                 // backward only: [C][nqp + 4] x 2 staging tiles (d(pre), z_l)
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int jch = blockIdx.x, jcd = blockIdx.y, b = blockIdx.z;
  const int* jd_desc = pi + g.off_jdesc[0];
  const int* jh_desc = pi + g.off_jdesc[1];
  const int* jw_desc = pi + g.off_jdesc[2];
  const int* kd_desc = pi + g.off_kdesc[0];
  const int* kh_desc = pi + g.off_kdesc[1];
  const int* kw_desc = pi + g.off_kdesc[2];
  const int rd[2] = {jcd, g.sin_d[jcd]}, rh[2] = {jch, g.sin_h[jch]};
  const int kd2[2] = {jd_desc[4 * jcd], jd_desc[4 * jcd + 1]};  // positions of +u_d / -u_d in the retained list (or -1)
  const int kh2[2] = {jh_desc[4 * jch], jh_desc[4 * jch + 1]};
  const long BCM = (long)B * C * g.Ld * g.Lh * g.Lw;

  // ---- tables, weights and the <= 4 rows of every channel
  {
    const float* src = pf + g.off_fullT_w;  // fullT [Wp][Jwp], then fullP [Jw][Wp] (adjacent in the plan)
    const unsigned dT = (unsigned)__cvta_generic_to_shared(fwT), dP = (unsigned)__cvta_generic_to_shared(fwP);
    const int nT = (g.Wp * g.Jwp) >> 2, nP = (g.Jw * g.Wp) >> 2;
    for (int c4 = tid; c4 < nT + nP; c4 += kCoreThreads)
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(c4 < nT ? dT + 16 * c4 : dP + 16 * (c4 - nT)), "l"(src + 4 * c4) : "memory");
  }
  asm volatile("cp.async.commit_group;" ::: "memory");
  // backward: the saved layer outputs z_L (-> sy) and z_{L-1} (-> sx) of this CTA's modes start their way from HBM now
  // (they were written by the forward pass, not by the kernel in front, so they may be requested before griddepcontrol.wait);
  // fetched at the top of every layer they cost ~12,000 cycles per layer (cycle counters: 3 x 13 k of the kernel's 107 k)
  constexpr int CHq = C / 2;
  const int pq = tid & 127, phalf = tid >> 7;
  const int TVSq = g.nqp + 4;
  const long cstr = (long)g.Ld * g.Lh * g.Lw;
  bool pvalid = false;
  long pgoff = 0;
  {
    const bool act = pq < g.nq;
    const int ms = act ? pq / g.Lw : 0, kw = act ? pq - ms * g.Lw : 0;
    const int kd = kd2[ms >> 1], kh = kh2[ms & 1];
    pvalid = act && kd >= 0 && kh >= 0;
    pgoff = pvalid ? (((long)b * C * g.Ld + kd) * g.Lh + kh) * g.Lw + kw + (long)(phalf * CHq) * cstr : 0;
  }
  auto prefetch_layer = [&](int layer, float* dst) {  // z_layer of this CTA's modes -> dst[c][TVS] (zeros for missing modes)
    if (pq < g.nqp) {
      const float* src = zall + (long)layer * BCM + pgoff;
      const unsigned d0 = (unsigned)__cvta_generic_to_shared(dst + (phalf * CHq) * TVSq + pq);
#pragma unroll
      for (int c = 0; c < CHq; ++c)
        asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(d0 + 4u * (c * TVSq)), "l"(src + c * cstr),
                     "r"(pvalid ? 4 : 0)
                     : "memory");
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };
  if (BWD) {
    prefetch_layer(L, sdp);                  // z_L lands where d(pre) of the first layer is formed (in place)
    prefetch_layer(L - 1, sdp + C * TVSq);
  }
  for (int idx = tid; idx < L * C * C; idx += kCoreThreads) {
    const int l = idx / (C * C), r = idx - l * C * C;
    if (BWD) {
      wts[idx] = __ldg(P.w[l] + r);
    } else {
      const int i = r / C, o = r - i * C;  // wts[l][i][o] = W_l[o][i]
      wts[idx] = __ldg(P.w[l] + o * C + i);
    }
  }
  if (BWD) asm volatile("cp.async.wait_group 2;" ::: "memory");  // the tables; the two layer tiles stay in flight
  else cp_async_wait_all();
  CORE_STAMP(1)
  asm volatile("griddepcontrol.wait;" ::: "memory");
  CORE_STAMP(2)  // T2 (the H analysis / H^T synthesis in front) is read from here on
  __syncthreads();
  // ---- W analysis: Tw[c][r][jw] = sum_w rows[c][r][w] fwT[w][jw]; a thread owns the 4 rows of a channel x 4 columns jw
  {
    const int jq = g.Jwp >> 2;
    for (int item = tid; item < C * jq; item += kCoreThreads) {
      const int c = item / jq, j4 = (item - c * jq) * 4;
      // the (cos|sin d) x (cos|sin h) rows of this channel in T2 (L2 resident: the H analysis just wrote them); the 8 threads
      // of a channel read the same addresses.  A missing sin row reads row 0 with weight 0.
      const float* rp[4];
      float live[4];
#pragma unroll
      for (int r = 0; r < 4; ++r) {
        const int jd = rd[r >> 1], jh = rh[r & 1];
        live[r] = (jd >= 0 && jh >= 0) ? 1.f : 0.f;
        rp[r] = T2 + (long)(b * C + c) * g.slablen + (long)(jh >= 0 ? jh : 0) * g.rowlen + (long)(jd >= 0 ? jd : 0) * g.Wp;
      }
      float2 acc[4][2];
#pragma unroll
      for (int r = 0; r < 4; ++r) acc[r][0] = acc[r][1] = make_float2(0.f, 0.f);
      for (int w4 = 0; w4 < g.Wp; w4 += 4) {
        float4 x[4];
#pragma unroll
        for (int r = 0; r < 4; ++r) {
          x[r] = *reinterpret_cast<const float4*>(rp[r] + w4);
          // pad columns (w >= W) hold whatever the workspace held: never let them reach the sums (0 * NaN)
          if (w4 + 1 >= g.W) x[r].y = 0.f;
          if (w4 + 2 >= g.W) x[r].z = 0.f;
          if (w4 + 3 >= g.W) x[r].w = 0.f;
          if (live[r] == 0.f) x[r] = make_float4(0.f, 0.f, 0.f, 0.f);
        }
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const float4 f = *reinterpret_cast<const float4*>(fwT + (w4 + k) * g.Jwp + j4);
          const float2 f01 = make_float2(f.x, f.y), f23 = make_float2(f.z, f.w);
#pragma unroll
          for (int r = 0; r < 4; ++r) {
            const float xv = k == 0 ? x[r].x : (k == 1 ? x[r].y : (k == 2 ? x[r].z : x[r].w));
            acc[r][0] = ffma2(dup2(xv), f01, acc[r][0]);
            acc[r][1] = ffma2(dup2(xv), f23, acc[r][1]);
          }
        }
      }
#pragma unroll
      for (int r = 0; r < 4; ++r)
        *reinterpret_cast<float4*>(Tw + (c * 4 + r) * g.Jwp + j4) = make_float4(acc[r][0].x, acc[r][0].y, acc[r][1].x, acc[r][1].y);
    }
  }
  __syncthreads();
  CORE_STAMP(3)

  // ---- cas recombination (dht_plan.h): Z[kd, kh, kw] = scale * sum over the 8 products; missing sin rows are zero rows and
  //      their sigma is 0, so the full expression is evaluated for every mode.  A thread owns one mode: its signs and row
  //      indices are set up once and reused for all C channels.
  if ((tid & 127) < g.nq) {  // two threads per mode: channels [0, C/2) and [C/2, C)
    const int q = tid & 127, c0 = (tid >> 7) * (C / 2);
    const int ms = q / g.Lw, kw = q - ms * g.Lw;
    const int kd = kd2[ms >> 1], kh = kh2[ms & 1];
    if (kd >= 0 && kh >= 0) {
      const float gd = (float)kd_desc[4 * kd + 2], gh = (float)kh_desc[4 * kh + 2];
      const int cw = kw_desc[4 * kw], sw = kw_desc[4 * kw + 1];
      const float gw = sw >= 0 ? (float)kw_desc[4 * kw + 2] : 0.f;
      const int swi = sw >= 0 ? sw : cw;
      // rows: 0 = (cos d, cos h), 1 = (cos d, sin h), 2 = (sin d, cos h), 3 = (sin d, sin h)
      const float k0c = scale_in, k0s = scale_in * gw, k1c = scale_in * gh, k1s = -scale_in * gh * gw;
      const float k2c = scale_in * gd, k2s = -scale_in * gd * gw, k3c = -scale_in * gd * gh, k3s = -scale_in * gd * gh * gw;
#pragma unroll 4
      for (int c = c0; c < c0 + C / 2; ++c) {
        const float* t = Tw + c * 4 * g.Jwp;
        float v = k0c * t[cw];
        v = fmaf(k0s, t[swi], v);
        v = fmaf(k1c, t[g.Jwp + cw], v);
        v = fmaf(k1s, t[g.Jwp + swi], v);
        v = fmaf(k2c, t[2 * g.Jwp + cw], v);
        v = fmaf(k2s, t[2 * g.Jwp + swi], v);
        v = fmaf(k3c, t[3 * g.Jwp + cw], v);
        v = fmaf(k3s, t[3 * g.Jwp + swi], v);
        zq[c * g.nqp + q] = v;
      }
    } else {
      for (int c = c0; c < c0 + C / 2; ++c) zq[c * g.nqp + q] = 0.f;
    }
  }
  __syncthreads();
  CORE_STAMP(4)

  // ---- the shared-weight mixes on this CTA's modes.  TWO threads per mode (q = tid & 127, half = tid >> 7): each computes
  //      half of the C output channels from all C inputs, exchanging the layer outputs through zq (its own half only is
  //      written, both halves are read: one barrier per layer).  One thread per mode left 4.5 of 8 warps waiting at the
  //      barrier for 2,100 serial instructions (ncu: barrier stalls 2.75 per issue, profiles/r2w).
  {
    constexpr int CH = C / 2;              // channels per half
    const int q = tid & 127, half = tid >> 7;
    const bool active = q < g.nq;
    const int ms = active ? q / g.Lw : 0, kw = active ? q - ms * g.Lw : 0;
    const int kd = kd2[ms >> 1], kh = kh2[ms & 1];
    const bool valid = active && kd >= 0 && kh >= 0;
    const long cstride = (long)g.Ld * g.Lh * g.Lw;
    const long goff = valid ? (((long)b * C * g.Ld + kd) * g.Lh + kh) * g.Lw + kw + (long)(half * CH) * cstride : 0;
    if (!BWD) {
      if (valid && zall != nullptr) {
#pragma unroll
        for (int c = 0; c < CH; ++c) zall[goff + c * cstride] = zq[(half * CH + c) * g.nqp + q];
      }
      for (int l = 0; l < L; ++l) {
        float2 acc[CH / 2];
        if (valid) {
          float x[C];
#pragma unroll
          for (int c = 0; c < C; ++c) x[c] = zq[c * g.nqp + q];
#pragma unroll
          for (int h = 0; h < CH / 2; ++h) acc[h] = make_float2(x[half * CH + 2 * h], x[half * CH + 2 * h + 1]);  // residual
          const float4* w4 = reinterpret_cast<const float4*>(wts + l * C * C + half * CH);
#pragma unroll
          for (int i = 0; i < C; ++i) {
            const float2 xi = dup2(x[i]);
#pragma unroll
            for (int h = 0; h < CH / 4; ++h) {
              const float4 w = w4[i * (C / 4) + h];
              acc[2 * h] = ffma2(make_float2(w.x, w.y), xi, acc[2 * h]);
              acc[2 * h + 1] = ffma2(make_float2(w.z, w.w), xi, acc[2 * h + 1]);
            }
          }
        }
        __syncthreads();  // every input of this layer has been read
        if (valid) {
          float* op = zall != nullptr ? zall + (long)(l + 1) * BCM + goff : nullptr;
#pragma unroll
          for (int h = 0; h < CH / 2; ++h) {
            const float2 y = selu2(acc[h]);
            zq[(half * CH + 2 * h) * g.nqp + q] = y.x;
            zq[(half * CH + 2 * h + 1) * g.nqp + q] = y.y;
            if (op != nullptr) {
              op[(2 * h) * cstride] = y.x;
              op[(2 * h + 1) * cstride] = y.y;
            }
          }
        }
        __syncthreads();
      }
    } else {
      // backward of the chain (k_modechain_bwd on this CTA's modes): d(pre) = g * selu'(z_{l+1}), dW_l += d(pre) z_l^T,
      // g <- W_l^T d(pre) + d(pre); g lives in zq, a thread handles the CH channels of its half
      const int TVS = g.nqp + 4;
      float* sx = sdp + C * TVS;  // z_l (input of layer l): this tile and the one parked on fwT | Tw take turns
      float* sy = sdp;            // z_{l+1} (its SELU output): first layer in the d(pre) tile itself, then the previous z_l tile
      float* spare = fwT;         // free from here until the recombination^T below
      constexpr int TI = C % 3 == 0 ? 3 : 2;  // weight-gradient tile: one output row, TI input columns per thread
      constexpr int NI = C / TI;
      static_assert(C % TI == 0 && C * NI <= kCoreThreads, "weight-gradient tiling");
      const int wo = tid / NI, wi0 = (tid - wo * NI) * TI;
      const bool wactive = tid < C * NI;
      const int cta = (blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x;
      for (int l = L - 1; l >= 0; --l) {
        if (l == L - 1) { CORE_STAMP(8) }
        asm volatile("cp.async.wait_group 0;" ::: "memory");  // this thread's elements of sy / sx have landed
        if (q < g.nqp) {
#pragma unroll
          for (int c = 0; c < CH; ++c) {
            const int cc = half * CH + c;
            sdp[cc * TVS + q] = valid ? zq[cc * g.nqp + q] * selu_grad_from_out(sy[cc * TVS + q]) : 0.f;
          }
        }
        __syncthreads();
        // z_{l-1} replaces z_{l+1} (every thread has read its own elements of it) while this layer computes
        float* nxt = sy == sdp ? spare : sy;
        if (l > 0) prefetch_layer(l - 1, nxt);
        if (l == L - 1) { CORE_STAMP(9) }
        if (wactive) {
          float acc[TI];
#pragma unroll
          for (int r = 0; r < TI; ++r) acc[r] = 0.f;
          for (int v = 0; v < g.nqp; v += 4) {
            const float4 d = *reinterpret_cast<const float4*>(sdp + wo * TVS + v);
#pragma unroll
            for (int r = 0; r < TI; ++r) {
              const float4 xv = *reinterpret_cast<const float4*>(sx + (wi0 + r) * TVS + v);
              acc[r] = fmaf(d.x, xv.x, fmaf(d.y, xv.y, fmaf(d.z, xv.z, fmaf(d.w, xv.w, acc[r]))));
            }
          }
          float* pr = partials + ((long)cta * L + l) * C * C + wo * C + wi0;
#pragma unroll
          for (int r = 0; r < TI; ++r) pr[r] = acc[r];
        }
        if (l == L - 1) { CORE_STAMP(10) }
        if (valid) {
          float2 a2[CH / 2];
#pragma unroll
          for (int h = 0; h < CH / 2; ++h)
            a2[h] = make_float2(sdp[(half * CH + 2 * h) * TVS + q], sdp[(half * CH + 2 * h + 1) * TVS + q]);
          const float4* w4 = reinterpret_cast<const float4*>(wts + l * C * C + half * CH);
#pragma unroll
          for (int o = 0; o < C; ++o) {
            const float2 d = dup2(sdp[o * TVS + q]);
#pragma unroll
            for (int h = 0; h < CH / 4; ++h) {
              const float4 wv = w4[o * (C / 4) + h];
              a2[2 * h] = ffma2(make_float2(wv.x, wv.y), d, a2[2 * h]);
              a2[2 * h + 1] = ffma2(make_float2(wv.z, wv.w), d, a2[2 * h + 1]);
            }
          }
#pragma unroll
          for (int h = 0; h < CH / 2; ++h) {
            zq[(half * CH + 2 * h) * g.nqp + q] = a2[h].x;
            zq[(half * CH + 2 * h + 1) * g.nqp + q] = a2[h].y;
          }
        }
        if (l == L - 1) { CORE_STAMP(11) }
        __syncthreads();  // the staging tiles are rewritten by the next layer
        if (l == L - 1) { CORE_STAMP(12) }
        sy = sx;  // next layer: y = z_l (this layer's input tile), x = z_{l-1} (landing in the tile freed above)
        sx = nxt;
      }
    }
  }
  __syncthreads();
  CORE_STAMP(5)

  // ---- recombination^T (k_dht_tail_adj): Tt[c][jw][r] = sign * scale * sum_{+-d, +-h, +-w} (+-) z[c][mode], stored with the
  //      4 rows of a channel adjacent so that the synthesis reads them with one LDS.128.  A thread owns one (r, jw): the
  //      (up to 8) source modes and their signs are set up once and reused for all C channels.
  float* Tt = Tw;  // [C][Jwp][4]
  for (int item = tid & 127; item < 4 * g.Jwp; item += 128) {  // two threads per (r, jw): half of the channels each
    const int c0 = (tid >> 7) * (C / 2);
    const int r = item & 3, jw = item >> 2;
    const int isd = r >> 1, ish = r & 1;
    int src[8];
    float coef[8];
    int n = 0;
    if (jw < g.Jw && rd[isd] >= 0 && rh[ish] >= 0) {
      const int isw = jw >= g.JCw ? 1 : 0;
      const float sign = ((isd + ish + isw >= 2) ? -1.f : 1.f) * scale_out;
#pragma unroll
      for (int a = 0; a < 2; ++a)
#pragma unroll
        for (int bb = 0; bb < 2; ++bb)
#pragma unroll
          for (int cc = 0; cc < 2; ++cc) {
            const int kw = jw_desc[4 * jw + cc];
            const float f = ((isd && a == 1) ? -1.f : 1.f) * ((ish && bb == 1) ? -1.f : 1.f) * ((isw && cc == 1) ? -1.f : 1.f);
            const bool ok = kd2[a] >= 0 && kh2[bb] >= 0 && kw >= 0;
            src[n] = ok ? (a * 2 + bb) * g.Lw + kw : 0;
            coef[n] = ok ? sign * f : 0.f;
            ++n;
          }
    } else {
#pragma unroll
      for (int k = 0; k < 8; ++k) src[k] = 0, coef[k] = 0.f;
    }
#pragma unroll 2
    for (int c = c0; c < c0 + C / 2; ++c) {
      const float* zp = zq + c * g.nqp;
      float acc = 0.f;
#pragma unroll
      for (int k = 0; k < 8; ++k) acc = fmaf(coef[k], zp[src[k]], acc);
      Tt[(c * g.Jwp + jw) * 4 + r] = acc;
    }
  }
  __syncthreads();
  CORE_STAMP(6)

  // ---- W synthesis: T2[c][r][w] = sum_jw Tt[c][jw][r] fwP[jw][w]; a thread owns the 4 rows of a channel x 4 columns w
  {
    const int q4 = g.Wp >> 2;
    for (int item = tid; item < C * q4; item += kCoreThreads) {
      const int c = item / q4, c4 = item - c * q4;
      const float* tp = Tt + c * g.Jwp * 4;
      const float* fp = fwP + 4 * c4;
      float2 acc[4][2];
#pragma unroll
      for (int r = 0; r < 4; ++r) acc[r][0] = acc[r][1] = make_float2(0.f, 0.f);
#pragma unroll 2
      for (int jw = 0; jw < g.Jw; ++jw) {
        const float4 f = *reinterpret_cast<const float4*>(fp + jw * g.Wp);
        const float4 t = *reinterpret_cast<const float4*>(tp + jw * 4);
        const float2 f01 = make_float2(f.x, f.y), f23 = make_float2(f.z, f.w);
        acc[0][0] = ffma2(dup2(t.x), f01, acc[0][0]);
        acc[0][1] = ffma2(dup2(t.x), f23, acc[0][1]);
        acc[1][0] = ffma2(dup2(t.y), f01, acc[1][0]);
        acc[1][1] = ffma2(dup2(t.y), f23, acc[1][1]);
        acc[2][0] = ffma2(dup2(t.z), f01, acc[2][0]);
        acc[2][1] = ffma2(dup2(t.z), f23, acc[2][1]);
        acc[3][0] = ffma2(dup2(t.w), f01, acc[3][0]);
        acc[3][1] = ffma2(dup2(t.w), f23, acc[3][1]);
      }
#pragma unroll
      for (int r = 0; r < 4; ++r) {
        const int jd = rd[r >> 1], jh = rh[r & 1];
        if (jd < 0 || jh < 0) continue;
        *reinterpret_cast<float4*>(T2 + (long)(b * C + c) * g.slablen + (long)jh * g.rowlen + (long)jd * g.Wp + 4 * c4) =
            make_float4(acc[r][0].x, acc[r][0].y, acc[r][1].x, acc[r][1].y);
      }
    }
  }
  __syncthreads();
  CORE_STAMP(7)
}

// ------------------------------------------------------------------------------------------------ host side
static void core_prof_print(const char* what, cudaStream_t st) {
  static const bool on = getenv("HNO_CORE_PROF") && atoi(getenv("HNO_CORE_PROF")) != 0;
  if (!on) return;
  cudaStreamSynchronize(st);
  long long h[16];
  cudaMemcpyFromSymbol(h, g_core_prof, sizeof(h));
  fprintf(stderr, "[core_prof %s] prologue %lld | pdl wait %lld | analysis %lld | recombine %lld | chain %lld | recombT %lld | synthesis %lld | total %lld\n",
          what, h[1] - h[0], h[2] - h[1], h[3] - h[2], h[4] - h[3], h[5] - h[4], h[6] - h[5], h[7] - h[6], h[7] - h[0]);
  fprintf(stderr, "   chain layer L-1: stage+sync %lld | wgrad (thread 0) %lld | W^T product %lld | final sync %lld\n", h[9] - h[8], h[10] - h[9], h[11] - h[10], h[12] - h[11]);
}

static bool make_core_geom(const DhtPlanHeader* h, CoreGeom* out) {
  CoreGeom g;
  g.W = h->ax[2].n;
  g.Wp = (g.W + 3) & ~3;
  g.Jw = h->ax[2].J;
  g.Jwp = (g.Jw + 3) & ~3;
  g.JCw = h->ax[2].JC;
  g.Ld = h->ax[0].L;
  g.Lh = h->ax[1].L;
  g.Lw = h->ax[2].L;
  g.rowlen = (long)h->ax[0].J * g.Wp;
  g.slablen = (long)h->ax[1].J * g.rowlen;
  g.off_fullT_w = h->ax[2].off_fullT;
  g.off_fullP_w = h->ax[2].off_fullP;
  for (int a = 0; a < 3; ++a) {
    g.off_kdesc[a] = h->ax[a].off_kdesc;
    g.off_jdesc[a] = h->ax[a].off_jdesc;
  }
  g.nq = 4 * g.Lw;
  g.nqp = (g.nq + 3) & ~3;
  *out = g;
  if (g.Jwp != 32 && g.Jwp > 32) return false;             // lane = jw in the W analysis
  if (g.nq > kCoreThreads / 2) return false;               // two threads per local mode
  if (h->ax[0].JC > 32 || h->ax[1].JC > 32) return false;
  // fullT ([n4][J4]) must be followed directly by fullP ([J][n4]) in the blob (one copy loop)
  if (g.off_fullP_w != g.off_fullT_w + g.Wp * g.Jwp) return false;
  const int* iw = reinterpret_cast<const int*>(h);
  for (int a = 0; a < 2; ++a) {
    const DhtAxis& ax = h->ax[a];
    int* dst = a == 0 ? g.sin_d : g.sin_h;
    for (int j = 0; j < 32; ++j) dst[j] = -1;
    for (int jc = 0; jc < ax.JC; ++jc) {
      const int u = iw[ax.off_jdesc + 4 * jc + 3];
      for (int j = ax.JC; j < ax.J; ++j)
        if (iw[ax.off_jdesc + 4 * j + 3] == u) dst[jc] = j;
    }
  }
  *out = g;
  return true;
}

static size_t core_smem(const CoreGeom& g, int C, int L, bool bwd) {
  size_t f = (size_t)g.Wp * g.Jwp + (size_t)g.Jw * g.Wp + (size_t)C * 4 * g.Jwp + (size_t)C * g.nqp + (size_t)L * C * C;
  if (bwd) f += (size_t)2 * C * (g.nqp + 4);  // d(pre) and z_l staging tiles (the third is parked on fwT | Tw)
  return f * sizeof(float);
}

bool spectral_core_eligible(const void* plan_host, int C, int L, int B) {
  const auto* h = reinterpret_cast<const DhtPlanHeader*>(plan_host);
  CoreGeom g;
  if (!make_core_geom(h, &g)) return false;
  if (C != 24 && C != 8) return false;
  if (L < 1 || L > kCoreMaxLayers || B < 1 || B > 65535) return false;
  if ((size_t)g.Wp * g.Jwp + (size_t)C * 4 * g.Jwp < (size_t)C * (g.nqp + 4)) return false;  // the parked staging tile
  return core_smem(g, C, L, true) <= 75 * 1024 && g.nq <= 128;
}

size_t spectral_core_partials_floats(const void* plan_host, int C, int L, int B) {
  const auto* h = reinterpret_cast<const DhtPlanHeader*>(plan_host);
  return (size_t)h->ax[1].JC * h->ax[0].JC * B * L * C * C;
}

int reduce_chain_partials(const float* partials, int nrows, int L, int n, float* const* dweights, int accumulate,
                          cudaStream_t st);

// T2 [B*C][Jh][Jd][Wp] in place.  Forward: zall (may be null) receives z_0..z_L.  Backward: zall is read, dweights written.
int spectral_core(const void* plan_host, const void* plan_dev, float* T2, float* zall, const float* const* weights,
                  float* const* dweights, float* partials, int B, int C, int L, float scale_in, float scale_out,
                  bool backward, int accumulate_dw, cudaStream_t st) {
  const auto* h = reinterpret_cast<const DhtPlanHeader*>(plan_host);
  CoreGeom g;
  HNO_CHECK(make_core_geom(h, &g) && spectral_core_eligible(plan_host, C, L, B), "spectral_core: configuration not eligible");
  HNO_CHECK(T2 && weights && (!backward || (zall && dweights && partials)), "spectral_core: null pointer");
  CorePtrs P{};
  for (int l = 0; l < L; ++l) {
    HNO_CHECK(weights[l] != nullptr, "spectral_core: null weight pointer");
    P.w[l] = weights[l];
  }
  const size_t smem = core_smem(g, C, L, backward);
  static const bool core_pdl = !(getenv("HNO_CORE_PDL") && atoi(getenv("HNO_CORE_PDL")) == 0);
  dim3 grid(h->ax[1].JC, h->ax[0].JC, B);
  const float* pf = reinterpret_cast<const float*>(plan_dev);
  const int* pi = reinterpret_cast<const int*>(plan_dev);
#define HNO_CORE_LAUNCH(CC, BW)                                                                                     \
  {                                                                                                                 \
    auto kern = k_spectral_core<CC, BW>;                                                                            \
    HNO_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));                   \
    HNO_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared)); \
    cudaLaunchConfig_t cfg = {};                                                                                    \
    cfg.gridDim = grid;                                                                                             \
    cfg.blockDim = dim3(kCoreThreads);                                                                              \
    cfg.dynamicSmemBytes = smem;                                                                                    \
    cfg.stream = st;                                                                                                \
    cudaLaunchAttribute attr[1];                                                                                    \
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;                                                \
    attr[0].val.programmaticStreamSerializationAllowed = 1;                                                         \
    cfg.attrs = attr;                                                                                               \
    cfg.numAttrs = core_pdl ? 1 : 0;                                                                                \
    HNO_CUDA(cudaLaunchKernelEx(&cfg, kern, T2, zall, partials, pf, pi, g, P, L, B, scale_in, scale_out));          \
    if (getenv("HNO_CORE_PROF")) {                                                                                  \
      int nb = 0;                                                                                                   \
      cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, kern, kCoreThreads, smem);                                 \
      fprintf(stderr, "[core_prof] smem %zu B, blocks/SM %d, grid %u\n", smem, nb, grid.x * grid.y * grid.z);       \
    }                                                                                                               \
  }
  if (C == 24) {
    if (backward) HNO_CORE_LAUNCH(24, true) else HNO_CORE_LAUNCH(24, false)
  } else {
    if (backward) HNO_CORE_LAUNCH(8, true) else HNO_CORE_LAUNCH(8, false)
  }
#undef HNO_CORE_LAUNCH
  HNO_LAUNCH_CHECK();
  core_prof_print(backward ? "bwd" : "fwd", st);
  if (backward) return reduce_chain_partials(partials, (int)(grid.x * grid.y * grid.z), L, C * C, dweights, accumulate_dw, st);
  return 0;
}

}  // namespace hno
