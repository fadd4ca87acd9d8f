// Fused middle stages of the truncated 3-D DHT (sm_100a): everything between the two HBM-bound D-axis contractions.
//
//   forward :  G1[slab][jd][h][w]  --H analysis-->  --W analysis-->  T[jd][jh][jw]  --8-term cas recombination-->  Z
//   adjoint :  Z  --recombination^T-->  T  --W synthesis-->  --H synthesis-->  G1[slab][jd][h][w]
//
// (math: dht_plan.h; replaces the middle of nets/hnosegxs.py:378-410 / :454-494 evaluated as separable contractions.)
// The three separate kernels of dht_kernels.cu (k_analysis_outer + k_analysis_inner + k_combine and their
// transposes) each ran ~1 wave of latency-bound threads over 38 MB / 9 MB / 3 MB of L2-resident data and together cost
// as much as the HBM-bound stage (ncu: 33 + 32 + 12 us vs 75 us).  Here ONE CTA owns one (slab, |u_d|) pair, i.e. the
// cos and the sin plane of the same D frequency: it stages a 121 x 78 plane in shared memory, runs both contractions
// out of shared memory with register-tiled broadcast loads, keeps T for the two planes on chip and finishes with the
// recombination for the (at most two) output frequencies kd = +u, -u that need exactly these two planes.
// All arithmetic is exact fp32 FFMA with the fp64-generated tables of the plan.
#include "common.cuh"
#include "dht_plan.h"

namespace hno {

constexpr int kMidThreads = 256;

struct MidGeom {
  int H, W, Jd, JCd, Jh, Jw, Ld, Lh, Lw;
  int Jhp, Jwp;      // Jh, Jw rounded up to 4
  int Hp, Wp;        // H, W rounded up to 4
  long P;            // plane pitch of G1 (floats)
  int off_full_h, off_full_w;
  int off_kdesc[3], off_jdesc[3];
  // even/odd folded tables of the H axis: fcos [(nh+1)][JCp], fsin [(nh+1)][JSp]
  int JCh, JSh, JChp, JShp, nhh, off_fcos_h, off_fsin_h;
  int NHp;           // nh + 1 rounded up to an even number (pitch of the transposed tables of the adjoint)
};

static inline int r4(int v) { return (v + 3) & ~3; }

// shared-memory layout (floats); forward and adjoint use the same regions
struct MidSmem {
  int plane, fh, fw, t2, T, total;
};
static MidSmem mid_smem_fwd(const MidGeom& g) {
  MidSmem s;
  s.plane = 0;
  int cur = g.H * g.Wp;               // plane rows padded to Wp
  s.fh = cur; cur += (g.nhh + 1) * (g.JChp + g.JShp);  // folded cos table [i][JChp], then folded sin table [i][JShp]
  s.fw = cur; cur += g.W * g.Jwp;     // [w][Jwp]
  s.t2 = cur; cur += g.Wp * (g.JChp + g.JShp);  // transposed: [w][JChp + JShp]
  s.T = cur; cur += r4(2 * g.Jh * g.Jw);
  s.total = cur;
  return s;
}
static MidSmem mid_smem_adj(const MidGeom& g) {
  MidSmem s;
  s.plane = 0;
  int cur = 0;
  s.fh = cur; cur += r4(g.Jh * g.NHp);  // transposed folded tables: cos rows [j][NHp], then sin rows [j][NHp]
  s.fw = cur; cur += g.Jw * g.Wp;     // [jw][Wp]
  s.t2 = cur; cur += g.Jh * g.Wp;     // [jh][Wp]
  s.T = cur; cur += r4(2 * g.Jh * g.Jw);
  s.total = cur;
  return s;
}

// sin row of the D axis that shares |u| with cos row jc (or -1)
__device__ __forceinline__ int find_sin_row(const int* jdesc, int Jd, int JCd, int jc) {
  const int u = jdesc[4 * jc + 3];
  for (int j = JCd; j < Jd; ++j)
    if (jdesc[4 * j + 3] == u) return j;
  return -1;
}

// ------------------------------------------------------------------------------------------------ forward
__global__ void __launch_bounds__(kMidThreads, 3) k_dht_mid_fwd(const float* __restrict__ G1, float* __restrict__ Z,
                                                               const float* __restrict__ pf,
                                                               const int* __restrict__ pi, const MidGeom g,
                                                               const MidSmem sm, float scale) {
  extern __shared__ float4 smem4[];
  float* s = reinterpret_cast<float*>(smem4);
  float* plane = s + sm.plane;
  float* fh = s + sm.fh;
  float* fw = s + sm.fw;
  float* t2 = s + sm.t2;
  float* T = s + sm.T;
  const int tid = threadIdx.x;
  const int jc = blockIdx.x;
  const long slab = blockIdx.y;
  const int* jd_desc = pi + g.off_jdesc[0];
  const int js = find_sin_row(jd_desc, g.Jd, g.JCd, jc);
  const int HW = g.H * g.W;

  // tables: the folded H tables are stored [i][row] in the plan, i.e. one LDS.128 yields four output rows
  float* fcs = fh;
  float* fsn = fh + (g.nhh + 1) * g.JChp;
  for (int idx = tid; idx < (g.nhh + 1) * g.JChp; idx += kMidThreads) fcs[idx] = __ldg(pf + g.off_fcos_h + idx);
  for (int idx = tid; idx < (g.nhh + 1) * g.JShp; idx += kMidThreads) fsn[idx] = __ldg(pf + g.off_fsin_h + idx);
  for (int j = tid >> 5; j < g.Jwp; j += kMidThreads / 32)  // fw[w][j] <- full[j][w]: a warp per table row
    for (int w = tid & 31; w < g.W; w += 32) fw[w * g.Jwp + j] = j < g.Jw ? __ldg(pf + g.off_full_w + j * g.W + w) : 0.f;

  for (int pass = 0; pass < 2; ++pass) {
    const int row = pass == 0 ? jc : js;
    float* Tp = T + pass * g.Jh * g.Jw;
    if (row < 0) break;
    __syncthreads();  // previous pass is done with `plane` and `t2`
    {
      const float* src = G1 + (slab * g.Jd + row) * g.P;
      const int wrp = tid >> 5, ln = tid & 31;
      if ((g.W & 1) == 0) {  // 8-byte pieces: rows of W floats -> rows of Wp floats (a warp per row, no divisions)
        const int hw2 = g.W >> 1;
        for (int h = wrp; h < g.H; h += kMidThreads / 32) {
          const float* sr = src + h * g.W;
          const unsigned dr = (unsigned)__cvta_generic_to_shared(plane + h * g.Wp);
          for (int c = ln; c < hw2; c += 32)
            asm volatile("cp.async.ca.shared.global [%0], [%1], 8;\n" ::"r"(dr + 8 * c), "l"(sr + 2 * c) : "memory");
        }
      } else {
        for (int h = wrp; h < g.H; h += kMidThreads / 32) {
          const float* sr = src + h * g.W;
          const unsigned dr = (unsigned)__cvta_generic_to_shared(plane + h * g.Wp);
          for (int c = ln; c < g.W; c += 32)
            asm volatile("cp.async.ca.shared.global [%0], [%1], 4;\n" ::"r"(dr + 4 * c), "l"(sr + c) : "memory");
        }
      }
      if (g.Wp != g.W && ln < g.Wp - g.W)  // zero the row padding (read by the 4-wide tiles)
        for (int h = wrp; h < g.H; h += kMidThreads / 32) plane[h * g.Wp + g.W + ln] = 0.f;
      cp_async_wait_all();
    }
    __syncthreads();
    // ---- H analysis with even/odd folding (cos rows see x[i] + x[n-i], sin rows x[i] - x[n-i]):
    //      tile = 4 rows x 4 columns, one step = 3 LDS.128 + 4 adds + 8 FFMA2 for 32 multiply-adds of the plain form
    {
      const int gc = g.JChp >> 2, gs = g.JSh > 0 ? (g.JShp >> 2) : 0;
      const int wq = g.Wp >> 2;
      const int ntile = (gc + gs) * wq;
      const int n = g.H, npair = (n - 1) >> 1;
      for (int tile = tid; tile < ntile; tile += kMidThreads) {
        // row groups run fastest: the eight threads of a store phase then write 128 contiguous bytes of one t2 column
        const int wg = tile / (gc + gs), jg = tile - wg * (gc + gs);
        const bool is_sin = jg >= gc;
        const int jq = is_sin ? jg - gc : jg;
        const int pitch = is_sin ? g.JShp : g.JChp;
        const float* tab = (is_sin ? fsn : fcs) + 4 * jq;
        const float* pp = plane + 4 * wg;
        float2 acc[4][2];  // [row][column pair]
#pragma unroll
        for (int r = 0; r < 4; ++r) acc[r][0] = acc[r][1] = make_float2(0.f, 0.f);
        auto step = [&](const float4 x, const float4 f) {
          const float2 x01 = make_float2(x.x, x.y), x23 = make_float2(x.z, x.w);
          acc[0][0] = ffma2(dup2(f.x), x01, acc[0][0]);
          acc[0][1] = ffma2(dup2(f.x), x23, acc[0][1]);
          acc[1][0] = ffma2(dup2(f.y), x01, acc[1][0]);
          acc[1][1] = ffma2(dup2(f.y), x23, acc[1][1]);
          acc[2][0] = ffma2(dup2(f.z), x01, acc[2][0]);
          acc[2][1] = ffma2(dup2(f.z), x23, acc[2][1]);
          acc[3][0] = ffma2(dup2(f.w), x01, acc[3][0]);
          acc[3][1] = ffma2(dup2(f.w), x23, acc[3][1]);
        };
        if (!is_sin) step(*reinterpret_cast<const float4*>(pp), *reinterpret_cast<const float4*>(tab));
#pragma unroll 2
        for (int i = 1; i <= npair; ++i) {
          const float4 a = *reinterpret_cast<const float4*>(pp + i * g.Wp);
          const float4 c = *reinterpret_cast<const float4*>(pp + (n - i) * g.Wp);
          const float4 f = *reinterpret_cast<const float4*>(tab + i * pitch);
          const float4 x = is_sin ? make_float4(a.x - c.x, a.y - c.y, a.z - c.z, a.w - c.w)
                                  : make_float4(a.x + c.x, a.y + c.y, a.z + c.z, a.w + c.w);
          step(x, f);
        }
        if (!is_sin && (n & 1) == 0 && n > 1)  // Nyquist sample pairs with itself, its sine vanishes
          step(*reinterpret_cast<const float4*>(pp + (n >> 1) * g.Wp),
               *reinterpret_cast<const float4*>(tab + (n >> 1) * pitch));
        // t2 is kept transposed, [w][JT] with the cos rows at [0, JChp) and the sin rows at [JChp, JChp + JShp):
        // one STS.128 per column here, one LDS.128 per (w, 4 rows) in the W analysis
        const int JT = g.JChp + g.JShp;
        float* o = t2 + (4 * wg) * JT + (is_sin ? g.JChp : 0) + 4 * jq;
        if (4 * wg + 0 < g.W) *reinterpret_cast<float4*>(o) = make_float4(acc[0][0].x, acc[1][0].x, acc[2][0].x, acc[3][0].x);
        if (4 * wg + 1 < g.W) *reinterpret_cast<float4*>(o + JT) = make_float4(acc[0][0].y, acc[1][0].y, acc[2][0].y, acc[3][0].y);
        if (4 * wg + 2 < g.W) *reinterpret_cast<float4*>(o + 2 * JT) = make_float4(acc[0][1].x, acc[1][1].x, acc[2][1].x, acc[3][1].x);
        if (4 * wg + 3 < g.W) *reinterpret_cast<float4*>(o + 3 * JT) = make_float4(acc[0][1].y, acc[1][1].y, acc[2][1].y, acc[3][1].y);
      }
    }
    __syncthreads();
    // ---- W analysis: T[jh][jw] = sum_w t2[w][jh] fw[w][jw]; tile = 4 rows jh x 4 columns jw (2 LDS.128 + 8 FFMA2 a step)
    {
      const int JT = g.JChp + g.JShp;
      const int gh = JT >> 2, gw = g.Jwp >> 2;
      const int ntile = gh * gw;
      for (int tile = tid; tile < ntile; tile += kMidThreads) {
        const int hq = tile / gw, wq4 = tile - hq * gw;
        const float* tp = t2 + 4 * hq;
        const float* fp = fw + 4 * wq4;
        float2 acc[4][2];
#pragma unroll
        for (int r = 0; r < 4; ++r) acc[r][0] = acc[r][1] = make_float2(0.f, 0.f);
#pragma unroll 2
        for (int w = 0; w < g.W; ++w) {
          const float4 x = *reinterpret_cast<const float4*>(tp + w * JT);     // 4 rows jh
          const float4 f = *reinterpret_cast<const float4*>(fp + w * g.Jwp);  // 4 columns jw
          const float2 f01 = make_float2(f.x, f.y), f23 = make_float2(f.z, f.w);
          acc[0][0] = ffma2(dup2(x.x), f01, acc[0][0]);
          acc[0][1] = ffma2(dup2(x.x), f23, acc[0][1]);
          acc[1][0] = ffma2(dup2(x.y), f01, acc[1][0]);
          acc[1][1] = ffma2(dup2(x.y), f23, acc[1][1]);
          acc[2][0] = ffma2(dup2(x.z), f01, acc[2][0]);
          acc[2][1] = ffma2(dup2(x.z), f23, acc[2][1]);
          acc[3][0] = ffma2(dup2(x.w), f01, acc[3][0]);
          acc[3][1] = ffma2(dup2(x.w), f23, acc[3][1]);
        }
        // rows of the padded numbering -> rows of T (cos rows [0, JCh), sin rows [JCh, Jh))
#pragma unroll
        for (int r = 0; r < 4; ++r) {
          const int jp = 4 * hq + r;
          int jh = -1;
          if (jp < g.JChp) {
            if (jp < g.JCh) jh = jp;
          } else if (jp - g.JChp < g.JSh) {
            jh = g.JCh + jp - g.JChp;
          }
          if (jh >= 0) {
            const float v[4] = {acc[r][0].x, acc[r][0].y, acc[r][1].x, acc[r][1].y};
#pragma unroll
            for (int c = 0; c < 4; ++c)
              if (4 * wq4 + c < g.Jw) Tp[jh * g.Jw + 4 * wq4 + c] = v[c];
          }
        }
      }
    }
  }
  __syncthreads();
  // ---- 8-term recombination for kd = +u and -u (same expression as k_combine, T from shared memory)
  const int* kd_desc = pi + g.off_kdesc[0];
  const int* kh_desc = pi + g.off_kdesc[1];
  const int* kw_desc = pi + g.off_kdesc[2];
  const float* Tc = T;
  const float* Ts = T + g.Jh * g.Jw;
  for (int a = 0; a < 2; ++a) {
    const int kd = jd_desc[4 * jc + a];
    if (kd < 0) continue;
    const int sd = kd_desc[4 * kd + 1];
    const float gd = (float)kd_desc[4 * kd + 2];
    float* zo = Z + ((slab * g.Ld + kd) * g.Lh) * (long)g.Lw;
    const int n = g.Lh * g.Lw;
    for (int o = tid; o < n; o += kMidThreads) {
      const int kh = o / g.Lw, kw = o - kh * g.Lw;
      const int ch = kh_desc[4 * kh], sh = kh_desc[4 * kh + 1];
      const float gh = (float)kh_desc[4 * kh + 2];
      const int cw = kw_desc[4 * kw], sw = kw_desc[4 * kw + 1];
      const float gw = (float)kw_desc[4 * kw + 2];
      float v = Tc[ch * g.Jw + cw];
      if (sh >= 0 && sw >= 0) v -= gh * gw * Tc[sh * g.Jw + sw];
      if (sd >= 0 && sw >= 0) v -= gd * gw * Ts[ch * g.Jw + sw];
      if (sd >= 0 && sh >= 0) v -= gd * gh * Ts[sh * g.Jw + cw];
      if (sd >= 0) v += gd * Ts[ch * g.Jw + cw];
      if (sh >= 0) v += gh * Tc[sh * g.Jw + cw];
      if (sw >= 0) v += gw * Tc[ch * g.Jw + sw];
      if (sd >= 0 && sh >= 0 && sw >= 0) v -= gd * gh * gw * Ts[sh * g.Jw + sw];
      zo[o] = scale * v;
    }
  }
}

// ------------------------------------------------------------------------------------------------ adjoint
__global__ void __launch_bounds__(kMidThreads, 2) k_dht_mid_adj(const float* __restrict__ Z, float* __restrict__ G1,
                                                               const float* __restrict__ pf,
                                                               const int* __restrict__ pi, const MidGeom g,
                                                               const MidSmem sm, float scale) {
  extern __shared__ float4 smem4[];
  float* s = reinterpret_cast<float*>(smem4);
  float* fh = s + sm.fh;   // [jh][Hp]
  float* fw = s + sm.fw;   // [jw][Wp]
  float* g2 = s + sm.t2;   // [jh][Wp]
  float* T = s + sm.T;
  const int tid = threadIdx.x;
  const int jc = blockIdx.x;
  const long slab = blockIdx.y;
  const int* jd_desc = pi + g.off_jdesc[0];
  const int* jh_desc = pi + g.off_jdesc[1];
  const int* jw_desc = pi + g.off_jdesc[2];
  const int js = find_sin_row(jd_desc, g.Jd, g.JCd, jc);

  for (int idx = tid; idx < g.Jh * g.NHp; idx += kMidThreads) {  // fh[j][i]: cos rows then sin rows, i = 0..nh
    const int j = idx / g.NHp, i = idx - j * g.NHp;
    float v = 0.f;
    if (i <= g.nhh)
      v = j < g.JCh ? __ldg(pf + g.off_fcos_h + i * g.JChp + j) : __ldg(pf + g.off_fsin_h + i * g.JShp + (j - g.JCh));
    fh[idx] = v;
  }
  for (int idx = tid; idx < g.Jw * g.Wp; idx += kMidThreads) {
    const int j = idx / g.Wp, w = idx - j * g.Wp;
    fw[idx] = w < g.W ? __ldg(pf + g.off_full_w + (long)j * g.W + w) : 0.f;
  }
  // ---- recombination^T (same expression as k_combine_t) for the rows jc and js
  const float* Zs = Z + slab * (long)g.Ld * g.Lh * g.Lw;
  for (int pass = 0; pass < 2; ++pass) {
    const int jd = pass == 0 ? jc : js;
    if (jd < 0) break;
    const int isd = jd_desc[4 * jd + 2];
    const int n = g.Jh * g.Jw;
    for (int o = tid; o < n; o += kMidThreads) {
      const int jh = o / g.Jw, jw = o - jh * g.Jw;
      const int ish = jh_desc[4 * jh + 2], isw = jw_desc[4 * jw + 2];
      const float sign = (isd + ish + isw >= 2) ? -1.f : 1.f;
      float acc = 0.f;
#pragma unroll
      for (int a = 0; a < 2; ++a) {
        const int kd = jd_desc[4 * jd + a];
        if (kd < 0) continue;
        const float fd = (isd && a == 1) ? -1.f : 1.f;
#pragma unroll
        for (int b = 0; b < 2; ++b) {
          const int kh = jh_desc[4 * jh + b];
          if (kh < 0) continue;
          const float fhs = (ish && b == 1) ? -fd : fd;
#pragma unroll
          for (int c = 0; c < 2; ++c) {
            const int kw = jw_desc[4 * jw + c];
            if (kw < 0) continue;
            const float fws = (isw && c == 1) ? -fhs : fhs;
            acc = fmaf(fws, __ldg(Zs + ((long)kd * g.Lh + kh) * g.Lw + kw), acc);
          }
        }
      }
      T[pass * n + o] = sign * scale * acc;
    }
  }
  for (int pass = 0; pass < 2; ++pass) {
    const int row = pass == 0 ? jc : js;
    if (row < 0) break;
    const float* Tp = T + pass * g.Jh * g.Jw;
    __syncthreads();  // T ready / previous pass done with g2
    // ---- W synthesis: g2[jh][w] = sum_jw fw[jw][w] T[jh][jw]; tile = 1 row jh x 4 columns w
    {
      const int q = g.Wp >> 2;
      const int ntile = g.Jh * q;
      for (int tile = tid; tile < ntile; tile += kMidThreads) {
        const int jh = tile / q, wg = tile - jh * q;
        const float* tp = Tp + jh * g.Jw;
        const float4* f4 = reinterpret_cast<const float4*>(fw + 4 * wg);
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 4
        for (int jw = 0; jw < g.Jw; ++jw) {
          const float x = tp[jw];
          const float4 f = f4[jw * q];
          acc.x = fmaf(f.x, x, acc.x);
          acc.y = fmaf(f.y, x, acc.y);
          acc.z = fmaf(f.z, x, acc.z);
          acc.w = fmaf(f.w, x, acc.w);
        }
        *reinterpret_cast<float4*>(g2 + jh * g.Wp + 4 * wg) = acc;
      }
    }
    __syncthreads();
    // ---- H synthesis with even/odd folding: e = sum_cos fc[j][i] g2[j][w], o = sum_sin fs[j][i] g2[j][w],
    //      out[i] = e + o, out[n-i] = e - o; tile = 2 values of i (4 output rows) x 4 columns
    {
      float* dst = G1 + (slab * g.Jd + row) * g.P;
      const int n = g.H;
      const int wq = g.Wp >> 2;
      const int niq = (g.nhh + 2) >> 1;  // pairs of i covering 0..nh
      const int ntile = niq * wq;
      for (int tile = tid; tile < ntile; tile += kMidThreads) {
        const int ig = tile / wq, wg = tile - ig * wq;
        const int i0 = 2 * ig;
        const float* gp = g2 + 4 * wg;
        float2 e[2][2], o[2][2];  // [i][column pair]
#pragma unroll
        for (int a = 0; a < 2; ++a) e[a][0] = e[a][1] = o[a][0] = o[a][1] = make_float2(0.f, 0.f);
#pragma unroll 3
        for (int j = 0; j < g.JCh; ++j) {
          const float4 x = *reinterpret_cast<const float4*>(gp + j * g.Wp);
          const float2 f = *reinterpret_cast<const float2*>(fh + j * g.NHp + i0);
          const float2 x01 = make_float2(x.x, x.y), x23 = make_float2(x.z, x.w);
          e[0][0] = ffma2(dup2(f.x), x01, e[0][0]);
          e[0][1] = ffma2(dup2(f.x), x23, e[0][1]);
          e[1][0] = ffma2(dup2(f.y), x01, e[1][0]);
          e[1][1] = ffma2(dup2(f.y), x23, e[1][1]);
        }
#pragma unroll 2
        for (int j = g.JCh; j < g.Jh; ++j) {
          const float4 x = *reinterpret_cast<const float4*>(gp + j * g.Wp);
          const float2 f = *reinterpret_cast<const float2*>(fh + j * g.NHp + i0);
          const float2 x01 = make_float2(x.x, x.y), x23 = make_float2(x.z, x.w);
          o[0][0] = ffma2(dup2(f.x), x01, o[0][0]);
          o[0][1] = ffma2(dup2(f.x), x23, o[0][1]);
          o[1][0] = ffma2(dup2(f.y), x01, o[1][0]);
          o[1][1] = ffma2(dup2(f.y), x23, o[1][1]);
        }
#pragma unroll
        for (int a = 0; a < 2; ++a) {
          const int i = i0 + a;
          if (i > g.nhh) continue;
          const float lo[4] = {e[a][0].x + o[a][0].x, e[a][0].y + o[a][0].y, e[a][1].x + o[a][1].x, e[a][1].y + o[a][1].y};
          const float hi[4] = {e[a][0].x - o[a][0].x, e[a][0].y - o[a][0].y, e[a][1].x - o[a][1].x, e[a][1].y - o[a][1].y};
          const bool mirror = i != 0 && i != n - i;
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            if (4 * wg + c < g.W) {
              dst[i * g.W + 4 * wg + c] = lo[c];
              if (mirror) dst[(n - i) * g.W + 4 * wg + c] = hi[c];
            }
          }
        }
      }
    }
  }
}

// block-wide variant of find_sin_row: one round trip instead of a chain of up to JS dependent global loads at the very
// start of every CTA (every thread must call it)
__device__ __forceinline__ int find_sin_row_block(const int* jdesc, int Jd, int JCd, int jc) {
  __shared__ int s_js;
  if (threadIdx.x == 0) s_js = -1;
  __syncthreads();
  const int u = jdesc[4 * jc + 3];
  const int j = JCd + (int)threadIdx.x;
  if (j < Jd && jdesc[4 * j + 3] == u) s_js = j;
  __syncthreads();
  return s_js;
}

// ================================================================================================ tails
// Tensor-core H stage variant (dht_kernels.cu: dht_hsplit_*): the D stage keeps its result as G1h[slab][h][jd][Wp] and
// the SAME streamed tcgen05 kernel contracts H with (jd, w) as the contiguous axis, T2[slab][jh][jd][Wp].  What is left
// for CUDA cores is 2 % of the arithmetic: the W contraction of 29 short rows per plane and the cas recombination.
// One CTA owns one (slab, |u_d|) pair as above.

struct TailGeom {
  int W, Wp, Jd, JCd, Jh, Jhp, Jw, Jwp, Ld, Lh, Lw;
  long rowlen;       // Jd * Wp: distance between jh rows of T2
  long slablen;      // Jh * rowlen
  int off_full_w, off_fullT_w, off_fullP_w;
  int off_kdesc[3], off_jdesc[3];
};

constexpr int kTailThreads = 256;

// forward: T2 -> W analysis -> recombination -> Z
__global__ void __launch_bounds__(kTailThreads) k_dht_tail_fwd(const float* __restrict__ T2, float* __restrict__ Z,
                                                              const float* __restrict__ pf,
                                                              const int* __restrict__ pi, const TailGeom g,
                                                              float scale) {
  extern __shared__ float4 smem4[];
  float* s = reinterpret_cast<float*>(smem4);
  float* t2s = s;                              // [2][Jhp][Wp]
  float* fw = t2s + 2 * g.Jhp * g.Wp;          // [Wp][Jwp]   (rows >= W are zero)
  float* T = fw + g.Wp * g.Jwp;                // [2][Jh * Jw]
  const int tid = threadIdx.x;
  const int jc = blockIdx.x;
  const long slab = blockIdx.y;
  const int* jd_desc = pi + g.off_jdesc[0];
  const int js = find_sin_row_block(jd_desc, g.Jd, g.JCd, jc);
  const int npass = js >= 0 ? 2 : 1;
  // ---- the 2 x Jh rows of this CTA (16-byte asynchronous copies; pad columns and pad rows are zeroed afterwards)
  {
    const int q = g.Wp >> 2;
    const int wrp = tid >> 5, ln = tid & 31;
    for (int r = wrp; r < npass * g.Jh; r += kTailThreads / 32) {  // a warp per row: no index divisions
      const int pass = r >= g.Jh ? 1 : 0, jh = r - pass * g.Jh;
      const int row = pass == 0 ? jc : js;
      const float* src = T2 + slab * g.slablen + (long)jh * g.rowlen + (long)row * g.Wp;
      const unsigned dst = (unsigned)__cvta_generic_to_shared(t2s + (pass * g.Jhp + jh) * g.Wp);
      for (int c = ln; c < q; c += 32)
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(dst + 16 * c), "l"(src + 4 * c) : "memory");
    }
    // fw[w][j]: the plan's transposed, zero-padded copy of the W rows ([Wp][Jwp] floats, 16-byte pieces)
    const float* fsrc = pf + g.off_fullT_w;
    const unsigned fdst = (unsigned)__cvta_generic_to_shared(fw);
    for (int c = tid; c < (g.Wp * g.Jwp) >> 2; c += kTailThreads)
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(fdst + 16 * c), "l"(fsrc + 4 * c) : "memory");
  }
  for (int idx = tid; idx < 2 * (g.Jhp - g.Jh) * g.Wp; idx += kTailThreads) {  // pad rows jh >= Jh
    const int pass = idx / ((g.Jhp - g.Jh) * g.Wp), r = idx - pass * (g.Jhp - g.Jh) * g.Wp;
    t2s[(pass * g.Jhp + g.Jh) * g.Wp + r] = 0.f;
  }
  cp_async_wait_all();
  __syncthreads();
  if (g.Wp != g.W) {  // pad columns hold whatever the workspace held: zero them (0 * NaN would poison the sums)
    const int np = g.Wp - g.W;
    for (int idx = tid; idx < npass * g.Jh * np; idx += kTailThreads) {
      const int r = idx / np, c = idx - r * np;
      const int pass = r / g.Jh, jh = r - pass * g.Jh;
      t2s[(pass * g.Jhp + jh) * g.Wp + g.W + c] = 0.f;
    }
    __syncthreads();
  }
  // ---- W analysis: T[pass][jh][jw] = sum_w t2s[pass][jh][w] fw[w][jw]
  //      item = (pass, 4 rows jh, 4 columns jw, half of the w range); the two halves sit in adjacent lanes
  {
    const int gh = g.Jhp >> 2, gw = g.Jwp >> 2;
    const int nitem = npass * gh * gw * 2;
    const int wq = g.Wp >> 2;
    const int wq_half = (wq + 1) >> 1;
    for (int item0 = 0; item0 < nitem; item0 += kTailThreads) {
      const int item = item0 + tid;
      const bool act = item < nitem;
      const int half = item & 1;
      const int tile = act ? item >> 1 : 0;
      const int pass = tile / (gh * gw), r = tile - pass * gh * gw;
      const int hq = r / gw, wq4 = r - hq * gw;
      float2 acc[4][2];
#pragma unroll
      for (int a = 0; a < 4; ++a) acc[a][0] = acc[a][1] = make_float2(0.f, 0.f);
      if (act) {
        const float* tp = t2s + (pass * g.Jhp + 4 * hq) * g.Wp;
        const float* fp = fw + 4 * wq4;
        const int q0 = half * wq_half, q1 = half ? wq : wq_half;
        for (int qd = q0; qd < q1; ++qd) {
          float4 x[4];
#pragma unroll
          for (int a = 0; a < 4; ++a) x[a] = *reinterpret_cast<const float4*>(tp + a * g.Wp + 4 * qd);
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const float4 f = *reinterpret_cast<const float4*>(fp + (4 * qd + k) * g.Jwp);
            const float2 f01 = make_float2(f.x, f.y), f23 = make_float2(f.z, f.w);
#pragma unroll
            for (int a = 0; a < 4; ++a) {
              const float xv = k == 0 ? x[a].x : (k == 1 ? x[a].y : (k == 2 ? x[a].z : x[a].w));
              acc[a][0] = ffma2(dup2(xv), f01, acc[a][0]);
              acc[a][1] = ffma2(dup2(xv), f23, acc[a][1]);
            }
          }
        }
      }
      // both halves end up with the full sums; the even lane stores rows 0-1, the odd lane rows 2-3
#pragma unroll
      for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int c = 0; c < 2; ++c) {
          acc[a][c].x += __shfl_xor_sync(0xffffffffu, acc[a][c].x, 1);
          acc[a][c].y += __shfl_xor_sync(0xffffffffu, acc[a][c].y, 1);
        }
      if (act) {
        float* Tp = T + pass * g.Jh * g.Jw;
#pragma unroll
        for (int a2 = 0; a2 < 2; ++a2) {
          const int a = 2 * half + a2;
          const int jh = 4 * hq + a;
          if (jh < g.Jh) {
            const float2 v01 = half ? acc[2 + a2][0] : acc[a2][0];
            const float2 v23 = half ? acc[2 + a2][1] : acc[a2][1];
            const float v[4] = {v01.x, v01.y, v23.x, v23.y};
#pragma unroll
            for (int c = 0; c < 4; ++c)
              if (4 * wq4 + c < g.Jw) Tp[jh * g.Jw + 4 * wq4 + c] = v[c];
          }
        }
      }
    }
  }
  __syncthreads();
  // ---- 8-term recombination for kd = +u and -u (same expression as k_combine, T from shared memory)
  const int* kd_desc = pi + g.off_kdesc[0];
  const int* kh_desc = pi + g.off_kdesc[1];
  const int* kw_desc = pi + g.off_kdesc[2];
  const float* Tc = T;
  const float* Ts = T + g.Jh * g.Jw;
  for (int a = 0; a < 2; ++a) {
    const int kd = jd_desc[4 * jc + a];
    if (kd < 0) continue;
    const int sd = kd_desc[4 * kd + 1];
    const float gd = (float)kd_desc[4 * kd + 2];
    float* zo = Z + ((slab * g.Ld + kd) * g.Lh) * (long)g.Lw;
    for (int kh = tid >> 5; kh < g.Lh; kh += kTailThreads / 32)
     for (int kw = tid & 31; kw < g.Lw; kw += 32) {  // a warp per output row: no index divisions
      const int o = kh * g.Lw + kw;
      const int ch = kh_desc[4 * kh], sh = kh_desc[4 * kh + 1];
      const float gh = (float)kh_desc[4 * kh + 2];
      const int cw = kw_desc[4 * kw], sw = kw_desc[4 * kw + 1];
      const float gw = (float)kw_desc[4 * kw + 2];
      float v = Tc[ch * g.Jw + cw];
      if (sh >= 0 && sw >= 0) v -= gh * gw * Tc[sh * g.Jw + sw];
      if (sd >= 0 && sw >= 0) v -= gd * gw * Ts[ch * g.Jw + sw];
      if (sd >= 0 && sh >= 0) v -= gd * gh * Ts[sh * g.Jw + cw];
      if (sd >= 0) v += gd * Ts[ch * g.Jw + cw];
      if (sh >= 0) v += gh * Tc[sh * g.Jw + cw];
      if (sw >= 0) v += gw * Tc[ch * g.Jw + sw];
      if (sd >= 0 && sh >= 0 && sw >= 0) v -= gd * gh * gw * Ts[sh * g.Jw + sw];
      zo[o] = scale * v;
    }
  }
}

// adjoint: Z -> recombination^T -> W synthesis -> T2
__global__ void __launch_bounds__(kTailThreads) k_dht_tail_adj(const float* __restrict__ Z, float* __restrict__ T2,
                                                              const float* __restrict__ pf,
                                                              const int* __restrict__ pi, const TailGeom g,
                                                              float scale) {
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");  // the H-synthesis kernel may start its prologue
  extern __shared__ float4 smem4[];
  float* s = reinterpret_cast<float*>(smem4);
  float* fw = s;                         // [Jw][Wp]  (columns >= W are zero)
  float* Tt = fw + g.Jw * g.Wp;          // [2][Jw][Jhp]  transposed: one LDS.128 yields four rows jh
  const int tid = threadIdx.x;
  const int jc = blockIdx.x;
  const long slab = blockIdx.y;
  const int* jd_desc = pi + g.off_jdesc[0];
  const int* jh_desc = pi + g.off_jdesc[1];
  const int* jw_desc = pi + g.off_jdesc[2];
  const int js = find_sin_row_block(jd_desc, g.Jd, g.JCd, jc);
  const int npass = js >= 0 ? 2 : 1;
  {  // fw[jw][w]: the plan's zero-padded copy of the W rows ([Jw][Wp] floats, 16-byte pieces)
    const float* fsrc = pf + g.off_fullP_w;
    const unsigned fdst = (unsigned)__cvta_generic_to_shared(fw);
    for (int c = tid; c < (g.Jw * g.Wp) >> 2; c += kTailThreads)
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(fdst + 16 * c), "l"(fsrc + 4 * c) : "memory");
  }
  // ---- recombination^T (same expression as k_combine_t) for the rows jc and js
  const float* Zs = Z + slab * (long)g.Ld * g.Lh * g.Lw;
  for (int pass = 0; pass < npass; ++pass) {
    const int jd = pass == 0 ? jc : js;
    const int isd = jd_desc[4 * jd + 2];
    for (int jw = tid >> 5; jw < g.Jw; jw += kTailThreads / 32)
     for (int jh = tid & 31; jh < g.Jhp; jh += 32) {  // a warp per column of T: no index divisions
      const int o = jw * g.Jhp + jh;
      float acc = 0.f;
      float sign = 1.f;
      if (jh < g.Jh) {
        const int ish = jh_desc[4 * jh + 2], isw = jw_desc[4 * jw + 2];
        sign = (isd + ish + isw >= 2) ? -1.f : 1.f;
#pragma unroll
        for (int a = 0; a < 2; ++a) {
          const int kd = jd_desc[4 * jd + a];
          if (kd < 0) continue;
          const float fd = (isd && a == 1) ? -1.f : 1.f;
#pragma unroll
          for (int b = 0; b < 2; ++b) {
            const int kh = jh_desc[4 * jh + b];
            if (kh < 0) continue;
            const float fhs = (ish && b == 1) ? -fd : fd;
#pragma unroll
            for (int c = 0; c < 2; ++c) {
              const int kw = jw_desc[4 * jw + c];
              if (kw < 0) continue;
              const float fws = (isw && c == 1) ? -fhs : fhs;
              acc = fmaf(fws, __ldg(Zs + ((long)kd * g.Lh + kh) * g.Lw + kw), acc);
            }
          }
        }
      }
      Tt[pass * g.Jw * g.Jhp + o] = sign * scale * acc;
    }
  }
  cp_async_wait_all();
  __syncthreads();
  // ---- W synthesis: T2[jh][row][w] = sum_jw fw[jw][w] T[jh][jw]; tile = 4 rows jh x 4 columns w
  {
    const int gh = g.Jhp >> 2, q = g.Wp >> 2;
    const int nitem = npass * gh * q;
    for (int item = tid; item < nitem; item += kTailThreads) {
      const int pass = item / (gh * q), r = item - pass * gh * q;
      const int hq = r / q, wg = r - hq * q;
      const float* tp = Tt + pass * g.Jw * g.Jhp + 4 * hq;
      const float* fp = fw + 4 * wg;
      float2 acc[4][2];
#pragma unroll
      for (int a = 0; a < 4; ++a) acc[a][0] = acc[a][1] = make_float2(0.f, 0.f);
#pragma unroll 2
      for (int jw = 0; jw < g.Jw; ++jw) {
        const float4 x = *reinterpret_cast<const float4*>(tp + jw * g.Jhp);
        const float4 f = *reinterpret_cast<const float4*>(fp + jw * g.Wp);
        const float2 f01 = make_float2(f.x, f.y), f23 = make_float2(f.z, f.w);
        acc[0][0] = ffma2(dup2(x.x), f01, acc[0][0]);
        acc[0][1] = ffma2(dup2(x.x), f23, acc[0][1]);
        acc[1][0] = ffma2(dup2(x.y), f01, acc[1][0]);
        acc[1][1] = ffma2(dup2(x.y), f23, acc[1][1]);
        acc[2][0] = ffma2(dup2(x.z), f01, acc[2][0]);
        acc[2][1] = ffma2(dup2(x.z), f23, acc[2][1]);
        acc[3][0] = ffma2(dup2(x.w), f01, acc[3][0]);
        acc[3][1] = ffma2(dup2(x.w), f23, acc[3][1]);
      }
      const int row = pass == 0 ? jc : js;
#pragma unroll
      for (int a = 0; a < 4; ++a) {
        const int jh = 4 * hq + a;
        if (jh < g.Jh)
          *reinterpret_cast<float4*>(T2 + slab * g.slablen + (long)jh * g.rowlen + (long)row * g.Wp + 4 * wg) =
              make_float4(acc[a][0].x, acc[a][0].y, acc[a][1].x, acc[a][1].y);
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------ host side
static MidGeom make_geom(const DhtPlanHeader* h, long P) {
  MidGeom g;
  g.H = h->ax[1].n;
  g.W = h->ax[2].n;
  g.Jd = h->ax[0].J;
  g.JCd = h->ax[0].JC;
  g.Jh = h->ax[1].J;
  g.Jw = h->ax[2].J;
  g.Ld = h->ax[0].L;
  g.Lh = h->ax[1].L;
  g.Lw = h->ax[2].L;
  g.Jhp = r4(g.Jh);
  g.Jwp = r4(g.Jw);
  g.Hp = r4(g.H);
  g.Wp = r4(g.W);
  g.P = P;
  g.JCh = h->ax[1].JC;
  g.JSh = h->ax[1].JS;
  g.JChp = h->ax[1].JCp;
  g.JShp = h->ax[1].JSp;
  g.nhh = h->ax[1].nh;
  g.off_fcos_h = h->ax[1].off_fcos;
  g.off_fsin_h = h->ax[1].off_fsin;
  g.NHp = (g.nhh + 2) & ~1;
  g.off_full_h = h->ax[1].off_full;
  g.off_full_w = h->ax[2].off_full;
  for (int a = 0; a < 3; ++a) {
    g.off_kdesc[a] = h->ax[a].off_kdesc;
    g.off_jdesc[a] = h->ax[a].off_jdesc;
  }
  return g;
}

constexpr size_t kMidSmemLimit = 110 * 1024;  // at least two CTAs per SM

bool dht_mid_eligible(const void* plan_host, long P, int nslab) {
  const auto* h = reinterpret_cast<const DhtPlanHeader*>(plan_host);
  const MidGeom g = make_geom(h, P);
  if (nslab < 1 || nslab > 65535) return false;
  if (P % 4) return false;  // 16-byte cp.async of whole planes
  return (size_t)mid_smem_fwd(g).total * 4 <= kMidSmemLimit && (size_t)mid_smem_adj(g).total * 4 <= kMidSmemLimit;
}

// G1 [nslab][Jd][P] -> z [nslab][Ld][Lh][Lw]
int dht_mid_forward(const void* plan_host, const void* plan_dev, const float* G1, long P, float* z, int nslab,
                    float scale, cudaStream_t st) {
  const auto* h = reinterpret_cast<const DhtPlanHeader*>(plan_host);
  const MidGeom g = make_geom(h, P);
  const MidSmem sm = mid_smem_fwd(g);
  const size_t bytes = (size_t)sm.total * 4;
  if (bytes > 48 * 1024)
    HNO_CUDA(cudaFuncSetAttribute(k_dht_mid_fwd, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
  k_dht_mid_fwd<<<dim3(g.JCd, nslab), kMidThreads, bytes, st>>>(G1, z, reinterpret_cast<const float*>(plan_dev),
                                                                 reinterpret_cast<const int*>(plan_dev), g, sm, scale);
  HNO_LAUNCH_CHECK();
  return 0;
}

// z [nslab][Ld][Lh][Lw] -> G1 [nslab][Jd][P]
int dht_mid_adjoint(const void* plan_host, const void* plan_dev, const float* z, float* G1, long P, int nslab,
                    float scale, cudaStream_t st) {
  const auto* h = reinterpret_cast<const DhtPlanHeader*>(plan_host);
  const MidGeom g = make_geom(h, P);
  const MidSmem sm = mid_smem_adj(g);
  const size_t bytes = (size_t)sm.total * 4;
  if (bytes > 48 * 1024)
    HNO_CUDA(cudaFuncSetAttribute(k_dht_mid_adj, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
  k_dht_mid_adj<<<dim3(g.JCd, nslab), kMidThreads, bytes, st>>>(z, G1, reinterpret_cast<const float*>(plan_dev),
                                                                 reinterpret_cast<const int*>(plan_dev), g, sm, scale);
  HNO_LAUNCH_CHECK();
  return 0;
}

// ---- tails of the tensor-core H stage variant
static TailGeom make_tail_geom(const DhtPlanHeader* h) {
  TailGeom g;
  g.W = h->ax[2].n;
  g.Wp = r4(g.W);
  g.Jd = h->ax[0].J;
  g.JCd = h->ax[0].JC;
  g.Jh = h->ax[1].J;
  g.Jhp = r4(g.Jh);
  g.Jw = h->ax[2].J;
  g.Jwp = r4(g.Jw);
  g.Ld = h->ax[0].L;
  g.Lh = h->ax[1].L;
  g.Lw = h->ax[2].L;
  g.rowlen = (long)g.Jd * g.Wp;
  g.slablen = (long)g.Jh * g.rowlen;
  g.off_full_w = h->ax[2].off_full;
  g.off_fullT_w = h->ax[2].off_fullT;
  g.off_fullP_w = h->ax[2].off_fullP;
  for (int a = 0; a < 3; ++a) {
    g.off_kdesc[a] = h->ax[a].off_kdesc;
    g.off_jdesc[a] = h->ax[a].off_jdesc;
  }
  return g;
}
static size_t tail_smem_fwd(const TailGeom& g) {
  return (size_t)(2 * g.Jhp * g.Wp + g.Wp * g.Jwp + r4(2 * g.Jh * g.Jw)) * 4;
}
static size_t tail_smem_adj(const TailGeom& g) { return (size_t)(g.Jw * g.Wp + 2 * g.Jw * g.Jhp) * 4; }

bool dht_tail_eligible(const void* plan_host, int nslab) {
  const auto* h = reinterpret_cast<const DhtPlanHeader*>(plan_host);
  const TailGeom g = make_tail_geom(h);
  if (nslab < 1 || nslab > 65535) return false;
  return tail_smem_fwd(g) <= kMidSmemLimit && tail_smem_adj(g) <= kMidSmemLimit;
}

// T2 [nslab][Jh][Jd][Wp] -> z [nslab][Ld][Lh][Lw]
int dht_tail_forward(const void* plan_host, const void* plan_dev, const float* T2, float* z, int nslab, float scale,
                     cudaStream_t st) {
  const auto* h = reinterpret_cast<const DhtPlanHeader*>(plan_host);
  const TailGeom g = make_tail_geom(h);
  const size_t bytes = tail_smem_fwd(g);
  if (bytes > 48 * 1024)
    HNO_CUDA(cudaFuncSetAttribute(k_dht_tail_fwd, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
  k_dht_tail_fwd<<<dim3(g.JCd, nslab), kTailThreads, bytes, st>>>(T2, z, reinterpret_cast<const float*>(plan_dev),
                                                                   reinterpret_cast<const int*>(plan_dev), g, scale);
  HNO_LAUNCH_CHECK();
  return 0;
}

// z [nslab][Ld][Lh][Lw] -> T2 [nslab][Jh][Jd][Wp]
int dht_tail_adjoint(const void* plan_host, const void* plan_dev, const float* z, float* T2, int nslab, float scale,
                     cudaStream_t st) {
  const auto* h = reinterpret_cast<const DhtPlanHeader*>(plan_host);
  const TailGeom g = make_tail_geom(h);
  const size_t bytes = tail_smem_adj(g);
  if (bytes > 48 * 1024)
    HNO_CUDA(cudaFuncSetAttribute(k_dht_tail_adj, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
  k_dht_tail_adj<<<dim3(g.JCd, nslab), kTailThreads, bytes, st>>>(z, T2, reinterpret_cast<const float*>(plan_dev),
                                                                   reinterpret_cast<const int*>(plan_dev), g, scale);
  HNO_LAUNCH_CHECK();
  return 0;
}

}  // namespace hno
