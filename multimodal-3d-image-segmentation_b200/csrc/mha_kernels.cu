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

// z [B][cin][M], w [H][cd][cin], bias [H][cd] or null -> x_tok [B*H][Tp][Fp], x_chan [B*H][Fp][Tp]  (pre-zeroed buffers)
__global__ void __launch_bounds__(256) k_mha_project_fwd(const float* __restrict__ z, const float* __restrict__ w,
                                                         const float* __restrict__ bias, float* __restrict__ x_tok,
                                                         float* __restrict__ x_chan, MhaGeom g, int cin, int cd, int Fp) {
  const long idx = blockIdx.x * 256L + threadIdx.x;
  if (idx >= g.M) return;
  const int m = (int)idx;
  const int bh = blockIdx.y, b = bh / g.H, h = bh % g.H;
  int t, po;
  mha_locate(g, m, t, po);
  const int P = g.pd * g.ph * g.pw;
  const float* zp = z + (long)b * cin * g.M + m;
  const float* wp = w + (long)h * cd * cin;
  float* xt = x_tok + ((long)bh * g.Tp + t) * Fp + po;
  float* xc = x_chan + ((long)bh * Fp + po) * g.Tp + t;
  for (int c = 0; c < cd; ++c) {
    float acc = bias ? __ldg(bias + h * cd + c) : 0.f;
    for (int i = 0; i < cin; ++i) acc = fmaf(__ldg(wp + c * cin + i), __ldg(zp + (long)i * g.M), acc);
    xt[c * P] = acc;
    xc[(long)c * P * g.Tp] = acc;
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
    for (int c = 0; c < cd; ++c) acc = fmaf(__ldg(wp + c * cin), __ldg(dx + c * P), acc);
  }
  float* o = dz + ((long)b * cin + i) * g.M + m;
  *o = accumulate ? *o + acc : acc;
}

// dw [H][cd][cin] = sum_{b, m} dx_tok[b, h][t][c P + po] * z[b][i][m];  one CTA per (h, c, i); i == cin -> the bias gradient
__global__ void __launch_bounds__(256) k_mha_project_bwd_w(const float* __restrict__ dx_tok, const float* __restrict__ z,
                                                           float* __restrict__ dw, float* __restrict__ dbias, MhaGeom g,
                                                           int cin, int cd, int Fp) {
  __shared__ double sred[8];
  const int i = blockIdx.x, c = blockIdx.y, h = blockIdx.z;
  const bool is_bias = i == cin;
  const int P = g.pd * g.ph * g.pw;
  double acc = 0.0;
  for (int b = 0; b < g.B; ++b) {
    const float* dx = dx_tok + (long)(b * g.H + h) * g.Tp * Fp + c * P;
    const float* zp = z + ((long)b * cin + (is_bias ? 0 : i)) * g.M;
    float run = 0.f;
    int n = 0;
    for (int m = threadIdx.x; m < g.M; m += 256) {
      int t, po;
      mha_locate(g, m, t, po);
      const float d = __ldg(dx + (long)t * Fp + po);
      run = is_bias ? run + d : fmaf(d, __ldg(zp + m), run);
      if (++n == 32) {  // bounded fp32 runs, fp64 across them
        acc += (double)run;
        run = 0.f;
        n = 0;
      }
    }
    acc += (double)run;
  }
  acc = warp_sum_d(acc);
  if ((threadIdx.x & 31) == 0) sred[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    double tot = 0.0;
    for (int k = 0; k < 8; ++k) tot += sred[k];
    if (is_bias) dbias[h * cd + c] = (float)tot;
    else dw[((long)h * cd + c) * cin + i] = (float)tot;
  }
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
    for (int c = 0; c < cd; ++c) acc = fmaf(__ldg(wp + c), __ldg(op + c * P), acc);
  }
  y[((long)b * co + o) * g.M + m] = acc;
}

// dy [B][co][M] -> do_tok [B*H][Tp][Fp], do_chan [B*H][Fp][Tp] = sum_o wout[o][h cd + c] * dy[b][o][m]   (pre-zeroed)
__global__ void __launch_bounds__(256) k_mha_output_bwd_o(const float* __restrict__ dy, const float* __restrict__ wout,
                                                          float* __restrict__ do_tok, float* __restrict__ do_chan,
                                                          MhaGeom g, int co, int cd, int Fp) {
  const long idx = blockIdx.x * 256L + threadIdx.x;
  if (idx >= g.M) return;
  const int m = (int)idx;
  const int bh = blockIdx.y, b = bh / g.H, h = bh % g.H;
  int t, po;
  mha_locate(g, m, t, po);
  const int P = g.pd * g.ph * g.pw;
  const float* dp = dy + (long)b * co * g.M + m;
  float* xt = do_tok + ((long)bh * g.Tp + t) * Fp + po;
  float* xc = do_chan + ((long)bh * Fp + po) * g.Tp + t;
  for (int c = 0; c < cd; ++c) {
    float acc = 0.f;
    for (int o = 0; o < co; ++o) acc = fmaf(__ldg(wout + (long)o * g.H * cd + h * cd + c), __ldg(dp + (long)o * g.M), acc);
    xt[c * P] = acc;
    xc[(long)c * P * g.Tp] = acc;
  }
}

// dwout [co][H cd] = sum_{b, m} dy[b][o][m] * o_tok[b, h][t][c P + po];  one CTA per (j = h cd + c, o);  j == H cd -> bias
__global__ void __launch_bounds__(256) k_mha_output_bwd_w(const float* __restrict__ dy, const float* __restrict__ o_tok,
                                                          float* __restrict__ dwout, float* __restrict__ dbias, MhaGeom g,
                                                          int co, int cd, int Fp) {
  __shared__ double sred[8];
  const int j = blockIdx.x, o = blockIdx.y;
  const bool is_bias = j == g.H * cd;
  const int h = is_bias ? 0 : j / cd, c = is_bias ? 0 : j % cd;
  const int P = g.pd * g.ph * g.pw;
  double acc = 0.0;
  for (int b = 0; b < g.B; ++b) {
    const float* op = o_tok + (long)(b * g.H + h) * g.Tp * Fp + c * P;
    const float* dp = dy + ((long)b * co + o) * g.M;
    float run = 0.f;
    int n = 0;
    for (int m = threadIdx.x; m < g.M; m += 256) {
      const float d = __ldg(dp + m);
      if (is_bias) {
        run += d;
      } else {
        int t, po;
        mha_locate(g, m, t, po);
        run = fmaf(d, __ldg(op + (long)t * Fp + po), run);
      }
      if (++n == 32) {
        acc += (double)run;
        run = 0.f;
        n = 0;
      }
    }
    acc += (double)run;
  }
  acc = warp_sum_d(acc);
  if ((threadIdx.x & 31) == 0) sred[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    double tot = 0.0;
    for (int k = 0; k < 8; ++k) tot += sred[k];
    if (is_bias) dbias[o] = (float)tot;
    else dwout[(long)o * g.H * cd + j] = (float)tot;
  }
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
  const size_t bytes = (size_t)B * H * Tp * Fp * sizeof(float);
  HNO_CUDA(cudaMemsetAsync(x_tok, 0, bytes, st));
  HNO_CUDA(cudaMemsetAsync(x_chan, 0, bytes, st));
  dim3 grid(ceil_div(g.M, 256), B * H);
  k_mha_project_fwd<<<grid, 256, 0, st>>>(z, w, bias, x_tok, x_chan, g, cin, cd, Fp);
  HNO_LAUNCH_CHECK();
  return 0;
}

int mha_project_backward(const float* dx_tok, const float* z, const float* w, float* dz, float* dw, float* dbias, int B,
                         int H, int cin, int cd, int Ld, int Lh, int Lw, int pd, int ph, int pw, int Tp, int Fp,
                         int accumulate_dz, cudaStream_t st) {
  MhaGeom g;
  if (make_geom(&g, B, H, Ld, Lh, Lw, pd, ph, pw, Tp)) return -1;
  HNO_CHECK(dx_tok && z && w, "mha_project_backward: null pointer");
  HNO_CHECK(cin <= 65535 && B <= 65535, "mha_project_backward: too many channels");
  if (dz) {
    dim3 grid(ceil_div(g.M, 256), cin, B);
    k_mha_project_bwd_z<<<grid, 256, 0, st>>>(dx_tok, w, dz, g, cin, cd, Fp, accumulate_dz);
    HNO_LAUNCH_CHECK();
  }
  if (dw) {
    dim3 grid(cin + (dbias ? 1 : 0), cd, H);
    k_mha_project_bwd_w<<<grid, 256, 0, st>>>(dx_tok, z, dw, dbias, g, cin, cd, Fp);
    HNO_LAUNCH_CHECK();
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
  dim3 grid(ceil_div(g.M, 256), co, B);
  k_mha_output_fwd<<<grid, 256, 0, st>>>(o_tok, wout, bias, y, g, co, cd, Fp);
  HNO_LAUNCH_CHECK();
  return 0;
}

int mha_output_backward(const float* dy, const float* o_tok, const float* wout, float* do_tok, float* do_chan, float* dwout,
                        float* dbias, int B, int H, int co, int cd, int Ld, int Lh, int Lw, int pd, int ph, int pw, int Tp,
                        int Fp, cudaStream_t st) {
  MhaGeom g;
  if (make_geom(&g, B, H, Ld, Lh, Lw, pd, ph, pw, Tp)) return -1;
  HNO_CHECK(dy && o_tok && wout && do_tok && do_chan && dwout, "mha_output_backward: null pointer");
  HNO_CHECK(Fp >= cd * pd * ph * pw && Fp % 32 == 0, "mha_output_backward: bad feature pitch %d", Fp);
  const size_t bytes = (size_t)B * H * Tp * Fp * sizeof(float);
  HNO_CUDA(cudaMemsetAsync(do_tok, 0, bytes, st));
  HNO_CUDA(cudaMemsetAsync(do_chan, 0, bytes, st));
  dim3 grid(ceil_div(g.M, 256), B * H);
  k_mha_output_bwd_o<<<grid, 256, 0, st>>>(dy, wout, do_tok, do_chan, g, co, cd, Fp);
  HNO_LAUNCH_CHECK();
  dim3 gw(H * cd + (dbias ? 1 : 0), co);
  k_mha_output_bwd_w<<<gw, 256, 0, st>>>(dy, o_tok, dwout, dbias, g, co, cd, Fp);
  HNO_LAUNCH_CHECK();
  return 0;
}

}  // namespace hno
