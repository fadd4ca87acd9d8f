// extern "C" surface of libhno_b200.so (see include/hno_b200.h).  Thin: argument checks + dispatch.
#include "common.cuh"
#include "dht_plan.h"
#include "hno_b200.h"
#include "tc_stream.h"

#include <atomic>
#include <stdarg.h>
#include <stdio.h>

namespace hno {

static thread_local char g_err[1024] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

static std::atomic<long> g_launches{0};
void count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }

int sm_count() {
  static int cached = 0;
  if (cached > 0) return cached;
  int dev = 0, n = 0;
  if (cudaGetDevice(&dev) == cudaSuccess &&
      cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess && n > 0) {
    cached = n;
    return n;
  }
  (void)cudaGetLastError();
  return 148;  // B200; only used for sizing when no device is visible (CPU-side unit tests)
}

// implemented in the kernel translation units
int dht3_forward(const void*, const void*, const float*, long, long, float*, void*, int, float, cudaStream_t);
int dht3_adjoint(const void*, const void*, const float*, float*, long, long, void*, int, float, int, cudaStream_t);
size_t dht3_workspace_floats(const void*, long, int);
bool dht3_chain_eligible(const void*, const float*, long, long, int, int, int);
size_t dht3_chain_partials_bytes(const void*, int, int, int);
int dht3_chain(const void*, const void*, const float*, float*, long, long, const float* const*, float* const*, float*, void*,
               void*, int, int, int, float, float, int, int, int, cudaStream_t);
int pwconv_supported(int, int, int, int, int);
int pwconv_forward(const float*, const float*, const float*, const float*, float*, int, int, int, int, long, int, int,
                   cudaStream_t);
size_t pwconv_backward_workspace_bytes(int, int, int);
int pwconv_backward(const float*, const float*, const float*, const float*, const float*, float*, float*, float*,
                    float*, void*, int, int, int, int, long, long, long, int, int, int, cudaStream_t);
int modechain_supported(int C);
int modechain_forward(const float* z0, const float* const* weights, float* zs, int B, int C, long M, int L,
                      cudaStream_t st);
size_t modechain_backward_workspace_bytes(int B, int C, long M, int L);
int modechain_backward(const float* dzL, const float* z0, const float* zs, const float* const* weights, float* dz0,
                       float* const* dweights, void* workspace, int B, int C, long M, int L, int accumulate,
                       cudaStream_t st);
int hartley_conv_forward(const float*, const float*, float*, int, int, int, int, int, int, int, cudaStream_t);
int hartley_conv_backward(const float*, const float*, const float*, const float*, float*, float*, int, int, int, int,
                          int, int, int, cudaStream_t);
int hartley_conv_full_forward(const float*, const float*, const int*, float*, int, int, int, int, int, int, int, int, int,
                              int, cudaStream_t);
int hartley_conv_full_backward(const float*, const float*, const float*, const float*, const int*, float*, float*, int, int,
                               int, int, int, int, int, int, int, cudaStream_t);
int stem_supported(int, int);
int stem_forward(const float*, const float*, const float*, float*, int, int, int, int, int, int, long, cudaStream_t);
size_t stem_backward_workspace_bytes(int, int);
int stem_backward(const float*, const float*, float*, float*, void*, int, int, int, int, int, int, long, int,
                  cudaStream_t);
int stem_backward_input(const float*, const float*, float*, int, int, int, int, int, int, long, cudaStream_t);
size_t interp_tables_bytes(int, int, int, int, int, int);
int interp_tables_fill(void*, size_t, int, int, int, int, int, int);
int head_forward(const void*, const void*, const float*, float*, int, int, long, int, cudaStream_t);
int head_argmax(const void*, const void*, const float*, uint8_t*, int, int, long, cudaStream_t);
int head_direct_forward(const float*, float*, uint8_t*, int, int, int, int, int, long, int, cudaStream_t);
int head_direct_backward(const float*, const float*, float*, int, int, int, int, int, long, int, cudaStream_t);
int dsconv_forward(const float* const*, const int*, int, const float*, const float*, float*, const float*, float*, int, int,
                   long, int, cudaStream_t);
size_t dsconv_backward_workspace_bytes(int, int, int, long);
int dsconv_backward(const float* const*, float* const*, const int*, int, const float*, const float*, const float*,
                    const float*, float*, float*, float*, void*, int, int, long, long, long, int, cudaStream_t);
int mha_project_forward(const float*, const float*, const float*, float*, float*, int, int, int, int, int, int, int, int,
                        int, int, int, int, cudaStream_t);
int mha_project_backward(const float*, const float*, const float*, float*, float*, float*, void*, int, int, int, int, int,
                         int, int, int, int, int, int, int, int, cudaStream_t);
size_t mha_wgrad_workspace_bytes(int, int, int, int, int);
int mha_attention_forward(const float*, const float*, const float*, float*, float*, float*, int, int, int, int, float, int,
                          cudaStream_t);
int mha_attention_backward(const float*, const float*, const float*, const float*, const float*, const float*,
                           const float*, float*, float*, float*, float*, float*, int, int, int, int, float, int,
                           cudaStream_t);
int mha_output_forward(const float*, const float*, const float*, float*, int, int, int, int, int, int, int, int, int, int,
                       int, int, cudaStream_t);
int mha_output_backward(const float*, const float*, const float*, float*, float*, float*, float*, void*, int, int, int, int,
                        int, int, int, int, int, int, int, int, cudaStream_t);
size_t head_backward_workspace_bytes(const void*, int, int);
int head_backward(const void*, const void*, const float*, const float*, float*, void*, int, int, long, int,
                  cudaStream_t);
size_t loss_workspace_bytes(int, int);
int loss_forward(const float*, const float*, float*, float*, void*, int, int, long, int, float, cudaStream_t);
int loss_backward(const float*, const float*, const float*, const float*, float*, int, int, long, cudaStream_t);
int head_loss_forward(const void*, const void*, const float*, const uint8_t*, float*, float*, void*, int, int, long,
                      int, float, cudaStream_t);
size_t ce_loss_workspace_bytes(int);
int ce_loss_forward(const float*, const float*, const uint8_t*, float*, void*, int, int, long, cudaStream_t);
int ce_loss_backward(const float*, const float*, const uint8_t*, const float*, float*, int, int, long, cudaStream_t);
int head_loss_backward(const void*, const void*, const float*, const uint8_t*, const float*, const float*, float*,
                       void*, int, int, long, cudaStream_t);
size_t fourier_mix_workspace_bytes(int, int, long, int);
int fourier_mix_forward(const float*, const float*, const float*, const int*, const int*, const float*, float*, int, int, int, long,
                        long, int, cudaStream_t);
int fourier_mix_backward(const float*, const float*, const float*, const float*, const int*, const int*, const float*, float*,
                         float*, float*, void*, int, int, int, long, long, int, int, cudaStream_t);
int complex_modemix_forward(const float*, const float*, const float*, const float*, float*, float*, int, int, int, long,
                            cudaStream_t);
int complex_modemix_backward(const float*, const float*, const float*, const float*, const float*, const float*, float*,
                             float*, float*, float*, int, int, int, long, int, cudaStream_t);
int to_categorical(const void*, int, float*, int*, int, int, long, cudaStream_t);
size_t normalize_workspace_bytes(int);
int normalize_modalities(const void*, int, float*, void*, int, long, int, float, int, float, float, cudaStream_t);
int affine_resample_nn(const void*, void*, int, const double*, const int*, int, int, int, int, int, double, cudaStream_t);
int transpose2d(const void*, void*, int, long, int, int, cudaStream_t);
int adamax_step(float*, const float*, float*, float*, long, float, float, float, float, float, int, float,
                cudaStream_t);

}  // namespace hno

using namespace hno;
#define ST(s) reinterpret_cast<cudaStream_t>(s)

extern "C" {

int hno_version(void) { return HNO_B200_VERSION; }
int hno_set_tensor_cores(int enable) { return tc_set_enabled(enable); }
const char* hno_last_error(void) { return g_err; }

long hno_launch_count(int reset) {
  return reset ? g_launches.exchange(0) : g_launches.load();
}

int hno_device_check(void) {
  int dev = 0;
  HNO_CUDA(cudaGetDevice(&dev));
  int major = 0;
  HNO_CUDA(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev));
  HNO_CHECK(major == 10, "hno_b200 kernels are built for sm_100a only; device %d has compute capability %d.x", dev,
            major);
  return 0;
}

size_t hno_dht3_plan_bytes(int D, int H, int W, int Ld, int Lh, int Lw) {
  const int n[3] = {D, H, W}, L[3] = {Ld, Lh, Lw};
  return dht_plan_words(n, L) * 4;
}

int hno_dht3_plan_fill(void* host_buf, size_t bytes, int D, int H, int W, const int* kd, int Ld, const int* kh, int Lh,
                       const int* kw, int Lw) {
  HNO_CHECK(kd && kh && kw, "hno_dht3_plan_fill: null frequency list");
  const int n[3] = {D, H, W}, L[3] = {Ld, Lh, Lw};
  const int* const kl[3] = {kd, kh, kw};
  return dht_plan_fill(host_buf, bytes, n, kl, L);
}

size_t hno_dht3_workspace_bytes(const void* plan_host, long plane_pitch, int nslab) {
  if (!plan_host) return 0;
  return dht3_workspace_floats(plan_host, plane_pitch, nslab) * sizeof(float) + 256;
}

int hno_dht3_forward(const void* plan_host, const void* plan_dev, const float* x, long plane_pitch, long slab_stride,
                     float* z, void* workspace, int nslab, float scale, void* stream) {
  return dht3_forward(plan_host, plan_dev, x, plane_pitch, slab_stride, z, workspace, nslab, scale, ST(stream));
}

int hno_dht3_adjoint(const void* plan_host, const void* plan_dev, const float* z, float* x, long plane_pitch,
                     long slab_stride, void* workspace, int nslab, float scale, int epilogue, void* stream) {
  return dht3_adjoint(plan_host, plan_dev, z, x, plane_pitch, slab_stride, workspace, nslab, scale, epilogue,
                      ST(stream));
}

int hno_pwconv_supported(int ci1, int ci2, int co) { return pwconv_supported(ci1, ci2, co, -1, -1); }

int hno_pwconv_forward(const float* in1, const float* in2, const float* weight, const float* bias, float* out, int B,
                       int ci1, int ci2, int co, long S, int act, int residual, void* stream) {
  return pwconv_forward(in1, in2, weight, bias, out, B, ci1, ci2, co, S, act, residual, ST(stream));
}

size_t hno_pwconv_backward_workspace_bytes(int ci1, int ci2, int co) {
  return pwconv_backward_workspace_bytes(ci1, ci2, co);
}

int hno_pwconv_backward(const float* dy, const float* y, const float* in1, const float* in2, const float* weight,
                        float* din1, float* din2, float* dweight, float* dbias, void* workspace, int B, int ci1,
                        int ci2, int co, long S, long P, long HW, int act, int residual, int flags, void* stream) {
  return pwconv_backward(dy, y, in1, in2, weight, din1, din2, dweight, dbias, workspace, B, ci1, ci2, co, S, P, HW,
                         act, residual, flags, ST(stream));
}

int hno_modechain_supported(int C) { return modechain_supported(C); }

int hno_modechain_forward(const float* z0, const float* const* weights, float* zs, int B, int C, long M, int L,
                          void* stream) {
  return modechain_forward(z0, weights, zs, B, C, M, L, ST(stream));
}

size_t hno_modechain_backward_workspace_bytes(int B, int C, long M, int L) {
  return modechain_backward_workspace_bytes(B, C, M, L);
}

int hno_modechain_backward(const float* dzL, const float* z0, const float* zs, const float* const* weights, float* dz0,
                           float* const* dweights, void* workspace, int B, int C, long M, int L, int accumulate,
                           void* stream) {
  return modechain_backward(dzL, z0, zs, weights, dz0, dweights, workspace, B, C, M, L, accumulate, ST(stream));
}

int hno_hartley_conv_forward(const float* x, const float* w, float* out, int B, int ci, int co, int n0, int n1, int n2,
                             int residual_selu, void* stream) {
  return hartley_conv_forward(x, w, out, B, ci, co, n0, n1, n2, residual_selu, ST(stream));
}

int hno_hartley_conv_backward(const float* dout, const float* y, const float* x, const float* w, float* dx, float* dw,
                              int B, int ci, int co, int n0, int n1, int n2, int accumulate_dw, void* stream) {
  return hartley_conv_backward(dout, y, x, w, dx, dw, B, ci, co, n0, n1, n2, accumulate_dw, ST(stream));
}

int hno_stem_supported(int cin, int f) { return stem_supported(cin, f); }

int hno_stem_forward(const float* x, const float* weight, const float* bias, float* out, int B, int cin, int f, int Dx,
                     int Hx, int Wx, long P, void* stream) {
  return stem_forward(x, weight, bias, out, B, cin, f, Dx, Hx, Wx, P, ST(stream));
}

size_t hno_stem_backward_workspace_bytes(int cin, int f) { return stem_backward_workspace_bytes(cin, f); }

int hno_stem_backward(const float* dpre, const float* x, float* dweight, float* dbias, void* workspace, int B, int cin,
                      int f, int Dx, int Hx, int Wx, long P, int accumulate, void* stream) {
  return stem_backward(dpre, x, dweight, dbias, workspace, B, cin, f, Dx, Hx, Wx, P, accumulate, ST(stream));
}

int hno_stem_backward_input(const float* dpre, const float* weight, float* dx, int B, int cin, int f, int Dx, int Hx,
                            int Wx, long P, void* stream) {
  return stem_backward_input(dpre, weight, dx, B, cin, f, Dx, Hx, Wx, P, ST(stream));
}

size_t hno_interp_tables_bytes(int D, int H, int W, int Dx, int Hx, int Wx) {
  return interp_tables_bytes(D, H, W, Dx, Hx, Wx);
}

int hno_interp_tables_fill(void* host_buf, size_t bytes, int D, int H, int W, int Dx, int Hx, int Wx) {
  return interp_tables_fill(host_buf, bytes, D, H, W, Dx, Hx, Wx);
}

int hno_head_direct_forward(const float* logits, float* probs, int B, int C, int D, int H, int W, long P, int activation,
                            void* stream) {
  return head_direct_forward(logits, probs, nullptr, B, C, D, H, W, P, activation, ST(stream));
}

int hno_head_direct_argmax(const float* logits, unsigned char* labels, int B, int C, int D, int H, int W, long P,
                           void* stream) {
  return head_direct_forward(logits, nullptr, labels, B, C, D, H, W, P, 0, ST(stream));
}

int hno_head_direct_backward(const float* dprobs, const float* probs, float* dlogits, int B, int C, int D, int H, int W,
                             long P, int activation, void* stream) {
  return head_direct_backward(dprobs, probs, dlogits, B, C, D, H, W, P, activation, ST(stream));
}

int hno_head_argmax(const void* tables_host, const void* tables_dev, const float* logits_low, unsigned char* labels,
                    int B, int C, long P, void* stream) {
  return head_argmax(tables_host, tables_dev, logits_low, labels, B, C, P, ST(stream));
}

int hno_head_forward(const void* tables_host, const void* tables_dev, const float* logits_low, float* probs, int B,
                     int C, long P, int activation, void* stream) {
  return head_forward(tables_host, tables_dev, logits_low, probs, B, C, P, activation, ST(stream));
}

size_t hno_head_backward_workspace_bytes(const void* tables_host, int B, int C) {
  return head_backward_workspace_bytes(tables_host, B, C);
}

int hno_head_backward(const void* tables_host, const void* tables_dev, const float* dprobs, const float* probs,
                      float* dlogits_low, void* workspace, int B, int C, long P, int activation, void* stream) {
  return head_backward(tables_host, tables_dev, dprobs, probs, dlogits_low, workspace, B, C, P, activation,
                       ST(stream));
}

size_t hno_loss_workspace_bytes(int B, int C) { return loss_workspace_bytes(B, C); }

int hno_loss_forward(const float* y_pred, const float* y_true, float* loss, float* coef, void* workspace, int B, int C,
                     long N, int kind, float param, void* stream) {
  return loss_forward(y_pred, y_true, loss, coef, workspace, B, C, N, kind, param, ST(stream));
}

int hno_loss_backward(const float* y_pred, const float* y_true, const float* coef, const float* grad_loss,
                      float* dy_pred, int B, int C, long N, void* stream) {
  return loss_backward(y_pred, y_true, coef, grad_loss, dy_pred, B, C, N, ST(stream));
}

int hno_head_loss_forward(const void* tables_host, const void* tables_dev, const float* logits_low,
                          const uint8_t* labels, float* loss, float* coef, void* workspace, int B, int C, long P,
                          int kind, float param, void* stream) {
  return head_loss_forward(tables_host, tables_dev, logits_low, labels, loss, coef, workspace, B, C, P, kind, param,
                           ST(stream));
}

int hno_head_loss_backward(const void* tables_host, const void* tables_dev, const float* logits_low,
                           const uint8_t* labels, const float* coef, const float* grad_loss, float* dlogits_low,
                           void* workspace, int B, int C, long P, void* stream) {
  return head_loss_backward(tables_host, tables_dev, logits_low, labels, coef, grad_loss, dlogits_low, workspace, B, C,
                            P, ST(stream));
}

int hno_complex_modemix_forward(const float* re, const float* im, const float* w_real, const float* w_imag, float* a,
                                float* b, int B, int ci, int co, long M, void* stream) {
  return complex_modemix_forward(re, im, w_real, w_imag, a, b, B, ci, co, M, ST(stream));
}

int hno_complex_modemix_backward(const float* da, const float* db, const float* re, const float* im,
                                 const float* w_real, const float* w_imag, float* dre, float* dim, float* dw_real,
                                 float* dw_imag, int B, int ci, int co, long M, int accumulate_dw, void* stream) {
  return complex_modemix_backward(da, db, re, im, w_real, w_imag, dre, dim, dw_real, dw_imag, B, ci, co, M,
                                  accumulate_dw, ST(stream));
}

size_t hno_fourier_mix_workspace_bytes(int ci, int co, long MK, int B) { return fourier_mix_workspace_bytes(ci, co, MK, B); }

int hno_fourier_mix_forward(const float* z, const float* w_real, const float* w_imag, const int* lin_k, const int* lin_n,
                            const float* ck, float* hp, int B, int ci, int co, long MK, long MS, int individual, void* stream) {
  return fourier_mix_forward(z, w_real, w_imag, lin_k, lin_n, ck, hp, B, ci, co, MK, MS, individual, ST(stream));
}

int hno_fourier_mix_backward(const float* dhp, const float* z, const float* w_real, const float* w_imag, const int* lin_k,
                             const int* lin_n, const float* ck, float* dz, float* dw_real, float* dw_imag, void* workspace,
                             int B, int ci, int co, long MK, long MS, int individual, int accumulate_dw, void* stream) {
  return fourier_mix_backward(dhp, z, w_real, w_imag, lin_k, lin_n, ck, dz, dw_real, dw_imag, workspace, B, ci, co, MK, MS,
                              individual, accumulate_dw, ST(stream));
}

size_t hno_ce_loss_workspace_bytes(int B) { return ce_loss_workspace_bytes(B); }

int hno_ce_loss_forward(const float* y_pred, const float* y_true, const uint8_t* labels, float* loss, void* workspace,
                        int B, int C, long N, void* stream) {
  return ce_loss_forward(y_pred, y_true, labels, loss, workspace, B, C, N, ST(stream));
}

int hno_ce_loss_backward(const float* y_pred, const float* y_true, const uint8_t* labels, const float* grad_loss,
                         float* dy_pred, int B, int C, long N, void* stream) {
  return ce_loss_backward(y_pred, y_true, labels, grad_loss, dy_pred, B, C, N, ST(stream));
}

int hno_to_categorical(const void* labels, int label_bytes, float* onehot, int* bad_count, int B, int C, long N,
                       void* stream) {
  return to_categorical(labels, label_bytes, onehot, bad_count, B, C, N, ST(stream));
}

size_t hno_normalize_workspace_bytes(int rows) { return normalize_workspace_bytes(rows); }

int hno_normalize_modalities(const float* data, float* out, void* workspace, int rows, long n, int has_mask,
                             float mask_val, int has_clip, float clip_lo, float clip_hi, void* stream) {
  return normalize_modalities(data, 4, out, workspace, rows, n, has_mask, mask_val, has_clip, clip_lo, clip_hi,
                              ST(stream));
}

int hno_normalize_modalities_i16(const short* data, float* out, void* workspace, int rows, long n, int has_mask,
                                 float mask_val, int has_clip, float clip_lo, float clip_hi, void* stream) {
  return normalize_modalities(data, 2, out, workspace, rows, n, has_mask, mask_val, has_clip, clip_lo, clip_hi,
                              ST(stream));
}

int hno_transpose2d(const void* in, void* out, int elem_bytes, long n, int R, int C, void* stream) {
  return transpose2d(in, out, elem_bytes, n, R, C, ST(stream));
}

int hno_affine_resample_nn(const void* in, void* out, int elem_bytes, const double* xform, const int* flags, int B,
                           int C, int D, int H, int W, double cval, void* stream) {
  return affine_resample_nn(in, out, elem_bytes, xform, flags, B, C, D, H, W, cval, ST(stream));
}

int hno_hartley_conv_full_forward(const float* x_ext, const float* weight, const int* partner_table, float* out, int B,
                                  int ci, int co, int n0, int n1, int n2, int e0, int e1, int e2, int act, void* stream) {
  return hartley_conv_full_forward(x_ext, weight, partner_table, out, B, ci, co, n0, n1, n2, e0, e1, e2, act, ST(stream));
}
int hno_hartley_conv_full_backward(const float* dout, const float* y, const float* x_ext, const float* weight,
                                   const int* partner_table, float* dx_ext, float* dweight, int B, int ci, int co, int n0,
                                   int n1, int n2, int e0, int e1, int e2, void* stream) {
  return hartley_conv_full_backward(dout, y, x_ext, weight, partner_table, dx_ext, dweight, B, ci, co, n0, n1, n2, e0, e1,
                                    e2, ST(stream));
}

int hno_dht3_chain_eligible(const void* plan_host, const float* x, long plane_pitch, long slab_stride, int B, int C, int L) {
  return dht3_chain_eligible(plan_host, x, plane_pitch, slab_stride, B, C, L) ? 1 : 0;
}
size_t hno_dht3_chain_partials_bytes(const void* plan_host, int C, int L, int B) {
  return dht3_chain_partials_bytes(plan_host, C, L, B);
}
int hno_dht3_chain_forward(const void* plan_host, const void* plan_dev, const float* x, float* out, long plane_pitch,
                           long slab_stride, const float* const* weights, float* z_all, void* workspace, int B, int C, int L,
                           float scale_in, int epilogue, void* stream) {
  return dht3_chain(plan_host, plan_dev, x, out, plane_pitch, slab_stride, weights, nullptr, z_all, workspace, nullptr, B, C,
                    L, scale_in, 1.f, epilogue, 0, 0, ST(stream));
}
int hno_dht3_chain_backward(const void* plan_host, const void* plan_dev, const float* dt, float* out, long plane_pitch,
                            long slab_stride, const float* const* weights, float* const* dweights, const float* z_all,
                            void* workspace, void* partials, int B, int C, int L, float scale_out, int epilogue,
                            int accumulate_dw, void* stream) {
  return dht3_chain(plan_host, plan_dev, dt, out, plane_pitch, slab_stride, weights, dweights, const_cast<float*>(z_all),
                    workspace, partials, B, C, L, 1.f, scale_out, epilogue, 1, accumulate_dw, ST(stream));
}

int hno_dsconv_forward(const float* const* in, const int* ch, int n, const float* weight, const float* bias, float* out,
                       const float* weight2, float* out2, int B, int CO, long S, int act, void* stream) {
  return dsconv_forward(in, ch, n, weight, bias, out, weight2, out2, B, CO, S, act, ST(stream));
}
size_t hno_dsconv_backward_workspace_bytes(int ctot, int CO, int B, long S) {
  return dsconv_backward_workspace_bytes(ctot, CO, B, S);
}
int hno_dsconv_backward(const float* const* in, float* const* din, const int* ch, int n, const float* weight,
                        const float* dy, const float* y, const float* weight2, float* dweight, float* dbias,
                        float* dweight2, void* workspace, int B, int CO, long S, long P, long HW, int act, void* stream) {
  return dsconv_backward(in, din, ch, n, weight, dy, y, weight2, dweight, dbias, dweight2, workspace, B, CO, S, P, HW,
                         act, ST(stream));
}

int hno_mha_project_forward(const float* z, const float* weight, const float* bias, float* x_tok, float* x_chan, int B,
                            int H, int cin, int cd, int Ld, int Lh, int Lw, int pd, int ph, int pw, int Tp, int Fp,
                            void* stream) {
  return mha_project_forward(z, weight, bias, x_tok, x_chan, B, H, cin, cd, Ld, Lh, Lw, pd, ph, pw, Tp, Fp, ST(stream));
}
size_t hno_mha_wgrad_workspace_bytes(int B, int H, int cin, int cd, int T) { return mha_wgrad_workspace_bytes(B, H, cin, cd, T); }
int hno_mha_project_backward(const float* dx_tok, const float* z, const float* weight, float* dz, float* dweight,
                             float* dbias, void* workspace, int B, int H, int cin, int cd, int Ld, int Lh, int Lw, int pd,
                             int ph, int pw, int Tp, int Fp, int accumulate_dz, void* stream) {
  return mha_project_backward(dx_tok, z, weight, dz, dweight, dbias, workspace, B, H, cin, cd, Ld, Lh, Lw, pd, ph, pw, Tp,
                              Fp, accumulate_dz, ST(stream));
}
int hno_mha_attention_forward(const float* q_tok, const float* k_tok, const float* v_chan, float* P, float* PT,
                              float* o_tok, int BH, int Tp, int Fqp, int Fvp, float scale, int activation, void* stream) {
  return mha_attention_forward(q_tok, k_tok, v_chan, P, PT, o_tok, BH, Tp, Fqp, Fvp, scale, activation, ST(stream));
}
int hno_mha_attention_backward(const float* do_tok, const float* do_chan, const float* q_chan, const float* k_chan,
                               const float* v_tok, const float* P, const float* PT, float* dS, float* dST, float* dq_tok,
                               float* dk_tok, float* dv_tok, int BH, int Tp, int Fqp, int Fvp, float scale, int activation,
                               void* stream) {
  return mha_attention_backward(do_tok, do_chan, q_chan, k_chan, v_tok, P, PT, dS, dST, dq_tok, dk_tok, dv_tok, BH, Tp,
                                Fqp, Fvp, scale, activation, ST(stream));
}
int hno_mha_output_forward(const float* o_tok, const float* weight_out, const float* bias, float* y, int B, int H, int co,
                           int cd, int Ld, int Lh, int Lw, int pd, int ph, int pw, int Tp, int Fp, void* stream) {
  return mha_output_forward(o_tok, weight_out, bias, y, B, H, co, cd, Ld, Lh, Lw, pd, ph, pw, Tp, Fp, ST(stream));
}
int hno_mha_output_backward(const float* dy, const float* o_tok, const float* weight_out, float* do_tok, float* do_chan,
                            float* dweight_out, float* dbias, void* workspace, int B, int H, int co, int cd, int Ld, int Lh,
                            int Lw, int pd, int ph, int pw, int Tp, int Fp, void* stream) {
  return mha_output_backward(dy, o_tok, weight_out, do_tok, do_chan, dweight_out, dbias, workspace, B, H, co, cd, Ld, Lh, Lw,
                             pd, ph, pw, Tp, Fp, ST(stream));
}

int hno_adamax_step(float* param, const float* grad, float* exp_avg, float* exp_inf, long n, float lr, float beta1,
                    float beta2, float eps, float weight_decay, int step, float grad_scale, void* stream) {
  return adamax_step(param, grad, exp_avg, exp_inf, n, lr, beta1, beta2, eps, weight_decay, step, grad_scale,
                     ST(stream));
}

}  // extern "C"
