/* hno_b200 — C ABI of the B200 (sm_100a) spectral hot path for HNOSeg-XS.
 *
 * One shared object, libhno_b200.so, plain pointers and sizes, no torch types.  The reference
 * (IBM/multimodal-3d-image-segmentation) has no FFI of its own: its boundary for this path is the
 * Python module API of nets/ (SURVEY.md 8b).  Each entry point below names the reference code it
 * replaces; INTEGRATION.md shows the ctypes binding a maintainer of the reference would add.
 *
 * Conventions
 *   - every function returns 0 on success and a negative value on error; the message is
 *     available (per thread) from hno_last_error().  Nothing throws, nothing aborts.
 *   - all device work is enqueued on the `stream` argument (a cudaStream_t passed as void*); no
 *     implicit synchronisation, no allocation: the caller owns every buffer incl. workspaces.
 *   - activations are fp32, channel-planar  [B][C][D][P]  where P >= H*W is the plane pitch in
 *     floats (P == H*W for dense NCDHW tensors; the model pads P to a multiple of 32 so that
 *     128-bit accesses are legal on odd grids such as 121x121x78).  Columns [H*W, P) are padding.
 *   - "S" is the per-channel extent D*P.
 */
#ifndef HNO_B200_H_
#define HNO_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define HNO_B200_VERSION 100

int hno_version(void);
const char* hno_last_error(void);
/* 0 if the current CUDA device is compute capability 10.x (B200), error otherwise. */
int hno_device_check(void);
/* Number of CUDA kernels this library has launched in the process (optionally reset to 0). */
long hno_launch_count(int reset);
/* The HBM-bound contractions (pointwise convolutions, D-axis stages of the truncated DHT) run on the tcgen05 tensor
 * cores with 3xTF32 operand splitting whenever strides allow TMA (16-byte multiples).  This switch forces the fp32
 * CUDA-core kernels instead (A/B measurements, tests); returns the previous setting.  Default: enabled. */
int hno_set_tensor_cores(int enable);

/* ------------------------------------------------------------------------------------------
 * Truncated 3-D discrete Hartley transform.
 *   replaces nets/dht.py:16-66 (dhtn/dht3), nets/hnosegxs.py:332-410 (TransformCrop),
 *            nets/hnosegxs.py:413-494 (PadInverse)
 * A plan is a position-independent table blob built on the host (no GPU needed) for a grid
 * (D,H,W) and, per axis, the list of retained frequency indices in OUTPUT order (for
 * TransformCrop with m modes on an axis of length n: 0..m-1, n-m..n-1).  Upload the same bytes
 * to the device and pass both copies to the transform calls.
 * ------------------------------------------------------------------------------------------ */
size_t hno_dht3_plan_bytes(int D, int H, int W, int Ld, int Lh, int Lw);
int hno_dht3_plan_fill(void* host_buf, size_t bytes, int D, int H, int W, const int* kd, int Ld, const int* kh,
                       int Lh, const int* kw, int Lw);
size_t hno_dht3_workspace_bytes(const void* plan_host, long plane_pitch, int nslab);

/* z[slab][Ld][Lh][Lw] = scale * C * x[slab][D][P]     (TransformCrop forward with scale = 1/(D*H*W);
 *                                                        PadInverse backward with scale = 1) */
int hno_dht3_forward(const void* plan_host, const void* plan_dev, const float* x, long plane_pitch,
                     long slab_stride, float* z, void* workspace, int nslab, float scale, void* stream);
/* x[slab][D][P] (op)= scale * C^T * z                 (PadInverse forward with scale = 1;
 *                                                        TransformCrop backward with scale = 1/(D*H*W))
 * epilogue: 0 store, 1 accumulate into x, 2 store selu(.) (nets/hnosegxs.py:267-268 fused),
 *           3 x = selu(x + .) (nets/architectures.py:521-536: spectral branch + conv branch, then the activation) */
int hno_dht3_adjoint(const void* plan_host, const void* plan_dev, const float* z, float* x, long plane_pitch,
                     long slab_stride, void* workspace, int nslab, float scale, int epilogue, void* stream);

/* ------------------------------------------------------------------------------------------
 * One HNO-XS block's spectral part in one call       replaces nets/hnosegxs.py:260-263 (+ the SELU of :267-268):
 *   transform_crop -> n_XS x NeuralOperatorBlock (shared weights) -> pad_inverse
 * forward :  out = EPI( C^T chain( scale_in * C x ) ),  z_all [L+1][B][C][Ld][Lh][Lw] receives z_0 .. z_L (may be null)
 * backward:  out (op)= scale_out * C^T chain_bwd( C dt ),  dweights[l] [C][C] written (or accumulated)
 * Five launches (D analysis, H analysis, spectral core, H synthesis, D synthesis): the W stages, the cas recombination and
 * the mixes run in ONE kernel whose CTAs own the modes of one (sample, |u_d|, |u_h|) for all channels (csrc/spectral_core.cu).
 * x, out: planar [B][C][D][pitch]; weights / dweights: L device pointers to [C][C] matrices; epilogue as hno_dht3_adjoint;
 * workspace: hno_dht3_workspace_bytes; partials (backward only): hno_dht3_chain_partials_bytes.
 * hno_dht3_chain_eligible == 0: use hno_dht3_forward / hno_modechain_* / hno_dht3_adjoint instead.
 * ------------------------------------------------------------------------------------------ */
int hno_dht3_chain_eligible(const void* plan_host, const float* x, long plane_pitch, long slab_stride, int B, int C, int L);
size_t hno_dht3_chain_partials_bytes(const void* plan_host, int C, int L, int B);
int hno_dht3_chain_forward(const void* plan_host, const void* plan_dev, const float* x, float* out, long plane_pitch,
                           long slab_stride, const float* const* weights, float* z_all, void* workspace, int B, int C, int L,
                           float scale_in, int epilogue, void* stream);
int hno_dht3_chain_backward(const void* plan_host, const void* plan_dev, const float* dt, float* out, long plane_pitch,
                            long slab_stride, const float* const* weights, float* const* dweights, const float* z_all,
                            void* workspace, void* partials, int B, int C, int L, float scale_out, int epilogue,
                            int accumulate_dw, void* stream);

/* ------------------------------------------------------------------------------------------
 * Pointwise (1x1x1) channel mixing  y = act( W * [in1 ; in2] + bias (+ in1 if residual) )
 *   replaces nets/nets_utils.py:120-174 (ConvNormAct, kernel_size 1, SELU, no norm),
 *            nets/hartley_operator.py:287-292 + nets/hnosegxs.py:307-329 (shared-weight mode mix
 *            + residual + SELU: residual=1, bias=NULL), nets/hnosegxs.py:273-275 (concat conv:
 *            in1 = activated inverse transform, in2 = block input) and :254-255 (mapping conv).
 * Tensors are [B][C][S]; W is [CO][CI1+CI2] row-major (Conv3d weight flattened / HartleyOperator
 * weight); act: 0 none, 1 SELU.  P/HW describe padding columns (s % P >= HW), which contribute
 * nothing to weight gradients.  Supported channel counts: see hno_pwconv_supported().
 * ------------------------------------------------------------------------------------------ */
int hno_pwconv_supported(int ci1, int ci2, int co);
int hno_pwconv_forward(const float* in1, const float* in2, const float* weight, const float* bias, float* out,
                       int B, int ci1, int ci2, int co, long S, int act, int residual, void* stream);
size_t hno_pwconv_backward_workspace_bytes(int ci1, int ci2, int co);
/* dy, y: gradient and saved OUTPUT of the forward.  din1/din2 may be NULL (not needed).
 * flags bit0: accumulate into din1, bit1: accumulate into din2,
 *       bit2: in1 is itself a SELU output, return d/d(pre-activation of in1) in din1,
 *       bit3: accumulate into dweight/dbias instead of overwriting.
 * dweight [CO][CI1+CI2], dbias [CO] (NULL when the layer has no bias). */
int hno_pwconv_backward(const float* dy, const float* y, const float* in1, const float* in2, const float* weight,
                        float* din1, float* din2, float* dweight, float* dbias, void* workspace, int B, int ci1,
                        int ci2, int co, long S, long P, long HW, int act, int residual, int flags, void* stream);

/* ------------------------------------------------------------------------------------------
 * Per-mode ("individual") Hartley mixing with even/odd recombination
 *   replaces nets/hartley_operator.py:293-299,302-333 (hartley_conv + get_reverse on the cropped block)
 *   out(k) = 1/2 [ W(k) (X(k)+X(-k)) + W(-k) (X(k)-X(-k)) ],  -k = (n-j) mod n per axis.
 * x [B][CI][M], w [CO][CI][M], out [B][CO][M], M = n0*n1*n2.
 * residual_selu != 0 fuses NeuralOperatorBlock (nets/hnosegxs.py:307-329): out = selu(mix(x) + x); the
 * backward then takes the saved OUTPUT y (NULL for the plain op) and dout = gradient of that output.
 * ------------------------------------------------------------------------------------------ */
int hno_hartley_conv_forward(const float* x, const float* w, float* out, int B, int ci, int co, int n0, int n1,
                             int n2, int residual_selu, void* stream);
int hno_hartley_conv_backward(const float* dout, const float* y, const float* x, const float* w, float* dx,
                              float* dw, int B, int ci, int co, int n0, int n1, int n2, int accumulate_dw,
                              void* stream);
/* HartleyOperator(use_transform=True, weights_type='individual')  replaces nets/hartley_operator.py:196-241: the reversal
 * partner of X lives in the FULL spectrum (x_reverse = get_reverse(dht3(x)), :199), the weight is reversed inside its 2m block
 * (:200).  x_ext [B][ci][e0][e1][e2]: the retained modes plus, last on every axis with n > 2m, the frequency +m (the partner of
 * n - m); out / weight over the retained block [n0][n1][n2]; partner_table = [r0 (n0 ints) | r1 (n1) | r2 (n2)], device memory:
 * position in x_ext of the partner of retained position j.  act: 0 none, 1 SELU (the reference activates the padded spectrum,
 * :267).  backward: dout = gradient of out, y = out when act == 1 (else null); dx_ext is zero-filled and written. */
int hno_hartley_conv_full_forward(const float* x_ext, const float* weight, const int* partner_table, float* out, int B,
                                  int ci, int co, int n0, int n1, int n2, int e0, int e1, int e2, int act, void* stream);
int hno_hartley_conv_full_backward(const float* dout, const float* y, const float* x_ext, const float* weight,
                                   const int* partner_table, float* dx_ext, float* dweight, int B, int ci, int co, int n0,
                                   int n1, int n2, int e0, int e1, int e2, void* stream);

/* ------------------------------------------------------------------------------------------
 * Complex per-mode ("individual" weights) mixing of the retained Fourier half-spectrum
 *   replaces nets/fourier_operator.py:165-187 (weights_type == 'individual': four corner einsums
 *   'oidhw,bidhw->bodhw' with weight = complex(weight_real, weight_imag), :155) on real / imaginary
 *   mode tensors re, im [B][ci][M] -> a, b [B][co][M];  w_real, w_imag [co][ci][M] with
 *   M = 2 m0 * 2 m1 * m2 in the weights' own order (:67-72):   a + i b = (w_real + i w_imag)(re + i im).
 *   backward: dre/dim and dw_real/dw_imag are optional PAIRS (both NULL or both set);
 *   accumulate_dw != 0 adds into dw_*.
 * ------------------------------------------------------------------------------------------ */
int hno_complex_modemix_forward(const float* re, const float* im, const float* w_real, const float* w_imag,
                                float* a, float* b, int B, int ci, int co, long M, void* stream);
int hno_complex_modemix_backward(const float* da, const float* db, const float* re, const float* im,
                                 const float* w_real, const float* w_imag, float* dre, float* dim,
                                 float* dw_real, float* dw_imag, int B, int ci, int co, long M,
                                 int accumulate_dw, void* stream);

/* ------------------------------------------------------------------------------------------
 * Fourier spectral layer, shared complex weights (FNOSeg, BASELINE config 3): the mode-domain step between the two transforms
 *   replaces nets/fourier_operator.py:155, 165-209 (corner slicing of the rfftn output, 'oi,bi...->bo...' with
 *            w_real + i w_imag, zero-padded re-assembly before irfftn), evaluated on Hartley coefficients:
 *   z [B][ci][MS] = (1/N) DHT of the input on the symmetric mode set S (hno_dht3_forward);  lin_k / lin_n [MK]: positions in S
 *   of every rfft half-grid mode k and of its mirror image N - k;  ck [MK]: 1 on the k_w = 0 plane, else 2.
 *   re = (z[k] + z[N-k]) / 2, im = (z[N-k] - z[k]) / 2;  a + i b = (w_real + i w_imag)(re + i im);
 *   hp[k] += ck (a - b) / 2,  hp[N-k] += ck (a + b) / 2  ->  hp [B][co][MS], whose hno_dht3_adjoint is the layer output.
 *   individual == 0: shared weights w_* [co][ci];  individual != 0: per-mode weights w_* [co][ci][MK] (config_fno.ini, :165-187).
 *   ci, co multiples of 4.  backward: dz and the (dw_real, dw_imag) pair are optional; workspace (shared weights only):
 *   hno_fourier_mix_workspace_bytes.
 * ------------------------------------------------------------------------------------------ */
size_t hno_fourier_mix_workspace_bytes(int ci, int co, long MK, int B);
int hno_fourier_mix_forward(const float* z, const float* w_real, const float* w_imag, const int* lin_k, const int* lin_n,
                            const float* ck, float* hp, int B, int ci, int co, long MK, long MS, int individual, void* stream);
int hno_fourier_mix_backward(const float* dhp, const float* z, const float* w_real, const float* w_imag, const int* lin_k,
                             const int* lin_n, const float* ck, float* dz, float* dw_real, float* dw_imag, void* workspace,
                             int B, int ci, int co, long MK, long MS, int individual, int accumulate_dw, void* stream);

/* ------------------------------------------------------------------------------------------
 * The n_XS shared-weight mixes of one HNO-XS block as one launch
 *   replaces nets/hnosegxs.py:261-262 (loop over NeuralOperatorBlock.forward :307-329) with
 *            nets/hartley_operator.py:287-292 (weights_type 'shared'):  z_l = selu(W_l z_{l-1} + z_{l-1}), l = 1..L
 * z0 [B][C][M]; weights: HOST array of L device pointers to [C][C] matrices; zs [L][B][C][M] receives z_1..z_L
 * (all of them are needed by the backward).  C in {8, 24}, L <= 8.
 * backward: dzL = gradient of z_L; writes dz0 [B][C][M] and dweights[l] [C][C] (HOST array of L device pointers;
 * accumulate != 0 adds to them).
 * ------------------------------------------------------------------------------------------ */
int hno_modechain_supported(int C);
int hno_modechain_forward(const float* z0, const float* const* weights, float* zs, int B, int C, long M, int L,
                          void* stream);
size_t hno_modechain_backward_workspace_bytes(int B, int C, long M, int L);
int hno_modechain_backward(const float* dzL, const float* z0, const float* zs, const float* const* weights, float* dz0,
                           float* const* dweights, void* workspace, int B, int C, long M, int L, int accumulate,
                           void* stream);

/* ------------------------------------------------------------------------------------------
 * Stem: Conv3d(k=2, s=2, p=1) + bias + SELU      replaces nets/hnosegxs.py:102-105,150-151
 *   x [B][CIN][Dx][Hx][Wx] dense  ->  out [B][F][D][P],  D = Dx/2+1, H = Hx/2+1, W = Wx/2+1
 *   weight [F][CIN][2][2][2], bias [F].  Padding columns of out are written as 0.
 * ------------------------------------------------------------------------------------------ */
int hno_stem_supported(int cin, int f);
int hno_stem_forward(const float* x, const float* weight, const float* bias, float* out, int B, int cin, int f,
                     int Dx, int Hx, int Wx, long P, void* stream);
size_t hno_stem_backward_workspace_bytes(int cin, int f);
/* dpre: gradient w.r.t. the PRE-activation of the stem output, [B][F][D][P]. */
int hno_stem_backward(const float* dpre, const float* x, float* dweight, float* dbias, void* workspace, int B,
                      int cin, int f, int Dx, int Hx, int Wx, long P, int accumulate, void* stream);
/* gradient w.r.t. the image (the transposed convolution; autograd of nets/hnosegxs.py:150-151 w.r.t. x):
 * dx [B][cin][Dx][Hx][Wx] = sum_o weight[o][.][k] dpre[o][(z + 1) / 2], k = (z + 1) % 2 per axis.  Overwrites dx. */
int hno_stem_backward_input(const float* dpre, const float* weight, float* dx, int B, int cin, int f, int Dx, int Hx,
                            int Wx, long P, void* stream);

/* ------------------------------------------------------------------------------------------
 * Head: trilinear up-sampling (F.interpolate, align_corners=False) of the low-resolution logits
 * followed by softmax over channels           replaces nets/hnosegxs.py:174-180
 * (the 1x1x1 conv_out commutes with the interpolation and is applied first, at low resolution,
 * with hno_pwconv_forward).  `tables` is built by hno_interp_tables_fill and uploaded by the caller.
 * ------------------------------------------------------------------------------------------ */
size_t hno_interp_tables_bytes(int D, int H, int W, int Dx, int Hx, int Wx);
int hno_interp_tables_fill(void* host_buf, size_t bytes, int D, int H, int W, int Dx, int Hx, int Wx);
/* logits_low [B][C][D][P] -> probs [B][C][Dx][Hx][Wx] (dense).  C <= 8.  activation: 0 none, 1 softmax. */
int hno_head_forward(const void* tables_host, const void* tables_dev, const float* logits_low, float* probs, int B,
                     int C, long P, int activation, void* stream);
/* Inference head: logits_low [B][C][D][P] -> labels [B][Dx][Hx][Wx] uint8 = argmax over the classes of the up-sampled
 * logits, i.e. the label map experiments/train_test.py:402-408 computes with np.argmax(probs, 1) on the host after copying
 * the probabilities back; here one byte per voxel leaves the device. */
int hno_head_argmax(const void* tables_host, const void* tables_dev, const float* logits_low, unsigned char* labels,
                    int B, int C, long P, void* stream);
/* Full-resolution head of the use_resize=False models (nets/hnosegxs.py:102-109, 150, 174-180; nets/architectures.py:286-289,
 * 345-351): no interpolation; conv_out runs through hno_pwconv_forward and these apply the output activation between the
 * planar logits [B][C][D][P] and the dense probabilities [B][C][D][H][W] (activation 0 none, 1 softmax; C <= 8), produce
 * the uint8 argmax label map (first maximum, as np.argmax at experiments/train_test.py:402-408), and run the backward
 * (padding columns of dlogits = 0; probs may be NULL for activation 0). */
int hno_head_direct_forward(const float* logits, float* probs, int B, int C, int D, int H, int W, long P, int activation,
                            void* stream);
int hno_head_direct_argmax(const float* logits, unsigned char* labels, int B, int C, int D, int H, int W, long P,
                           void* stream);
int hno_head_direct_backward(const float* dprobs, const float* probs, float* dlogits, int B, int C, int D, int H, int W,
                             long P, int activation, void* stream);
size_t hno_head_backward_workspace_bytes(const void* tables_host, int B, int C);
/* dprobs, probs [B][C][Dx][Hx][Wx] -> dlogits_low [B][C][D][P] (padding columns = 0). */
int hno_head_backward(const void* tables_host, const void* tables_dev, const float* dprobs, const float* probs,
                      float* dlogits_low, void* workspace, int B, int C, long P, int activation, void* stream);

/* ------------------------------------------------------------------------------------------
 * Losses on probabilities                     replaces nets/custom_losses.py:17-111
 *   kind 0: DiceLoss, 1: PCCLoss, 2: ExpDiceLoss (param = its exponent, :114-133; ignored otherwise).
 *   y_pred, y_true [B][C][N] fp32 (y_true one-hot floats).
 *   forward writes the scalar loss to loss[0] and per-(b,c) backward coefficients
 *   (alpha, beta, gamma) to coef[B*C*3]:   dL/dy_pred = alpha + beta*y_true + gamma*y_pred.
 * ------------------------------------------------------------------------------------------ */
size_t hno_loss_workspace_bytes(int B, int C);
int hno_loss_forward(const float* y_pred, const float* y_true, float* loss, float* coef, void* workspace, int B,
                     int C, long N, int kind, float param, void* stream);
int hno_loss_backward(const float* y_pred, const float* y_true, const float* coef, const float* grad_loss,
                      float* dy_pred, int B, int C, long N, void* stream);

/* Fused head + loss on integer labels (the library's own training step; probabilities are never
 * written): logits_low -> (interp, softmax) -> loss sums;  backward recomputes the probabilities.
 *   replaces experiments/utils.py:74-97 (to_categorical) + nets/hnosegxs.py:174-180 +
 *            nets/custom_losses.py + their autograd.  labels: uint8 [B][Dx][Hx][Wx]. */
int hno_head_loss_forward(const void* tables_host, const void* tables_dev, const float* logits_low,
                          const uint8_t* labels, float* loss, float* coef, void* workspace, int B, int C, long P,
                          int kind, float param, void* stream);
int hno_head_loss_backward(const void* tables_host, const void* tables_dev, const float* logits_low,
                           const uint8_t* labels, const float* coef, const float* grad_loss, float* dlogits_low,
                           void* workspace, int B, int C, long P, void* stream);

/* ------------------------------------------------------------------------------------------
 * Cross entropy ON PROBABILITIES             replaces torch.nn.CrossEntropyLoss() as the reference
 *   configures and calls it: experiments/run.py:105-110 (loss_name looked up in torch.nn),
 *   experiments/train_test.py:159-160 (loss_fn(model(x), one_hot)): the model's softmax OUTPUT is
 *   the "input" of log-softmax, targets are class probabilities, reduction 'mean' over B*N:
 *     loss = 1/(B N) sum_{b,v} [ (sum_c t_c) logsumexp_c(p) - sum_c t_c p_c ].
 *   y_pred [B][C][N] fp32.  Targets: exactly one of y_true ([B][C][N] fp32) and labels ([B][N]
 *   uint8 class indices, i.e. the input of experiments/utils.py:74-97 instead of its output).
 *   backward: dy_pred = grad_loss * (softmax(p)_c * sum_c' t_c' - t_c) / (B N)  (grad_loss NULL = 1).
 * ------------------------------------------------------------------------------------------ */
size_t hno_ce_loss_workspace_bytes(int B);
int hno_ce_loss_forward(const float* y_pred, const float* y_true, const uint8_t* labels, float* loss,
                        void* workspace, int B, int C, long N, void* stream);
int hno_ce_loss_backward(const float* y_pred, const float* y_true, const uint8_t* labels,
                         const float* grad_loss, float* dy_pred, int B, int C, long N, void* stream);

/* ------------------------------------------------------------------------------------------
 * Input side of the step on the device
 *   hno_to_categorical        replaces experiments/utils.py:74-97 (called at train_test.py:152, :197):
 *     labels [B][N] (label_bytes 1 = uint8, 8 = int64) -> onehot fp32 [B][C][N] (dense channel-first;
 *     the reference returns the same values as a channels-last view).  bad_count (device int, may be
 *     NULL) receives the number of labels outside [0, C) -- the reference raises IndexError on those.
 *   hno_normalize_modalities  replaces experiments/utils.py:25-71 (x_processing, run.py:52-55):
 *     data [rows][n] fp32, one modality per row: optional np.clip(clip_lo, clip_hi), mean and
 *     population std over the voxels whose CLIPPED value differs from mask_val (all voxels without
 *     a mask), out = (x - mean) / std, masked voxels = 0.  out must not alias data.
 * ------------------------------------------------------------------------------------------ */
int hno_to_categorical(const void* labels, int label_bytes, float* onehot, int* bad_count, int B, int C, long N,
                       void* stream);
size_t hno_normalize_workspace_bytes(int rows);
int hno_normalize_modalities(const float* data, float* out, void* workspace, int rows, long n, int has_mask,
                             float mask_val, int has_clip, float clip_lo, float clip_hi, void* stream);
/* the same on raw int16 intensities (the NIfTI storage type of the BraTS volumes the reference reads,
 * experiments/utils.py:260-270 -> np.asarray(data, dtype=np.float32) at :47): the batch crosses PCIe at half the bytes and
 * the exact int16 -> fp32 conversion happens in the load */
int hno_normalize_modalities_i16(const short* data, float* out, void* workspace, int rows, long n, int has_mask,
                                 float mask_val, int has_clip, float clip_lo, float clip_hi, void* stream);
/* hno_affine_resample_nn  replaces experiments/data_io/dataset.py:205-237 (apply_transform: one SimpleITK
 * ResampleImageFilter.Execute per channel with an AffineTransform, sitkNearestNeighbor, default pixel value cval, unit
 * spacing / zero origin) and :240-245 (flip_axis) for a whole batch:
 *   in / out [B][C][D][H][W], elem_bytes 1 (uint8 labels), 2 (int16 raw intensities) or 4 (float32); out != in.
 *   xform  DEVICE [B][12] fp64: per sample the 3 x 4 matrix (row-major, SimpleITK's (x, y, z) = (W, H, D) order) that
 *          maps an output index to the continuous input index -- the affine part and offset the reference passes to
 *          sitk.AffineTransform (:220-224) after transform_matrix_offset_center (:195-202).
 *   flags  DEVICE [B] (may be NULL = 0): bit 0 / 1 / 2 flip the D / H / W axis of the RESAMPLED image (the reference
 *          flips after the resampling, :172-178), bit 3 = no geometric transform for this sample (flips only).
 *   out[d][h][w] = in[nearest(M (x, y, z, 1))] with round-half-up per axis, cval (cast to the element type) outside.
 * A 2-D image (C, H, W) is the D == 1 case with an identity z row. */
int hno_affine_resample_nn(const void* in, void* out, int elem_bytes, const double* xform, const int* flags, int B,
                           int C, int D, int H, int W, double cval, void* stream);
/* Batched 2-D transpose out[n][c][r] = in[n][r][c] (1-, 2- or 4-byte elements; out != in): the layout change
 * behind running a volume with its shortest spatial axis last (the networks are equivariant under axis permutations;
 * parallel.Trainer._axis_perm).  No counterpart in the reference, which runs whatever order the reader delivers
 * (experiments/utils.py:260-270: SimpleITK's (z, y, x)). */
int hno_transpose2d(const void* in, void* out, int elem_bytes, long n, int R, int C, void* stream);

/* ------------------------------------------------------------------------------------------
 * Deep-supervision convolution                 replaces nets/architectures.py:295-311, 330-343 and
 * nets/hnosegxs.py:110-125, 154-172:  x = conv_ds(torch.cat(tensors, 1))  -- a 1x1x1 ConvNormAct over the
 * concatenation of EVERY block output -- evaluated over the list of sources without forming the concatenation:
 *   out[b][o][s] = act( bias[o] + sum_i sum_c weight[o][off_i + c] * in_i[b][c][s] ),   act: 0 none, 1 SELU.
 * in[i]: [B][ch[i]][S] fp32, 16-byte aligned, S % 4 == 0 (the planar layout guarantees it); n <= 40 sources; CO <= 8.
 * weight2 [CO][CO] (optional, with out2): the bias-free conv_out that follows the head (architectures.py:311-313; it commutes
 * with the trilinear up-sampling and is applied here, at low resolution): out2 = weight2 * out in the same pass.
 * backward: dy = gradient of out2 when weight2 is given (else of out); din[i] (may be null per source) = W_i^T d(pre),
 * dweight [CO][sum ch], dbias [CO] (may be null), dweight2 [CO][CO]; P / HW: plane pitch and valid columns per plane (gradients
 * of the padding columns are zero); y = the forward's `out`.
 * ------------------------------------------------------------------------------------------ */
int hno_dsconv_forward(const float* const* in, const int* ch, int n, const float* weight, const float* bias, float* out,
                       const float* weight2, float* out2, int B, int CO, long S, int act, void* stream);
size_t hno_dsconv_backward_workspace_bytes(int ctot, int CO, int B, long S);
int hno_dsconv_backward(const float* const* in, float* const* din, const int* ch, int n, const float* weight,
                        const float* dy, const float* y, const float* weight2, float* dweight, float* dbias,
                        float* dweight2, void* workspace, int B, int CO, long S, long P, long HW, int act, void* stream);

/* ------------------------------------------------------------------------------------------
 * Hartley multi-head attention on the retained modes       replaces nets/hartley_mha.py:136-222
 * (HartleyMultiHeadAttention._call / _call_notransform between the forward and the inverse transform).
 * Geometry: mode block (Ld, Lh, Lw) = 2 * num_modes, patch (pd, ph, pw) (1, 1, 1 without grouping), tokens
 * T = (Ld/pd)(Lh/ph)(Lw/pw) padded to Tp (multiple of 128), features per head F = channels * pd*ph*pw padded to Fp (multiple
 * of 32; a value accepted by the GEMM tiles: <= 256 or a multiple of 128).  Attention operands live in two layouts,
 *   x_tok [B*H][Tp][Fp] (token major)   and   x_chan [B*H][Fp][Tp] (feature major),   zero in the padding.
 *   hno_mha_project_*    freq_conv3d (:310-334, per-head 1x1x1 conv over the cropped block, weight [H][cd][cin], optional
 *                        bias [H][cd]) + grouping3d (:473-498) into both layouts; backward gives dz (optionally accumulated:
 *                        self-attention feeds one mode tensor to Q, K and V), dw, dbias.
 *   hno_mha_attention_*  att = act(Q^T K * scale) (:198-201, scale = 1/sqrt(F), act SELU or none), out = att V (:203);
 *                        forward keeps att and its transpose for the backward (P, PT [B*H][Tp][Tp]); tcgen05 3xTF32 GEMMs.
 *   hno_mha_output_*     ungrouping3d (:501-524) + 'oi,bidhw->bodhw' with weight_out [co][H*cd] (+ bias [co]) (:207-216).
 * ------------------------------------------------------------------------------------------ */
int hno_mha_project_forward(const float* z, const float* weight, const float* bias, float* x_tok, float* x_chan, int B,
                            int H, int cin, int cd, int Ld, int Lh, int Lw, int pd, int ph, int pw, int Tp, int Fp,
                            void* stream);
/* workspace of the two backward entry points below (per-tile partial sums of the weight gradients; cin = co for the output
 * projection, T = number of tokens).  A NULL workspace selects the slower one-CTA-per-output weight-gradient kernel. */
size_t hno_mha_wgrad_workspace_bytes(int B, int H, int cin, int cd, int T);
int hno_mha_project_backward(const float* dx_tok, const float* z, const float* weight, float* dz, float* dweight,
                             float* dbias, void* workspace, int B, int H, int cin, int cd, int Ld, int Lh, int Lw, int pd,
                             int ph, int pw, int Tp, int Fp, int accumulate_dz, void* stream);
int hno_mha_attention_forward(const float* q_tok, const float* k_tok, const float* v_chan, float* P, float* PT,
                              float* o_tok, int BH, int Tp, int Fqp, int Fvp, float scale, int activation, void* stream);
int hno_mha_attention_backward(const float* do_tok, const float* do_chan, const float* q_chan, const float* k_chan,
                               const float* v_tok, const float* P, const float* PT, float* dS, float* dST, float* dq_tok,
                               float* dk_tok, float* dv_tok, int BH, int Tp, int Fqp, int Fvp, float scale, int activation,
                               void* stream);
int hno_mha_output_forward(const float* o_tok, const float* weight_out, const float* bias, float* y, int B, int H, int co,
                           int cd, int Ld, int Lh, int Lw, int pd, int ph, int pw, int Tp, int Fp, void* stream);
int hno_mha_output_backward(const float* dy, const float* o_tok, const float* weight_out, float* do_tok, float* do_chan,
                            float* dweight_out, float* dbias, void* workspace, int B, int H, int co, int cd, int Ld, int Lh,
                            int Lw, int pd, int ph, int pw, int Tp, int Fp, void* stream);

/* ------------------------------------------------------------------------------------------
 * Fused Adamax step on a flat parameter vector (torch.optim.Adamax semantics, the optimizer of
 * experiments/config_files/config_hnoseg_xs.ini:53-55).  step is 1-based.
 * ------------------------------------------------------------------------------------------ */
int hno_adamax_step(float* param, const float* grad, float* exp_avg, float* exp_inf, long n, float lr, float beta1,
                    float beta2, float eps, float weight_decay, int step, float grad_scale, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* HNO_B200_H_ */
