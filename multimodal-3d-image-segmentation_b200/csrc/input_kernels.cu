// Input side of the training step on the device, sm_100a: the two array transforms that sit between the
// loader and the network in the reference.
//
//   to_categorical          experiments/utils.py:74-97   integer labels (B,1,D,H,W) -> one-hot fp32 (B,C,D,H,W)
//                           (called every step at train_test.py:152 / :197 on the device tensor)
//   normalize_modalities    experiments/utils.py:25-71   per modality: optional clip, mean / std over the voxels that
//                           differ from mask_val (after clipping), (x - mean) / std, masked voxels -> 0
//                           (x_processing of the loader, experiments/run.py:52-55: mask_val = 0 by default)
//
// Both are HBM-bound streaming passes: labels are read once as uchar4 / int64 and the C one-hot planes written with
// 16-byte stores; the normalisation is one moments pass (count, sum, sum of squares in fp64 block partials) and one apply
// pass, i.e. 2 reads + 1 write of the volume (the numpy original makes ~8 passes and a masked-array copy).
#include "common.cuh"
#include "hno_b200.h"

#include <stdint.h>

namespace hno {

// ------------------------------------------------------------------------------------------ to_categorical
template <int V>
__global__ void __launch_bounds__(256) k_to_categorical_u8(const uint8_t* __restrict__ lab, float* __restrict__ out,
                                                           int* __restrict__ bad, int C, long N) {
  const long b = blockIdx.y;
  const uint8_t* l = lab + b * N;
  float* o = out + b * C * N;
  int nbad = 0;
  for (long i = blockIdx.x * 256L + threadIdx.x; i < N / V; i += (long)gridDim.x * 256L) {
    if (V == 4) {
      const uchar4 q = __ldg(reinterpret_cast<const uchar4*>(l) + i);
      nbad += (q.x >= C) + (q.y >= C) + (q.z >= C) + (q.w >= C);
      for (int c = 0; c < C; ++c)
        reinterpret_cast<float4*>(o + c * N)[i] =
            make_float4(q.x == c ? 1.f : 0.f, q.y == c ? 1.f : 0.f, q.z == c ? 1.f : 0.f, q.w == c ? 1.f : 0.f);
    } else {
      const int q = (int)__ldg(l + i);
      nbad += q >= C;
      for (int c = 0; c < C; ++c) o[c * N + i] = q == c ? 1.f : 0.f;
    }
  }
  if (bad && nbad) atomicAdd(bad, nbad);
}

__global__ void __launch_bounds__(256) k_to_categorical_i64(const long long* __restrict__ lab, float* __restrict__ out,
                                                            int* __restrict__ bad, int C, long N) {
  const long b = blockIdx.y;
  const long long* l = lab + b * N;
  float* o = out + b * C * N;
  int nbad = 0;
  for (long i = blockIdx.x * 256L + threadIdx.x; i < N; i += (long)gridDim.x * 256L) {
    const long long q = __ldg(l + i);
    nbad += q < 0 || q >= C;
    for (int c = 0; c < C; ++c) o[c * N + i] = q == c ? 1.f : 0.f;
  }
  if (bad && nbad) atomicAdd(bad, nbad);
}

int to_categorical(const void* labels, int label_bytes, float* onehot, int* bad_count, int B, int C, long N,
                   cudaStream_t st) {
  HNO_CHECK(labels && onehot, "to_categorical: null pointer");
  HNO_CHECK(label_bytes == 1 || label_bytes == 8, "to_categorical: labels must be uint8 or int64");
  HNO_CHECK(B >= 1 && B <= 65535 && C >= 1 && N >= 1, "to_categorical: bad sizes");
  HNO_CHECK(label_bytes == 8 || C <= 256, "to_categorical: uint8 labels cannot address %d classes", C);
  if (bad_count) HNO_CUDA(cudaMemsetAsync(bad_count, 0, sizeof(int), st));
  const bool v4 = label_bytes == 1 && N % 4 == 0 && (reinterpret_cast<uintptr_t>(labels) & 3) == 0 &&
                  (reinterpret_cast<uintptr_t>(onehot) & 15) == 0;
  const long items = v4 ? N / 4 : N;
  dim3 grid((unsigned)((items + 255) / 256 < 1184 ? (items + 255) / 256 : 1184), B);
  if (label_bytes == 8)
    k_to_categorical_i64<<<grid, 256, 0, st>>>(static_cast<const long long*>(labels), onehot, bad_count, C, N);
  else if (v4)
    k_to_categorical_u8<4><<<grid, 256, 0, st>>>(static_cast<const uint8_t*>(labels), onehot, bad_count, C, N);
  else
    k_to_categorical_u8<1><<<grid, 256, 0, st>>>(static_cast<const uint8_t*>(labels), onehot, bad_count, C, N);
  HNO_LAUNCH_CHECK();
  return 0;
}

// ------------------------------------------------------------------------------------------ normalize_modalities
constexpr int kNormChunks = 296;

struct NormArgs {
  int has_mask, has_clip;
  float mask_val, clip_lo, clip_hi;
};

__device__ __forceinline__ float norm_clip(float v, const NormArgs& a) {
  return a.has_clip ? fminf(fmaxf(v, a.clip_lo), a.clip_hi) : v;  // np.clip
}

// Storage types of the raw modalities: fp32 (what the reference's loader hands over) or int16 (the NIfTI storage type of
// BraTS volumes: shipping it halves the host -> device bytes of a batch; the conversion is exact)
__device__ __forceinline__ float4 norm_ld4(const float* p, long i) { return __ldg(reinterpret_cast<const float4*>(p) + i); }
__device__ __forceinline__ float4 norm_ld4(const int16_t* p, long i) {
  const short4 q = __ldg(reinterpret_cast<const short4*>(p) + i);
  return make_float4((float)q.x, (float)q.y, (float)q.z, (float)q.w);
}
__device__ __forceinline__ float norm_ld1(const float* p, long i) { return __ldg(p + i); }
__device__ __forceinline__ float norm_ld1(const int16_t* p, long i) { return (float)__ldg(p + i); }

// grid (chunks, rows): partials [rows][chunks][3] = (count, sum, sum of squares) of the unmasked, clipped voxels
template <int V, typename T>
__global__ void __launch_bounds__(256) k_norm_moments(const T* __restrict__ x, double* __restrict__ partials, long n,
                                                      NormArgs a) {
  __shared__ double sred[8][3];
  const T* p = x + (long)blockIdx.y * n;
  float s = 0.f, ss = 0.f;
  double cnt = 0.0, sd = 0.0, ssd = 0.0;
  int c32 = 0, run = 0;
  auto add = [&](float v) {
    v = norm_clip(v, a);
    if (!(a.has_mask && v == a.mask_val)) {
      s += v;
      ss = fmaf(v, v, ss);
      ++c32;
    }
  };
  for (long i = blockIdx.x * 256L + threadIdx.x; i < n / V; i += (long)gridDim.x * 256L) {
    if (V == 4) {
      const float4 q = norm_ld4(p, i);
      add(q.x), add(q.y), add(q.z), add(q.w);
    } else {
      add(norm_ld1(p, i));
    }
    if (++run == 32 / V) {  // bounded fp32 run length (raw intensities reach 1e3 - 1e4), then into fp64
      sd += (double)s, ssd += (double)ss, cnt += (double)c32;
      s = ss = 0.f, c32 = 0, run = 0;
    }
  }
  double m[3] = {cnt + (double)c32, sd + (double)s, ssd + (double)ss};
#pragma unroll
  for (int k = 0; k < 3; ++k) m[k] = warp_sum_d(m[k]);
  if ((threadIdx.x & 31) == 0)
#pragma unroll
    for (int k = 0; k < 3; ++k) sred[threadIdx.x >> 5][k] = m[k];
  __syncthreads();
  if (threadIdx.x < 3) {
    double t = 0.0;
    for (int w = 0; w < 8; ++w) t += sred[w][threadIdx.x];
    partials[((long)blockIdx.y * gridDim.x + blockIdx.x) * 3 + threadIdx.x] = t;
  }
}

// one warp per row: mean and (population, ddof = 0) standard deviation as fp32, like numpy's float32 result
__global__ void __launch_bounds__(32) k_norm_finalize(const double* __restrict__ partials, int chunks,
                                                      float* __restrict__ stats) {
  const int row = blockIdx.x;
  double m[3] = {0.0, 0.0, 0.0};
  for (int ch = threadIdx.x; ch < chunks; ch += 32)
#pragma unroll
    for (int k = 0; k < 3; ++k) m[k] += partials[((long)row * chunks + ch) * 3 + k];
#pragma unroll
  for (int k = 0; k < 3; ++k) m[k] = warp_sum_d(m[k]);
  if (threadIdx.x == 0) {
    const double mean = m[0] > 0.0 ? m[1] / m[0] : 0.0;
    const double var = m[0] > 0.0 ? fmax(m[2] / m[0] - mean * mean, 0.0) : 0.0;
    stats[row * 2 + 0] = (float)mean;
    stats[row * 2 + 1] = (float)sqrt(var);
  }
}

template <int V, typename T>
__global__ void __launch_bounds__(256) k_norm_apply(const T* __restrict__ x, float* __restrict__ out,
                                                    const float* __restrict__ stats, long n, NormArgs a) {
  const T* p = x + (long)blockIdx.y * n;
  float* o = out + (long)blockIdx.y * n;
  const float mean = stats[blockIdx.y * 2 + 0], sd = stats[blockIdx.y * 2 + 1];
  auto f = [&](float v) {
    v = norm_clip(v, a);
    return (a.has_mask && v == a.mask_val) ? 0.f : (v - mean) / sd;  // true division, as numpy
  };
  for (long i = blockIdx.x * 256L + threadIdx.x; i < n / V; i += (long)gridDim.x * 256L) {
    if (V == 4) {
      const float4 q = norm_ld4(p, i);
      reinterpret_cast<float4*>(o)[i] = make_float4(f(q.x), f(q.y), f(q.z), f(q.w));
    } else {
      o[i] = f(norm_ld1(p, i));
    }
  }
}

size_t normalize_workspace_bytes(int rows) {
  return (size_t)rows * kNormChunks * 3 * sizeof(double) + (size_t)rows * 2 * sizeof(float) + 256;
}

template <typename T>
static int normalize_modalities_t(const T* data, float* out, void* ws, int rows, long n, int has_mask, float mask_val,
                                  int has_clip, float clip_lo, float clip_hi, cudaStream_t st) {
  HNO_CHECK(data && out && ws, "normalize_modalities: null pointer");
  HNO_CHECK(rows >= 1 && rows <= 65535 && n >= 1, "normalize_modalities: bad sizes");
  HNO_CHECK(!has_clip || clip_lo <= clip_hi, "normalize_modalities: clip_val must be (min, max)");
  double* partials = reinterpret_cast<double*>(ws);
  float* stats = reinterpret_cast<float*>(partials + (size_t)rows * kNormChunks * 3);
  const NormArgs a{has_mask, has_clip, mask_val, clip_lo, clip_hi};
  const bool v4 = n % 4 == 0 && (reinterpret_cast<uintptr_t>(data) % (4 * sizeof(T))) == 0 &&
                  (reinterpret_cast<uintptr_t>(out) & 15) == 0;
  const long items = v4 ? n / 4 : n;
  const int chunks = (int)((items + 255) / 256 < kNormChunks ? (items + 255) / 256 : kNormChunks);
  dim3 g1(chunks, rows);
  if (v4) k_norm_moments<4, T><<<g1, 256, 0, st>>>(data, partials, n, a);
  else k_norm_moments<1, T><<<g1, 256, 0, st>>>(data, partials, n, a);
  HNO_LAUNCH_CHECK();
  k_norm_finalize<<<rows, 32, 0, st>>>(partials, chunks, stats);
  HNO_LAUNCH_CHECK();
  dim3 g2((unsigned)((items + 255) / 256 < 1184 ? (items + 255) / 256 : 1184), rows);
  if (v4) k_norm_apply<4, T><<<g2, 256, 0, st>>>(data, out, stats, n, a);
  else k_norm_apply<1, T><<<g2, 256, 0, st>>>(data, out, stats, n, a);
  HNO_LAUNCH_CHECK();
  return 0;
}

int normalize_modalities(const void* data, int elem_bytes, float* out, void* ws, int rows, long n, int has_mask,
                         float mask_val, int has_clip, float clip_lo, float clip_hi, cudaStream_t st) {
  if (elem_bytes == 4)
    return normalize_modalities_t(static_cast<const float*>(data), out, ws, rows, n, has_mask, mask_val, has_clip,
                                  clip_lo, clip_hi, st);
  HNO_CHECK(elem_bytes == 2, "normalize_modalities: element size must be 4 (float32) or 2 (int16), got %d", elem_bytes);
  return normalize_modalities_t(static_cast<const int16_t*>(data), out, ws, rows, n, has_mask, mask_val, has_clip,
                                clip_lo, clip_hi, st);
}

// ------------------------------------------------------------------------------------------ affine augmentation
// experiments/data_io/dataset.py:205-237 (apply_transform) + :240-245 (flip_axis) for a whole batch on the device.  The
// reference hands every channel of every sample to SimpleITK's ResampleImageFilter (AffineTransform, nearest neighbour,
// default pixel value cval, unit spacing, zero origin) inside the loader workers and flips the result with numpy views.
// Here one thread owns one output voxel: undo the flips, map the output index (x, y, z) = (w, h, d) through the sample's
// 3 x 4 matrix to the continuous input index, round half up (ITK's ConvertContinuousIndexToNearestIndex), and copy all
// channels of that input voxel -- or cval when the index leaves the buffer (ITK: continuous index in [-0.5, size - 0.5)).
// The coordinates are fp64 with an explicit operation order (no FMA contraction), so the CPU oracle reproduces every
// rounding decision.  A gather of 1 / 2 / 4-byte elements: HBM bound, one read + one write of the batch.
template <typename T, int V>
struct alignas(sizeof(T) * V) AugPack {
  T v[V];
};

// V consecutive output voxels (linear index) per thread: the gathers stay scalar, the store is one sizeof(T) * V word
// (4 B for uint8, 8 B for int16, 16 B for fp32 at V = 4), which is what keeps the write side coalesced for the narrow types.
template <typename T, int V>
__global__ void __launch_bounds__(256) k_affine_resample_nn(const T* __restrict__ in, T* __restrict__ out,
                                                            const double* __restrict__ xform,
                                                            const int* __restrict__ flags, int C, int D, int H, int W,
                                                            T cval) {
  const long N = (long)D * H * W;
  const int b = blockIdx.y;
  const T* ib = in + (long)b * C * N;
  T* ob = out + (long)b * C * N;
  const int fl = flags ? flags[b] : 0;
  double m[12];
#pragma unroll
  for (int k = 0; k < 12; ++k) m[k] = xform[b * 12 + k];
  for (long g = blockIdx.x * 256L + threadIdx.x; g < N / V; g += (long)gridDim.x * 256L) {
    const long i = g * V;
    const unsigned iu = (unsigned)i;  // N < 2^31 (checked by the launcher): 32-bit divisions
    const unsigned t = iu / (unsigned)W;
    int w = (int)(iu - t * (unsigned)W);
    int d = (int)(t / (unsigned)H);
    int h = (int)(t - (unsigned)d * (unsigned)H);
    long src[V];
#pragma unroll
    for (int e = 0; e < V; ++e) {
      const int px = (fl & 4) ? W - 1 - w : w;
      const int py = (fl & 2) ? H - 1 - h : h;
      const int pz = (fl & 1) ? D - 1 - d : d;
      if (fl & 8) {  // no geometric transform was drawn for this sample: flips only
        src[e] = ((long)pz * H + py) * W + px;
      } else {
        const double x = (double)px, y = (double)py, z = (double)pz;
        double c[3];
#pragma unroll
        for (int r = 0; r < 3; ++r) {
          double s = __dadd_rn(__dmul_rn(m[4 * r + 0], x), __dmul_rn(m[4 * r + 1], y));
          s = __dadd_rn(s, __dmul_rn(m[4 * r + 2], z));
          c[r] = floor(__dadd_rn(__dadd_rn(s, m[4 * r + 3]), 0.5));
        }
        const bool inside = c[0] >= 0.0 && c[0] < (double)W && c[1] >= 0.0 && c[1] < (double)H && c[2] >= 0.0 &&
                            c[2] < (double)D;
        src[e] = inside ? ((long)c[2] * H + (long)c[1]) * W + (long)c[0] : -1;
      }
      if (++w == W) {
        w = 0;
        if (++h == H) h = 0, ++d;
      }
    }
    for (int ch = 0; ch < C; ++ch) {
      AugPack<T, V> pk;
#pragma unroll
      for (int e = 0; e < V; ++e) pk.v[e] = src[e] >= 0 ? __ldg(ib + ch * N + src[e]) : cval;
      *reinterpret_cast<AugPack<T, V>*>(ob + ch * N + i) = pk;
    }
  }
}

template <typename T>
static void affine_resample_launch(const void* in, void* out, const double* xform, const int* flags, int B, int C, int D,
                                   int H, int W, T cval, cudaStream_t st) {
  const long N = (long)D * H * W;
  const bool v4 = N % 4 == 0 && reinterpret_cast<uintptr_t>(out) % (4 * sizeof(T)) == 0;
  const long items = v4 ? N / 4 : N;
  dim3 grid((unsigned)((items + 255) / 256 < 8L * sm_count() ? (items + 255) / 256 : 8L * sm_count()), B);
  if (v4)
    k_affine_resample_nn<T, 4><<<grid, 256, 0, st>>>(static_cast<const T*>(in), static_cast<T*>(out), xform, flags, C, D,
                                                     H, W, cval);
  else
    k_affine_resample_nn<T, 1><<<grid, 256, 0, st>>>(static_cast<const T*>(in), static_cast<T*>(out), xform, flags, C, D,
                                                     H, W, cval);
}

int affine_resample_nn(const void* in, void* out, int elem_bytes, const double* xform, const int* flags, int B, int C,
                       int D, int H, int W, double cval, cudaStream_t st) {
  HNO_CHECK(in && out && xform, "affine_resample_nn: null pointer");
  HNO_CHECK(in != out, "affine_resample_nn: the resampling is a gather and cannot run in place");
  HNO_CHECK(elem_bytes == 1 || elem_bytes == 2 || elem_bytes == 4,
            "affine_resample_nn: elements must be uint8 (1), int16 (2) or float32 (4)");
  HNO_CHECK(B >= 1 && B <= 65535 && C >= 1 && D >= 1 && H >= 1 && W >= 1, "affine_resample_nn: bad sizes");
  HNO_CHECK((long)D * H * W < (1L << 31), "affine_resample_nn: volume too large for 32-bit voxel indices");
  if (elem_bytes == 1) affine_resample_launch<uint8_t>(in, out, xform, flags, B, C, D, H, W, (uint8_t)cval, st);
  else if (elem_bytes == 2) affine_resample_launch<int16_t>(in, out, xform, flags, B, C, D, H, W, (int16_t)cval, st);
  else affine_resample_launch<float>(in, out, xform, flags, B, C, D, H, W, (float)cval, st);
  HNO_LAUNCH_CHECK();
  return 0;
}

// ------------------------------------------------------------------------------------------ axis permutation
// out[n][c][r] = in[n][r][c]: the batched 2-D transpose behind "shortest spatial axis last" (parallel.Trainer._axis_perm):
// (D, H, W) -> (H, W, D) is R = D, C = H * W per (sample, channel); (D, H, W) -> (D, W, H) is R = H, C = W per (sample, channel,
// slice).  32 x 32 tiles through padded shared memory, both sides coalesced.
template <typename T>
__global__ void __launch_bounds__(256) k_transpose2d(const T* __restrict__ in, T* __restrict__ out, int R, int C) {
  __shared__ T tile[32][33];
  const long base = (long)blockIdx.z * R * C;
  const int c0 = blockIdx.x * 32, r0 = blockIdx.y * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int r = r0 + ty + 8 * k, c = c0 + tx;
    if (r < R && c < C) tile[ty + 8 * k][tx] = __ldg(in + base + (long)r * C + c);
  }
  __syncthreads();
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int c = c0 + ty + 8 * k, r = r0 + tx;
    if (r < R && c < C) out[base + (long)c * R + r] = tile[tx][ty + 8 * k];
  }
}

int transpose2d(const void* in, void* out, int elem_bytes, long n, int R, int C, cudaStream_t st) {
  HNO_CHECK(in && out && in != out, "transpose2d: null or aliased pointers");
  HNO_CHECK(elem_bytes == 1 || elem_bytes == 2 || elem_bytes == 4, "transpose2d: elements must be 1, 2 or 4 bytes");
  HNO_CHECK(n >= 1 && R >= 1 && C >= 1 && (R + 31) / 32 <= 65535, "transpose2d: bad sizes");
  for (long n0 = 0; n0 < n; n0 += 65535) {  // gridDim.z limit
    const long nn = n - n0 < 65535 ? n - n0 : 65535;
    const long off = n0 * R * C;
    dim3 grid((C + 31) / 32, (R + 31) / 32, (unsigned)nn);
    if (elem_bytes == 1)
      k_transpose2d<uint8_t><<<grid, 256, 0, st>>>(static_cast<const uint8_t*>(in) + off, static_cast<uint8_t*>(out) + off, R, C);
    else if (elem_bytes == 2)
      k_transpose2d<int16_t><<<grid, 256, 0, st>>>(static_cast<const int16_t*>(in) + off, static_cast<int16_t*>(out) + off, R, C);
    else
      k_transpose2d<float><<<grid, 256, 0, st>>>(static_cast<const float*>(in) + off, static_cast<float*>(out) + off, R, C);
    HNO_LAUNCH_CHECK();
  }
  return 0;
}

}  // namespace hno
