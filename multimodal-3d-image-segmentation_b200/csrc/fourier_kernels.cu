// Mode-domain step of the Fourier spectral layer with SHARED complex weights (FNOSeg, BASELINE config 3), sm_100a.
//
// Replaces, as three launches, the chain the reference runs between rfftn and irfftn (nets/fourier_operator.py:155,
// 165-209: corner slicing, complex einsum with w_real + i w_imag, zero-padded re-assembly) in the form this package
// evaluates it: the layer input arrives as Hartley coefficients z on the symmetric mode set S (hno_dht3_forward), and
//   re(k) = (z[k] + z[N - k]) / 2,   im(k) = (z[N - k] - z[k]) / 2            (k in the rfft half-grid K)
//   a + i b = (w_real + i w_imag) (re + i im)                                 (contraction over the input channels)
//   hp[k] += c_k (a - b) / 2,   hp[N - k] += c_k (a + b) / 2                  (c_k = 1 on the k_w = 0 plane, else 2)
// gives the Hartley coefficients hp whose un-normalised inverse transform (hno_dht3_adjoint) is the layer output.
// Round 1 ran this as index_select x 2, four elementwise kernels, two pointwise-conv launches, zeros and index_add x 2.
//
// An entry of S receives at most TWO contributions (from k through lin_k and from its mirror image through lin_n, or
// both from the same self-conjugate k), added with atomics onto zeros: a two-term fp32 sum is order independent, so the
// results are deterministic.
#include "common.cuh"
#include "wgrad.cuh"

namespace hno {

constexpr int kFmOG = 4;        // output (forward) / input (backward) channels per thread
constexpr int kFmThreads = 128;

// grid (ceil(MK / 128), co / 4, B).  z [B][ci][MS], hp [B][co][MS] (zeroed by the launcher)
__global__ void __launch_bounds__(kFmThreads) k_fmix_fwd(const float* __restrict__ z, const float* __restrict__ wr,
                                                         const float* __restrict__ wi, const int* __restrict__ lin_k,
                                                         const int* __restrict__ lin_n, const float* __restrict__ ck,
                                                         float* __restrict__ hp, int ci, int co, int MK, long MS) {
  extern __shared__ float sw[];  // [2][4][ci]: w_real rows, then w_imag rows of this output group
  const int og = blockIdx.y, b = blockIdx.z;
  for (int idx = threadIdx.x; idx < kFmOG * ci; idx += kFmThreads) {
    sw[idx] = __ldg(wr + (long)og * kFmOG * ci + idx);
    sw[kFmOG * ci + idx] = __ldg(wi + (long)og * kFmOG * ci + idx);
  }
  __syncthreads();
  const int j = blockIdx.x * kFmThreads + threadIdx.x;
  if (j >= MK) return;
  const int ik = __ldg(lin_k + j), in = __ldg(lin_n + j);
  const float* zb = z + (long)b * ci * MS;
  float a[kFmOG], bb[kFmOG];
#pragma unroll
  for (int o = 0; o < kFmOG; ++o) a[o] = bb[o] = 0.f;
#pragma unroll 4
  for (int i = 0; i < ci; ++i) {
    const float hk = __ldg(zb + (long)i * MS + ik), hn = __ldg(zb + (long)i * MS + in);
    const float re = 0.5f * (hk + hn), im = 0.5f * (hn - hk);
#pragma unroll
    for (int o = 0; o < kFmOG; ++o) {
      const float r = sw[o * ci + i], q = sw[(kFmOG + o) * ci + i];
      a[o] = fmaf(r, re, fmaf(-q, im, a[o]));
      bb[o] = fmaf(q, re, fmaf(r, im, bb[o]));
    }
  }
  const float c = 0.5f * __ldg(ck + j);
  float* hb = hp + ((long)b * co + og * kFmOG) * MS;
#pragma unroll
  for (int o = 0; o < kFmOG; ++o) {
    atomicAdd(hb + (long)o * MS + ik, c * (a[o] - bb[o]));
    atomicAdd(hb + (long)o * MS + in, c * (a[o] + bb[o]));
  }
}

// d(a), d(b) of mode j and output o from the gradient of hp
__device__ __forceinline__ void fmix_dab(const float* __restrict__ dhb, long MS, int o, int ik, int in, float c, float& da,
                                         float& db) {
  const float gk = __ldg(dhb + (long)o * MS + ik), gn = __ldg(dhb + (long)o * MS + in);
  da = c * (gk + gn);
  db = c * (gn - gk);
}

// grid (ceil(MK / 128), ci / 4, B).  dz [B][ci][MS] (zeroed by the launcher)
__global__ void __launch_bounds__(kFmThreads) k_fmix_bwd_x(const float* __restrict__ dhp, const float* __restrict__ wr,
                                                           const float* __restrict__ wi, const int* __restrict__ lin_k,
                                                           const int* __restrict__ lin_n, const float* __restrict__ ck,
                                                           float* __restrict__ dz, int ci, int co, int MK, long MS) {
  extern __shared__ float sw[];  // [2][co][4]: w_real / w_imag columns of this input group
  const int ig = blockIdx.y, b = blockIdx.z;
  for (int idx = threadIdx.x; idx < co * kFmOG; idx += kFmThreads) {
    const int o = idx / kFmOG, i = idx - o * kFmOG;
    sw[idx] = __ldg(wr + (long)o * ci + ig * kFmOG + i);
    sw[co * kFmOG + idx] = __ldg(wi + (long)o * ci + ig * kFmOG + i);
  }
  __syncthreads();
  const int j = blockIdx.x * kFmThreads + threadIdx.x;
  if (j >= MK) return;
  const int ik = __ldg(lin_k + j), in = __ldg(lin_n + j);
  const float c = 0.5f * __ldg(ck + j);
  const float* dhb = dhp + (long)b * co * MS;
  float dre[kFmOG], dim[kFmOG];
#pragma unroll
  for (int i = 0; i < kFmOG; ++i) dre[i] = dim[i] = 0.f;
#pragma unroll 4
  for (int o = 0; o < co; ++o) {
    float da, db;
    fmix_dab(dhb, MS, o, ik, in, c, da, db);
#pragma unroll
    for (int i = 0; i < kFmOG; ++i) {
      const float r = sw[o * kFmOG + i], q = sw[(co + o) * kFmOG + i];
      dre[i] = fmaf(r, da, fmaf(q, db, dre[i]));
      dim[i] = fmaf(-q, da, fmaf(r, db, dim[i]));
    }
  }
  float* zb = dz + ((long)b * ci + ig * kFmOG) * MS;
#pragma unroll
  for (int i = 0; i < kFmOG; ++i) {
    atomicAdd(zb + (long)i * MS + ik, 0.5f * (dre[i] - dim[i]));
    atomicAdd(zb + (long)i * MS + in, 0.5f * (dre[i] + dim[i]));
  }
}

// dW_real[o][i] = sum d(a)[o] re[i] + d(b)[o] im[i],  dW_imag[o][i] = sum d(b)[o] re[i] - d(a)[o] im[i]  over (b, j).
// grid (ceil(MK / 128), B): a CTA stages d(a), d(b) [co][128] and re, im [ci][128] of its modes in shared memory, every thread
// then owns output pairs (o, i); one partial row [2 co ci] per CTA, summed in fp64 by k_reduce_partials.
constexpr int kFmTile = 128, kFmPitch = kFmTile + 4;
__global__ void __launch_bounds__(256) k_fmix_bwd_w(const float* __restrict__ dhp, const float* __restrict__ z,
                                                    const int* __restrict__ lin_k, const int* __restrict__ lin_n,
                                                    const float* __restrict__ ck, float* __restrict__ partials, int ci, int co,
                                                    int MK, long MS) {
  extern __shared__ float sm[];
  float* sda = sm;                       // [co][pitch]
  float* sdb = sda + co * kFmPitch;      // [co][pitch]
  float* sre = sdb + co * kFmPitch;      // [ci][pitch]
  float* sim = sre + ci * kFmPitch;      // [ci][pitch]
  const int b = blockIdx.y;
  const int j0 = blockIdx.x * kFmTile;
  {
    const int m = threadIdx.x & (kFmTile - 1), part = threadIdx.x >> 7;  // two threads per mode: half of the rows each
    const int j = j0 + m;
    const bool ok = j < MK;
    const int ik = ok ? __ldg(lin_k + j) : 0, in = ok ? __ldg(lin_n + j) : 0;
    const float c = ok ? 0.5f * __ldg(ck + j) : 0.f;
    const float* dhb = dhp + (long)b * co * MS;
    const float* zb = z + (long)b * ci * MS;
    for (int o = part; o < co; o += 2) {
      float da = 0.f, db = 0.f;
      if (ok) fmix_dab(dhb, MS, o, ik, in, c, da, db);
      sda[o * kFmPitch + m] = da;
      sdb[o * kFmPitch + m] = db;
    }
    for (int i = part; i < ci; i += 2) {
      const float hk = ok ? __ldg(zb + (long)i * MS + ik) : 0.f, hn = ok ? __ldg(zb + (long)i * MS + in) : 0.f;
      sre[i * kFmPitch + m] = 0.5f * (hk + hn);
      sim[i * kFmPitch + m] = 0.5f * (hn - hk);
    }
  }
  __syncthreads();
  float* prow = partials + ((long)blockIdx.y * gridDim.x + blockIdx.x) * (2L * co * ci);
  for (int idx = threadIdx.x; idx < co * ci; idx += 256) {
    const int o = idx / ci, i = idx - o * ci;
    const float4* pa = reinterpret_cast<const float4*>(sda + o * kFmPitch);
    const float4* pb = reinterpret_cast<const float4*>(sdb + o * kFmPitch);
    const float4* pr = reinterpret_cast<const float4*>(sre + i * kFmPitch);
    const float4* pi = reinterpret_cast<const float4*>(sim + i * kFmPitch);
    float gr = 0.f, gi = 0.f;
#pragma unroll 4
    for (int v = 0; v < kFmTile / 4; ++v) {
      const float4 a = pa[v], bq = pb[v], r = pr[v], q = pi[v];
      gr = fmaf(a.x, r.x, fmaf(bq.x, q.x, gr));
      gr = fmaf(a.y, r.y, fmaf(bq.y, q.y, gr));
      gr = fmaf(a.z, r.z, fmaf(bq.z, q.z, gr));
      gr = fmaf(a.w, r.w, fmaf(bq.w, q.w, gr));
      gi = fmaf(bq.x, r.x, fmaf(-a.x, q.x, gi));
      gi = fmaf(bq.y, r.y, fmaf(-a.y, q.y, gi));
      gi = fmaf(bq.z, r.z, fmaf(-a.z, q.z, gi));
      gi = fmaf(bq.w, r.w, fmaf(-a.w, q.w, gi));
    }
    prow[idx] = gr;
    prow[co * ci + idx] = gi;
  }
}

// ---- per-mode ('individual') complex weights w[o][i][j], j = position in the half-grid K (config_fno.ini, reference
//      fourier_operator.py:165-187): same three steps, the weights are read from global memory (coalesced along j) and the
//      weight gradient has no reduction over the modes.
__global__ void __launch_bounds__(kFmThreads) k_fmix_fwd_ind(const float* __restrict__ z, const float* __restrict__ wr,
                                                             const float* __restrict__ wi, const int* __restrict__ lin_k,
                                                             const int* __restrict__ lin_n, const float* __restrict__ ck,
                                                             float* __restrict__ hp, int ci, int co, int MK, long MS) {
  const int og = blockIdx.y, b = blockIdx.z;
  const int j = blockIdx.x * kFmThreads + threadIdx.x;
  if (j >= MK) return;
  const int ik = __ldg(lin_k + j), in = __ldg(lin_n + j);
  const float* zb = z + (long)b * ci * MS;
  const float* wrj = wr + (long)og * kFmOG * ci * MK + j;
  const float* wij = wi + (long)og * kFmOG * ci * MK + j;
  float a[kFmOG], bb[kFmOG];
#pragma unroll
  for (int o = 0; o < kFmOG; ++o) a[o] = bb[o] = 0.f;
#pragma unroll 2
  for (int i = 0; i < ci; ++i) {
    const float hk = __ldg(zb + (long)i * MS + ik), hn = __ldg(zb + (long)i * MS + in);
    const float re = 0.5f * (hk + hn), im = 0.5f * (hn - hk);
#pragma unroll
    for (int o = 0; o < kFmOG; ++o) {
      const float r = __ldg(wrj + ((long)o * ci + i) * MK), q = __ldg(wij + ((long)o * ci + i) * MK);
      a[o] = fmaf(r, re, fmaf(-q, im, a[o]));
      bb[o] = fmaf(q, re, fmaf(r, im, bb[o]));
    }
  }
  const float c = 0.5f * __ldg(ck + j);
  float* hb = hp + ((long)b * co + og * kFmOG) * MS;
#pragma unroll
  for (int o = 0; o < kFmOG; ++o) {
    atomicAdd(hb + (long)o * MS + ik, c * (a[o] - bb[o]));
    atomicAdd(hb + (long)o * MS + in, c * (a[o] + bb[o]));
  }
}

__global__ void __launch_bounds__(kFmThreads) k_fmix_bwd_x_ind(const float* __restrict__ dhp, const float* __restrict__ wr,
                                                               const float* __restrict__ wi, const int* __restrict__ lin_k,
                                                               const int* __restrict__ lin_n, const float* __restrict__ ck,
                                                               float* __restrict__ dz, int ci, int co, int MK, long MS) {
  const int ig = blockIdx.y, b = blockIdx.z;
  const int j = blockIdx.x * kFmThreads + threadIdx.x;
  if (j >= MK) return;
  const int ik = __ldg(lin_k + j), in = __ldg(lin_n + j);
  const float c = 0.5f * __ldg(ck + j);
  const float* dhb = dhp + (long)b * co * MS;
  const float* wrj = wr + (long)ig * kFmOG * MK + j;
  const float* wij = wi + (long)ig * kFmOG * MK + j;
  float dre[kFmOG], dim[kFmOG];
#pragma unroll
  for (int i = 0; i < kFmOG; ++i) dre[i] = dim[i] = 0.f;
#pragma unroll 2
  for (int o = 0; o < co; ++o) {
    float da, db;
    fmix_dab(dhb, MS, o, ik, in, c, da, db);
#pragma unroll
    for (int i = 0; i < kFmOG; ++i) {
      const float r = __ldg(wrj + ((long)o * ci + i) * MK), q = __ldg(wij + ((long)o * ci + i) * MK);
      dre[i] = fmaf(r, da, fmaf(q, db, dre[i]));
      dim[i] = fmaf(-q, da, fmaf(r, db, dim[i]));
    }
  }
  float* zb = dz + ((long)b * ci + ig * kFmOG) * MS;
#pragma unroll
  for (int i = 0; i < kFmOG; ++i) {
    atomicAdd(zb + (long)i * MS + ik, 0.5f * (dre[i] - dim[i]));
    atomicAdd(zb + (long)i * MS + in, 0.5f * (dre[i] + dim[i]));
  }
}

// grid (ceil(MK / 128), co / 4, ci / 4): a thread owns 4 x 4 (o, i) pairs of its mode and sums over the batch
__global__ void __launch_bounds__(kFmThreads) k_fmix_bwd_w_ind(const float* __restrict__ dhp, const float* __restrict__ z,
                                                               const int* __restrict__ lin_k, const int* __restrict__ lin_n,
                                                               const float* __restrict__ ck, float* __restrict__ dwr,
                                                               float* __restrict__ dwi, int B, int ci, int co, int MK, long MS,
                                                               int accumulate) {
  const int og = blockIdx.y, ig = blockIdx.z;
  const int j = blockIdx.x * kFmThreads + threadIdx.x;
  if (j >= MK) return;
  const int ik = __ldg(lin_k + j), in = __ldg(lin_n + j);
  const float c = 0.5f * __ldg(ck + j);
  float gr[kFmOG][kFmOG], gi[kFmOG][kFmOG];
#pragma unroll
  for (int o = 0; o < kFmOG; ++o)
#pragma unroll
    for (int i = 0; i < kFmOG; ++i) gr[o][i] = gi[o][i] = 0.f;
  for (int b = 0; b < B; ++b) {
    const float* dhb = dhp + (long)b * co * MS;
    const float* zb = z + ((long)b * ci + ig * kFmOG) * MS;
    float da[kFmOG], db[kFmOG], re[kFmOG], im[kFmOG];
#pragma unroll
    for (int o = 0; o < kFmOG; ++o) fmix_dab(dhb, MS, og * kFmOG + o, ik, in, c, da[o], db[o]);
#pragma unroll
    for (int i = 0; i < kFmOG; ++i) {
      const float hk = __ldg(zb + (long)i * MS + ik), hn = __ldg(zb + (long)i * MS + in);
      re[i] = 0.5f * (hk + hn);
      im[i] = 0.5f * (hn - hk);
    }
#pragma unroll
    for (int o = 0; o < kFmOG; ++o)
#pragma unroll
      for (int i = 0; i < kFmOG; ++i) {
        gr[o][i] = fmaf(da[o], re[i], fmaf(db[o], im[i], gr[o][i]));
        gi[o][i] = fmaf(db[o], re[i], fmaf(-da[o], im[i], gi[o][i]));
      }
  }
#pragma unroll
  for (int o = 0; o < kFmOG; ++o)
#pragma unroll
    for (int i = 0; i < kFmOG; ++i) {
      const long off = ((long)(og * kFmOG + o) * ci + ig * kFmOG + i) * MK + j;
      dwr[off] = accumulate ? dwr[off] + gr[o][i] : gr[o][i];
      dwi[off] = accumulate ? dwi[off] + gi[o][i] : gi[o][i];
    }
}

// ------------------------------------------------------------------------------------------------ host side
static int fmix_check(int B, int ci, int co, long MK, long MS) {
  HNO_CHECK(B >= 1 && B <= 65535 && ci >= kFmOG && co >= kFmOG && ci % kFmOG == 0 && co % kFmOG == 0,
            "fourier_mix: channels must be multiples of %d (got %d -> %d)", kFmOG, ci, co);
  HNO_CHECK(MK >= 1 && MS >= MK && MK < (1L << 30) && MS < (1L << 31), "fourier_mix: bad mode counts");
  HNO_CHECK((size_t)2 * (ci + co) * kFmPitch * sizeof(float) <= 200 * 1024 && (size_t)2 * kFmOG * (ci > co ? ci : co) * 4 <= 48 * 1024,
            "fourier_mix: too many channels");
  return 0;
}

size_t fourier_mix_workspace_bytes(int ci, int co, long MK, int B) {
  return (size_t)ceil_div(MK, kFmTile) * B * 2 * ci * co * sizeof(float) + 256;
}

int fourier_mix_forward(const float* z, const float* wr, const float* wi, const int* lin_k, const int* lin_n, const float* ck,
                        float* hp, int B, int ci, int co, long MK, long MS, int individual, cudaStream_t st) {
  HNO_CHECK(z && wr && wi && lin_k && lin_n && ck && hp, "fourier_mix_forward: null pointer");
  if (fmix_check(B, ci, co, MK, MS)) return -1;
  HNO_CUDA(cudaMemsetAsync(hp, 0, (size_t)B * co * MS * sizeof(float), st));
  dim3 grid(ceil_div(MK, kFmThreads), co / kFmOG, B);
  if (individual)
    k_fmix_fwd_ind<<<grid, kFmThreads, 0, st>>>(z, wr, wi, lin_k, lin_n, ck, hp, ci, co, (int)MK, MS);
  else
    k_fmix_fwd<<<grid, kFmThreads, (size_t)2 * kFmOG * ci * sizeof(float), st>>>(z, wr, wi, lin_k, lin_n, ck, hp, ci, co, (int)MK, MS);
  HNO_LAUNCH_CHECK();
  return 0;
}

int fourier_mix_backward(const float* dhp, const float* z, const float* wr, const float* wi, const int* lin_k, const int* lin_n,
                         const float* ck, float* dz, float* dwr, float* dwi, void* ws, int B, int ci, int co, long MK, long MS,
                         int individual, int accumulate_dw, cudaStream_t st) {
  HNO_CHECK(dhp && z && wr && wi && lin_k && lin_n && ck, "fourier_mix_backward: null pointer");
  HNO_CHECK((dwr == nullptr) == (dwi == nullptr), "fourier_mix_backward: dw_real and dw_imag come as a pair");
  if (fmix_check(B, ci, co, MK, MS)) return -1;
  if (dz) {
    HNO_CUDA(cudaMemsetAsync(dz, 0, (size_t)B * ci * MS * sizeof(float), st));
    dim3 grid(ceil_div(MK, kFmThreads), ci / kFmOG, B);
    if (individual)
      k_fmix_bwd_x_ind<<<grid, kFmThreads, 0, st>>>(dhp, wr, wi, lin_k, lin_n, ck, dz, ci, co, (int)MK, MS);
    else
      k_fmix_bwd_x<<<grid, kFmThreads, (size_t)2 * kFmOG * co * sizeof(float), st>>>(dhp, wr, wi, lin_k, lin_n, ck, dz, ci, co, (int)MK, MS);
    HNO_LAUNCH_CHECK();
  }
  if (dwr && individual) {
    HNO_CHECK(ci / kFmOG <= 65535 && co / kFmOG <= 65535, "fourier_mix_backward: too many channels");
    dim3 grid(ceil_div(MK, kFmThreads), co / kFmOG, ci / kFmOG);
    k_fmix_bwd_w_ind<<<grid, kFmThreads, 0, st>>>(dhp, z, lin_k, lin_n, ck, dwr, dwi, B, ci, co, (int)MK, MS, accumulate_dw);
    HNO_LAUNCH_CHECK();
    return 0;
  }
  if (dwr) {
    HNO_CHECK(ws, "fourier_mix_backward: the weight gradient needs the workspace");
    float* partials = reinterpret_cast<float*>(ws);
    dim3 grid(ceil_div(MK, kFmTile), B);
    const size_t smem = (size_t)2 * (ci + co) * kFmPitch * sizeof(float);
    HNO_CUDA(cudaFuncSetAttribute(k_fmix_bwd_w, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k_fmix_bwd_w<<<grid, 256, smem, st>>>(dhp, z, lin_k, lin_n, ck, partials, ci, co, (int)MK, MS);
    HNO_LAUNCH_CHECK();
    // rows of [co ci | co ci]: the first half sums into dw_real, the second into dw_imag
    return reduce_partials(partials, (int)(grid.x * grid.y), co * ci, co * ci, dwr, dwi, accumulate_dw, st);
  }
  return 0;
}

}  // namespace hno
