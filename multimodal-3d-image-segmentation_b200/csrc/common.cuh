// Shared helpers for the hno_b200 CUDA library (sm_100a only).
//
// Everything in csrc/ is compiled into ONE shared object, libhno_b200.so, whose
// only public surface is the extern "C" API declared in include/hno_b200.h.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stddef.h>

namespace hno {

// Thread-local error string returned by hno_last_error().
void set_error(const char* fmt, ...);

#define HNO_CHECK(cond, ...)                 \
  do {                                       \
    if (!(cond)) {                           \
      ::hno::set_error(__VA_ARGS__);         \
      return -1;                             \
    }                                        \
  } while (0)

#define HNO_CUDA(expr)                                                            \
  do {                                                                            \
    cudaError_t e_ = (expr);                                                      \
    if (e_ != cudaSuccess) {                                                      \
      ::hno::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e_),    \
                       __FILE__, __LINE__);                                       \
      return -2;                                                                  \
    }                                                                             \
  } while (0)

// Every kernel launch of the library goes through this macro; the counter backs hno_launch_count().
void count_launch();
#define HNO_LAUNCH_CHECK()        \
  do {                            \
    ::hno::count_launch();        \
    HNO_CUDA(cudaGetLastError()); \
  } while (0)

// SELU constants exactly as PyTorch defines them (aten/src/ATen/native/Activation.cpp);
// the reference applies F.selu everywhere (nets/nets_utils.py:127-133, nets/hnosegxs.py:267-268,325-327).
constexpr float kSeluAlpha = 1.6732632423543772848170429916717f;
constexpr float kSeluScale = 1.0507009873554804934193349852946f;
constexpr float kSeluNeg = kSeluAlpha * kSeluScale;

// Packed fp32 math (Blackwell FFMA2 / FMUL2, PTX fma.rn.f32x2): two independent operations per issue slot.  The
// 3-operand scalar FFMA issues at half rate on sm_100 (register-file read ports), so every FMA-heavy loop uses it.
__device__ __forceinline__ float2 ffma2(float2 a, float2 b, float2 c) { return __ffma2_rn(a, b, c); }
__device__ __forceinline__ float2 fmul2(float2 a, float2 b) { return __fmul2_rn(a, b); }
__device__ __forceinline__ float2 dup2(float x) { return make_float2(x, x); }
// The same on operands that LIVE as 64-bit registers: values built lane by lane (x in one statement, y in another) are
// kept in unrelated scalar registers by the compiler, which then re-packs them with two MOVs in front of EVERY packed
// FMA (measured: 2 of 3 instructions of the input-gradient loops).  A b64 register is an aligned pair by construction.
typedef unsigned long long pk2;
__device__ __forceinline__ pk2 pack2(float lo, float hi) {
  pk2 r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ float2 unpack2(pk2 v) {
  float2 r;
  asm("mov.b64 {%0, %1}, %2;" : "=f"(r.x), "=f"(r.y) : "l"(v));
  return r;
}
__device__ __forceinline__ pk2 ffma2p(pk2 a, pk2 b, pk2 c) {
  pk2 d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
  return d;
}

// SELU with a purpose-built expm1 for the negative branch (one SFU op and ~11 ALU ops instead of ~29 with several
// SFU / conversion ops for libm's expm1f -- the activation epilogues were 44 % of the dynamic instructions of the
// 48->24 channel mix):
//   x in (-0.5, 0]: x * q(x), q = degree-5 minimax fit of (e^x - 1) / x on [-0.5, 0] (fit error 1e-9; evaluated in
//                   fp32 Horner form the relative error is <= 1.04e-7, i.e. rounding limited, no cancellation)
//   x <= -0.5     : ex2.approx(x * log2 e) - 1   (relative error of the result <= 1.6 * 2^-22)
// A plain exp(x) - 1 is NOT acceptable: its cancellation near 0 (relative error ~6e-8/|x|) triples the gradient
// error of the whole network.  Measured on the oracle (fp32, flat gradient rel-L2 vs the fp64 oracle, 120x112x77):
// libm expm1 3.8e-4, this scheme 3.7e-4, exp(x) - 1 1.05e-3.
constexpr float kSeluQ0 = kSeluNeg * 0.9999999987618426f;
constexpr float kSeluQ1 = kSeluNeg * 0.49999982068866156f;
constexpr float kSeluQ2 = kSeluNeg * 0.1666624492152664f;
constexpr float kSeluQ3 = kSeluNeg * 0.041630243386182875f;
constexpr float kSeluQ4 = kSeluNeg * 0.008190002775750417f;
constexpr float kSeluQ5 = kSeluNeg * 0.001123715104247875f;
constexpr float kLog2e = 1.4426950408889634f;

__device__ __forceinline__ float ex2_approx(float x) {
  float e;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(x));
  return e;
}
__device__ __forceinline__ float selu_f(float x) {
  float p = kSeluQ5;
  p = fmaf(p, x, kSeluQ4);
  p = fmaf(p, x, kSeluQ3);
  p = fmaf(p, x, kSeluQ2);
  p = fmaf(p, x, kSeluQ1);
  p = fmaf(p, x, kSeluQ0);
  p *= x;
  const float e = fmaf(kSeluNeg, ex2_approx(x * kLog2e), -kSeluNeg);
  const float neg = x > -0.5f ? p : e;
  return x > 0.f ? kSeluScale * x : neg;
}
// two values at once on the packed pipe (same arithmetic, bit-identical to selu_f per element)
__device__ __forceinline__ float2 selu2(float2 x) {
  float2 p = dup2(kSeluQ5);
  p = ffma2(p, x, dup2(kSeluQ4));
  p = ffma2(p, x, dup2(kSeluQ3));
  p = ffma2(p, x, dup2(kSeluQ2));
  p = ffma2(p, x, dup2(kSeluQ1));
  p = ffma2(p, x, dup2(kSeluQ0));
  p = fmul2(p, x);
  const float2 t = fmul2(x, dup2(kLog2e));
  const float2 e = ffma2(dup2(kSeluNeg), make_float2(ex2_approx(t.x), ex2_approx(t.y)), dup2(-kSeluNeg));
  const float2 pos = fmul2(x, dup2(kSeluScale));
  float2 r;
  r.x = x.x > 0.f ? pos.x : (x.x > -0.5f ? p.x : e.x);
  r.y = x.y > 0.f ? pos.y : (x.y > -0.5f ? p.y : e.y);
  return r;
}
// d selu / d x expressed from the OUTPUT y = selu(x): scale for y>0, y + scale*alpha otherwise.
__device__ __forceinline__ float selu_grad_from_out(float y) {
  return y > 0.f ? kSeluScale : y + kSeluNeg;
}

template <int V>
struct Vec;
template <>
struct Vec<1> {
  float v[1];
  __device__ __forceinline__ static Vec ld(const float* p) {
    Vec r;
    r.v[0] = __ldg(p);
    return r;
  }
  __device__ __forceinline__ void st(float* p) const { *p = v[0]; }
};
template <>
struct Vec<2> {
  float v[2];
  __device__ __forceinline__ static Vec ld(const float* p) {
    float2 t = __ldg(reinterpret_cast<const float2*>(p));
    Vec r;
    r.v[0] = t.x;
    r.v[1] = t.y;
    return r;
  }
  __device__ __forceinline__ void st(float* p) const {
    *reinterpret_cast<float2*>(p) = make_float2(v[0], v[1]);
  }
};
template <>
struct Vec<4> {
  float v[4];
  __device__ __forceinline__ static Vec ld(const float* p) {
    float4 t = __ldg(reinterpret_cast<const float4*>(p));
    Vec r;
    r.v[0] = t.x;
    r.v[1] = t.y;
    r.v[2] = t.z;
    r.v[3] = t.w;
    return r;
  }
  __device__ __forceinline__ void st(float* p) const {
    *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
  }
};

// ---- Ampere-style asynchronous global -> shared copies (LDGSTS).  Used as PER-THREAD rings: a thread only reads
// back bytes it copied itself, so cp.async.wait_group is the only synchronisation the pipelines need.
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait_group() {
  asm volatile("cp.async.wait_group %0;\n" ::"n"(N) : "memory");
}
template <int V>
__device__ __forceinline__ void cp_async_vec(float* smem_dst, const float* gsrc) {
  const unsigned dst = (unsigned)__cvta_generic_to_shared(smem_dst);
  if (V == 4)
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(dst), "l"(gsrc) : "memory");
  else if (V == 2)
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;\n" ::"r"(dst), "l"(gsrc) : "memory");
  else
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;\n" ::"r"(dst), "l"(gsrc) : "memory");
}

__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;\n" ::: "memory"); }

// Largest vector width in {4,2,1} that divides every given stride / count and the pointer alignment.
inline int pick_vec(const void* const* ptrs, int nptr, const long* counts, int ncount) {
  int v = 4;
  for (; v > 1; v >>= 1) {
    bool ok = true;
    for (int i = 0; i < nptr && ok; ++i)
      if (ptrs[i] && (reinterpret_cast<uintptr_t>(ptrs[i]) % (sizeof(float) * v))) ok = false;
    for (int i = 0; i < ncount && ok; ++i)
      if (counts[i] % v) ok = false;
    if (ok) break;
  }
  return v;
}

inline int ceil_div(long a, long b) { return (int)((a + b - 1) / b); }

// Number of SMs of the current device (cached); B200 = 148.
int sm_count();

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double warp_sum_d(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

}  // namespace hno
