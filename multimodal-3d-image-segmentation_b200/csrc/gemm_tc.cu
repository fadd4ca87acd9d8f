// Batched "TN" GEMM for sm_100a (tcgen05 + TMEM + TMA, 3xTF32) with fused attention epilogues, and its CUDA-core twin.
//
//   C[b][m][n] = EPI( alpha * sum_k A[b][m][k] * B[b][n][k] )          (both operands K-major: k contiguous in memory)
//
// This is the contraction engine of the Hartley multi-head attention (reference nets/hartley_mha.py:198-204):
//   att = einsum('bzcq,bzck->bzqk', query, key) / sqrt(c);  att = selu(att);  out = einsum('bzqk,bzck->bzcq', att, value)
// and of its backward.  Unlike the streamed contractions of the transform (tc_stream.cu: one huge operand read once, a
// tiny resident one), BOTH operands are streamed here and the work is FLOP-bound (1,960 tokens x 96 features x 4 heads:
// 5.9 GFLOP per sample and block forward), so it is tiled like a classical tensor-core GEMM:
//   * CTA tile 128 (m) x BN (n), K walked in chunks of 32 floats = one 128-byte swizzle row per operand row;
//   * warp 8 (one thread) fills an NST-deep ring with two TMA tensor copies per chunk (SWIZZLE_128B boxes of
//     [rows][32 floats]; the landed bytes ARE the K-major "hi" operands: the tensor core ignores the 13 low mantissa bits);
//   * warps 4-7 write the "lo" images (x - trunc_tf32(x), same swizzled offsets) next to them;
//   * warp 9 (one thread) issues per 8-wide k-step  lo*hi + hi*lo + hi*hi  as three tcgen05.mma.kind::tf32 into one of two
//     TMEM accumulators (fp32) and commits to mbarriers that free the ring stage / publish the accumulator;
//   * warps 0-3 run the epilogue of tile t while the MMAs of tile t+1 execute: tcgen05.ld (one TMEM lane = one row m per
//     thread), scale / SELU / multiplication with selu'(E), then the row-major tile C (128-byte runs per thread) and / or
//     its transpose C^T (coalesced 128-byte rows per column) -- the transposed copy is what lets every later GEMM of the
//     attention backward read K-major operands too.
// Precision (stated choice, as everywhere in this library): 3xTF32, relative error per product ~2^-22.
#include "common.cuh"
#include "tc_common.cuh"
#include "tc_stream.h"
#include "gemm_tc.h"

#include <stdlib.h>

namespace hno {

using namespace tc;

constexpr int kGmBM = 128;
constexpr int kGmKC = 32;
constexpr int kGmThreads = 320;  // warps 0-3 epilogue, 4-7 operand split, 8 TMA producer, 9 MMA issuer

struct GmDev {
  float* c;
  long ldc, sc;
  float* ct;
  long ldct, sct;
  const float* e;
  long lde, se;
  float alpha;
  int epi;
  int nk;           // K / 32
  int kseg;         // chunks per accumulation segment (== nk: one segment)
  int mt, nt;       // tiles along m and n
  int total_tiles;  // batch * mt * nt
};

__device__ __forceinline__ void gm_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

template <int BN>
struct GmCfg {
  static constexpr int kStageRaw = kGmBM * 128 + BN * 128;  // bytes: A rows then B rows, 128 bytes each
  static constexpr int kStageBytes = 2 * kStageRaw;          // + the lo images
  static constexpr int NST = (3 * kStageBytes + 2048 <= 227 * 1024) ? 3 : 2;
  static constexpr uint32_t kTmemCols = 2 * BN <= 32 ? 32 : (2 * BN <= 64 ? 64 : (2 * BN <= 128 ? 128 : (2 * BN <= 256 ? 256 : 512)));
  static size_t smem_bytes() { return 1024 + (size_t)NST * kStageBytes; }
};

template <int BN>
__global__ void __launch_bounds__(kGmThreads, 1) k_gemm_tn_tc(const __grid_constant__ CUtensorMap tma,
                                                             const __grid_constant__ CUtensorMap tmb, const GmDev p) {
  constexpr int NST = GmCfg<BN>::NST;
  constexpr int kStageRaw = GmCfg<BN>::kStageRaw;
  constexpr int kStageBytes = GmCfg<BN>::kStageBytes;
  constexpr uint32_t kTmemCols = GmCfg<BN>::kTmemCols;
  constexpr uint32_t kIdesc = make_idesc_tf32(128, BN, 0, 0);
  static_assert(BN % 32 == 0 && BN >= 32 && BN <= 256, "BN must be a multiple of 32 in [32, 256]");
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  __shared__ __align__(8) uint64_t bar_full[NST];    // TMA bytes landed                       (1 arrival + tx)
  __shared__ __align__(8) uint64_t bar_split[NST];   // lo images written                      (128 arrivals)
  __shared__ __align__(8) uint64_t bar_done[NST];    // MMAs of the chunk retired              (tcgen05.commit)
  __shared__ __align__(8) uint64_t bar_accfull[2];   // all MMAs of a tile retired             (tcgen05.commit)
  __shared__ __align__(8) uint64_t bar_accfree[2];   // accumulator drained by the epilogue    (128 arrivals)
  __shared__ uint32_t tmem_slot;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  if (tid == 0) {
#pragma unroll
    for (int s = 0; s < NST; ++s) {
      mbar_init(&bar_full[s], 1);
      mbar_init(&bar_split[s], 128);
      mbar_init(&bar_done[s], 1);
    }
    mbar_init(&bar_accfull[0], 1);
    mbar_init(&bar_accfull[1], 1);
    mbar_init(&bar_accfree[0], 128);
    mbar_init(&bar_accfree[1], 128);
    mbar_fence_init();
  }
  if (warp == 8) tmem_alloc(&tmem_slot, kTmemCols);
  fence_proxy_async_smem();
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem = tmem_slot;
  const int my_tiles = p.total_tiles > (int)blockIdx.x ? (p.total_tiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
  const int nk = p.nk;
  const int per_b = p.mt * p.nt;

  if (warp == 8) {
    // =============================================================== TMA producer (one thread)
    if (lane == 0) {
      tma_prefetch_desc(&tma);
      tma_prefetch_desc(&tmb);
      int it = 0, s = 0;
      uint32_t ph = 0;
      for (int ti = 0; ti < my_tiles; ++ti) {
        const int tile = blockIdx.x + ti * gridDim.x;
        const int b = tile / per_b;
        const int r = tile - b * per_b;
        const int m0 = (r / p.nt) * kGmBM, n0 = (r % p.nt) * BN;
        for (int c = 0; c < nk; ++c) {
          if (it >= NST) mbar_wait(&bar_done[s], ph ^ 1);  // previous use of the stage (raw + lo) fully consumed
          uint8_t* dst = smem + s * kStageBytes;
          mbar_expect_tx(&bar_full[s], kStageRaw);
          tma_load_3d(dst, &tma, c * kGmKC, m0, b, &bar_full[s]);
          tma_load_3d(dst + kGmBM * 128, &tmb, c * kGmKC, n0, b, &bar_full[s]);
          ++it;
          if (++s == NST) {
            s = 0;
            ph ^= 1;
          }
        }
      }
    }
    __syncwarp();
  } else if (warp == 9) {
    // =============================================================== MMA issuer (one thread)
    if (lane == 0) {
      int s = 0;
      uint32_t ph = 0;
      const int nseg = (nk + p.kseg - 1) / p.kseg;
      int segc = 0;  // running segment index of this CTA: accumulator buffer segc & 1
      for (int ti = 0; ti < my_tiles; ++ti) {
        for (int sg = 0; sg < nseg; ++sg, ++segc) {
          const int buf = segc & 1;
          if (segc >= 2) mbar_wait(&bar_accfree[buf], (uint32_t)(((segc >> 1) - 1) & 1));
          tc_fence_after_sync();
          const uint32_t acc = tmem + buf * BN;
          const int c0 = sg * p.kseg, c1 = min(nk, c0 + p.kseg);
          for (int c = c0; c < c1; ++c) {
            mbar_wait(&bar_split[s], ph);
            tc_fence_after_sync();
            const uint32_t a_hi = smem_u32(smem + s * kStageBytes), b_hi = a_hi + kGmBM * 128;
            const uint32_t a_lo = a_hi + kStageRaw, b_lo = b_hi + kStageRaw;
#pragma unroll
            for (int j = 0; j < kGmKC / 8; ++j) {  // one MMA = 8 k = 32 bytes along the swizzled 128-byte row
              const uint64_t dah = make_smem_desc(a_hi + j * 32, 16, 1024, kLayoutSw128);
              const uint64_t dal = make_smem_desc(a_lo + j * 32, 16, 1024, kLayoutSw128);
              const uint64_t dbh = make_smem_desc(b_hi + j * 32, 16, 1024, kLayoutSw128);
              const uint64_t dbl = make_smem_desc(b_lo + j * 32, 16, 1024, kLayoutSw128);
              mma_tf32(acc, dal, dbh, kIdesc, !(c == c0 && j == 0));
              mma_tf32(acc, dah, dbl, kIdesc, true);
              mma_tf32(acc, dah, dbh, kIdesc, true);
            }
            mma_commit(&bar_done[s]);
            if (c == c1 - 1) mma_commit(&bar_accfull[buf]);
            if (++s == NST) {
              s = 0;
              ph ^= 1;
            }
          }
        }
      }
    }
    __syncwarp();
  } else if (warp >= 4) {
    // =============================================================== operand split: lo = x - trunc_tf32(x)
    const int t = tid - 128;
    int s = 0;
    uint32_t ph = 0;
    for (int ti = 0; ti < my_tiles; ++ti) {
      for (int c = 0; c < nk; ++c) {
        mbar_wait(&bar_full[s], ph);
        const float4* r4 = reinterpret_cast<const float4*>(smem + s * kStageBytes);
        float4* l4 = reinterpret_cast<float4*>(smem + s * kStageBytes + kStageRaw);
#pragma unroll
        for (int i = 0; i < kStageRaw / 16 / 128; ++i) {
          const float4 x = r4[t + i * 128];
          l4[t + i * 128] = make_float4(tf32_lo(x.x), tf32_lo(x.y), tf32_lo(x.z), tf32_lo(x.w));
        }
        fence_proxy_async_smem();
        gm_arrive(&bar_split[s]);
        if (++s == NST) {
          s = 0;
          ph ^= 1;
        }
      }
    }
  } else {
    // =============================================================== epilogue (warp w owns TMEM lanes 32 w .. 32 w + 31)
    // Long contractions (K = 2,048 tokens in att V and the three token-contracting backward GEMMs) are accumulated in
    // SEGMENTS of kseg chunks: the tensor core adds into its fp32 accumulator with truncation, one rounding per MMA, so 768
    // sequential MMAs leave a bias of ~1.6e-5 relative (measured against the fp64 oracle; the CUDA-core route: 8e-7).  Each
    // segment lands in one of the two TMEM buffers and is folded into registers here with round-to-nearest adds.
    const int nseg = (nk + p.kseg - 1) / p.kseg;
    int segc = 0;
    for (int ti = 0; ti < my_tiles; ++ti) {
      const int tile = blockIdx.x + ti * gridDim.x;
      const int b = tile / per_b;
      const int r = tile - b * per_b;
      const int m = (r / p.nt) * kGmBM + warp * 32 + lane;
      const int nbase = (r % p.nt) * BN;
      float* crow = p.c ? p.c + (long)b * p.sc + (long)m * p.ldc + nbase : nullptr;
      float* ctcol = p.ct ? p.ct + (long)b * p.sct + (long)nbase * p.ldct + m : nullptr;
      const float* erow = p.e ? p.e + (long)b * p.se + (long)m * p.lde + nbase : nullptr;
      auto finish = [&](float (&v)[32], int n0) {
        if (p.epi == 1) {
#pragma unroll
          for (int j = 0; j < 32; j += 2) {
            const float2 q = selu2(make_float2(p.alpha * v[j], p.alpha * v[j + 1]));
            v[j] = q.x;
            v[j + 1] = q.y;
          }
        } else if (p.epi == 2) {
#pragma unroll
          for (int j = 0; j < 32; j += 4) {
            const float4 y = __ldg(reinterpret_cast<const float4*>(erow + n0 + j));
            v[j] *= p.alpha * selu_grad_from_out(y.x);
            v[j + 1] *= p.alpha * selu_grad_from_out(y.y);
            v[j + 2] *= p.alpha * selu_grad_from_out(y.z);
            v[j + 3] *= p.alpha * selu_grad_from_out(y.w);
          }
        } else {
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] *= p.alpha;
        }
        if (crow) {
#pragma unroll
          for (int j = 0; j < 32; j += 4)
            *reinterpret_cast<float4*>(crow + n0 + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
        }
        if (ctcol) {
#pragma unroll
          for (int j = 0; j < 32; ++j) ctcol[(long)(n0 + j) * p.ldct] = v[j];
        }
      };
      if (nseg == 1) {
        const int buf = segc & 1;
        mbar_wait(&bar_accfull[buf], (uint32_t)((segc >> 1) & 1));
        tc_fence_after_sync();
        const uint32_t acc = tmem + buf * BN + ((uint32_t)(warp * 32) << 16);
#pragma unroll 1
        for (int n0 = 0; n0 < BN; n0 += 32) {
          float v[32];
          tmem_ld32(acc + n0, v);
          if (n0 + 32 >= BN) {  // last read of this accumulator buffer
            tc_fence_before_sync();
            gm_arrive(&bar_accfree[buf]);
          }
          finish(v, n0);
        }
        ++segc;
      } else if constexpr (BN <= 128) {
        float sum[BN / 32][32];
#pragma unroll
        for (int q = 0; q < BN / 32; ++q)
#pragma unroll
          for (int j = 0; j < 32; ++j) sum[q][j] = 0.f;
        for (int sg = 0; sg < nseg; ++sg, ++segc) {
          const int buf = segc & 1;
          mbar_wait(&bar_accfull[buf], (uint32_t)((segc >> 1) & 1));
          tc_fence_after_sync();
          const uint32_t acc = tmem + buf * BN + ((uint32_t)(warp * 32) << 16);
#pragma unroll
          for (int q = 0; q < BN / 32; ++q) {
            float v[32];
            tmem_ld32(acc + 32 * q, v);
#pragma unroll
            for (int j = 0; j < 32; ++j) sum[q][j] += v[j];
          }
          tc_fence_before_sync();
          gm_arrive(&bar_accfree[buf]);
        }
#pragma unroll
        for (int q = 0; q < BN / 32; ++q) finish(sum[q], 32 * q);
      }
    }
  }
  tc_fence_before_sync();
  __syncthreads();
  if (warp == 8) tmem_dealloc(tmem, kTmemCols);
}

// ------------------------------------------------------------------------------------------------ CUDA-core twin
// Plain shared-memory tiled fp32 GEMM (64 x 64 x 16 tiles, 4 x 4 outputs per thread) with the same epilogues: the route
// of hno_set_tensor_cores(0) and the cross-check of the tensor-core kernel in the GPU tests.  Any M, N, K.
__global__ void __launch_bounds__(256) k_gemm_tn_ffma(const float* __restrict__ A, long lda, long sa,
                                                      const float* __restrict__ B, long ldb, long sb, GmDev p, int M, int N,
                                                      int K) {
  __shared__ float sA[16][64 + 4], sB[16][64 + 4];
  const int b = blockIdx.z;
  const int m0 = blockIdx.y * 64, n0 = blockIdx.x * 64;
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  const float* a = A + (long)b * sa;
  const float* bb = B + (long)b * sb;
  float acc[4][4] = {};
  for (int k0 = 0; k0 < K; k0 += 16) {
    for (int i = threadIdx.x; i < 64 * 16; i += 256) {
      const int r = i >> 4, k = i & 15;
      sA[k][r] = (m0 + r < M && k0 + k < K) ? __ldg(a + (long)(m0 + r) * lda + k0 + k) : 0.f;
      sB[k][r] = (n0 + r < N && k0 + k < K) ? __ldg(bb + (long)(n0 + r) * ldb + k0 + k) : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < 16; ++k) {
      float av[4], bv[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        av[i] = sA[k][ty * 4 + i];
        bv[i] = sB[k][tx * 4 + i];
      }
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int m = m0 + ty * 4 + i;
    if (m >= M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int n = n0 + tx * 4 + j;
      if (n >= N) continue;
      float v = p.alpha * acc[i][j];
      if (p.epi == 1) v = selu_f(v);
      else if (p.epi == 2) v *= selu_grad_from_out(__ldg(p.e + (long)b * p.se + (long)m * p.lde + n));
      if (p.c) p.c[(long)b * p.sc + (long)m * p.ldc + n] = v;
      if (p.ct) p.ct[(long)b * p.sct + (long)n * p.ldct + m] = v;
    }
  }
}

// ------------------------------------------------------------------------------------------------ host side
static bool gemm_tc_eligible(const GemmArgs& g) {
  if (!tc_enabled()) return false;
  if (g.M % kGmBM || g.K % kGmKC || g.N % 32) return false;
  if (g.N > 256 && g.N % 128) return false;
  if (g.lda % 4 || g.ldb % 4 || g.sa % 4 || g.sb % 4) return false;
  if ((reinterpret_cast<uintptr_t>(g.a) | reinterpret_cast<uintptr_t>(g.b)) & 15) return false;
  if (g.c && ((reinterpret_cast<uintptr_t>(g.c) & 15) || g.ldc % 4 || g.sc % 4)) return false;
  if (g.e && ((reinterpret_cast<uintptr_t>(g.e) & 15) || g.lde % 4 || g.se % 4)) return false;
  return true;
}

template <int BN>
static int gemm_tc_launch(const GemmArgs& g, cudaStream_t st) {
  CUtensorMap tma, tmb;
  {
    const uint64_t dims[3] = {(uint64_t)g.K, (uint64_t)g.M, (uint64_t)g.batch};
    const uint64_t strides[2] = {(uint64_t)g.lda * 4, (uint64_t)(g.batch > 1 ? g.sa : (long)g.M * g.lda) * 4};
    const uint32_t box[3] = {kGmKC, kGmBM, 1};
    if (int rc = encode_tensor_map(&tma, g.a, 3, dims, strides, box, 1)) return rc;
  }
  {
    const uint64_t dims[3] = {(uint64_t)g.K, (uint64_t)g.N, (uint64_t)g.batch};
    const uint64_t strides[2] = {(uint64_t)g.ldb * 4, (uint64_t)(g.batch > 1 ? g.sb : (long)g.N * g.ldb) * 4};
    const uint32_t box[3] = {kGmKC, (uint32_t)BN, 1};
    if (int rc = encode_tensor_map(&tmb, g.b, 3, dims, strides, box, 1)) return rc;
  }
  GmDev p;
  p.c = g.c, p.ldc = g.ldc, p.sc = g.sc;
  p.ct = g.ct, p.ldct = g.ldct, p.sct = g.sct;
  p.e = g.e, p.lde = g.lde, p.se = g.se;
  p.alpha = g.alpha;
  p.epi = g.epi;
  p.nk = g.K / kGmKC;
  // segments of 4 chunks = 128 k = 48 MMAs (see the epilogue); tiles wider than 128 columns keep one segment (their
  // contractions are short: K = features per head)
  p.kseg = (BN <= 128 && p.nk > 8) ? 4 : p.nk;
  p.mt = g.M / kGmBM;
  p.nt = g.N / BN;
  p.total_tiles = g.batch * p.mt * p.nt;
  const size_t smem = GmCfg<BN>::smem_bytes();
  auto kern = k_gemm_tn_tc<BN>;
  HNO_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int grid = sm_count();
  if (grid > p.total_tiles) grid = p.total_tiles;
  kern<<<grid, kGmThreads, smem, st>>>(tma, tmb, p);
  HNO_LAUNCH_CHECK();
  return 0;
}

int gemm_tn(const GemmArgs& g, cudaStream_t st) {
  HNO_CHECK(g.a && g.b && (g.c || g.ct), "gemm_tn: null pointer");
  HNO_CHECK(g.batch >= 1 && g.M >= 1 && g.N >= 1 && g.K >= 1, "gemm_tn: bad sizes");
  HNO_CHECK(g.epi >= 0 && g.epi <= 2 && (g.epi != 2 || g.e), "gemm_tn: bad epilogue");
  if (gemm_tc_eligible(g)) {
    const int bn = g.N <= 256 ? g.N : 128;
    switch (bn) {
      case 32: return gemm_tc_launch<32>(g, st);
      case 64: return gemm_tc_launch<64>(g, st);
      case 96: return gemm_tc_launch<96>(g, st);
      case 128: return gemm_tc_launch<128>(g, st);
      case 160: return gemm_tc_launch<160>(g, st);
      case 192: return gemm_tc_launch<192>(g, st);
      case 224: return gemm_tc_launch<224>(g, st);
      case 256: return gemm_tc_launch<256>(g, st);
      default: break;
    }
  }
  GmDev p;
  p.c = g.c, p.ldc = g.ldc, p.sc = g.sc;
  p.ct = g.ct, p.ldct = g.ldct, p.sct = g.sct;
  p.e = g.e, p.lde = g.lde, p.se = g.se;
  p.alpha = g.alpha;
  p.epi = g.epi;
  p.nk = p.kseg = p.mt = p.nt = p.total_tiles = 0;
  HNO_CHECK(g.batch <= 65535, "gemm_tn: batch too large for the CUDA-core route");
  dim3 grid(ceil_div(g.N, 64), ceil_div(g.M, 64), g.batch);
  k_gemm_tn_ffma<<<grid, 256, 0, st>>>(g.a, g.lda, g.sa, g.b, g.ldb, g.sb, p, g.M, g.N, g.K);
  HNO_LAUNCH_CHECK();
  return 0;
}

}  // namespace hno
