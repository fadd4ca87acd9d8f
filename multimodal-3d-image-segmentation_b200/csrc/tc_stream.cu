// Streamed tall-skinny tensor-core contraction for sm_100a (tcgen05 + TMEM + TMA), 3xTF32.
//
//   out[g][n][m] = EPI( sum_k  A[g][k][m] * Bm[n][k] )        m = contiguous axis (voxels / plane columns)
//
// A is the big streamed operand: channel-planar activations [g][k][m] read exactly once from HBM by TMA (m contiguous,
// i.e. "MN-major" for the tensor core).  Bm is tiny (weights or cas/cos/sin basis rows) and stays resident in shared
// memory.  The same kernel therefore serves
//   * the pointwise 1x1x1 convolutions      (k = input channel,  n = output channel)   nets/nets_utils.py:120-174
//   * the D-axis analysis of the truncated DHT (k = d, n = retained cos/sin row)         nets/hnosegxs.py:378-410
//   * the D-axis synthesis of the adjoint DHT  (k = retained row, n = d)                 nets/hnosegxs.py:454-494
// which together move > 80 % of the bytes of an HNOSeg-XS training step.
//
// Precision (stated choice): 3xTF32.  Every fp32 operand x is split into hi = rna_tf32(x) and lo = rna_tf32(x - hi)
// (both exactly representable in TF32, so the tensor core's input truncation is a no-op) and each product is
// evaluated as lo*hi + hi*lo + hi*hi with fp32 accumulation in TMEM: relative error per product ~2^-22, unbiased.
// Plain TF32 (one MMA) does NOT meet the parity bar of this model.
//
// Structure: persistent CTAs of 128 threads, 2-3 resident per SM.  Work items are (tile of 128 m, chunk of KC k-rows).
// Thread 0 keeps NST-1 TMA chunk loads in flight (mbarrier complete_tx), all threads split the landed chunk in place
// (hi) and into a second buffer (lo), thread 0 issues the tcgen05.mma instructions and commits them to an mbarrier
// that (a) frees the stage for the next TMA load and (b) releases the epilogue, which reads the accumulator with
// tcgen05.ld (one TMEM lane = one m per thread) and writes coalesced 128-byte rows per output channel.
#include "common.cuh"
#include "tc_common.cuh"
#include "tc_stream.h"

#include <atomic>
#include <mutex>

namespace hno {

using namespace tc;

constexpr int kTcThreads = 128;

struct TcDev {
  const float* b;
  long ldbn, ldbk;
  int nvalid, kvalid;
  float scale;
  const float* bias;
  float* out;
  long ldo, gso;
  int nout;
  long mext, valid_m;
  int nchunk, chunks_per_src;
  int tiles_per_slab;
  long total_tiles;
  int act, epi;
};

template <int KC, int NPAD, int NST>
struct TcSmem {
  static constexpr int kChunkBytes = KC * 512;  // 4 blocks of 32 m x KC rows x 4 B
  static size_t bytes(int nchunk) {
    return 1024 /* alignment slack */ + (size_t)(NST + 2) * kChunkBytes + (size_t)2 * NPAD * nchunk * KC * 4 + NPAD * 4;
  }
};

template <int KC, int NPAD, int NST>
__global__ void __launch_bounds__(kTcThreads) k_tc_stream(const __grid_constant__ CUtensorMap tm0,
                                                           const __grid_constant__ CUtensorMap tm1, const TcDev p) {
  constexpr int kChunkBytes = KC * 512;
  constexpr int NKG = KC / 8;
  constexpr uint32_t kIdesc = make_idesc_tf32(128, NPAD, 1, 0);
  constexpr uint32_t kTmemCols = NPAD < 32 ? 32 : NPAD;
  static_assert(KC % 8 == 0 && NPAD % 16 == 0 && NPAD <= 256, "bad tile configuration");
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* raw = smem;                                     // [NST][kChunkBytes]
  uint8_t* lob = raw + (size_t)NST * kChunkBytes;          // [2][kChunkBytes]
  float* bhi = reinterpret_cast<float*>(lob + 2 * kChunkBytes);
  const int ktot = p.nchunk * KC;
  float* blo = bhi + (size_t)NPAD * ktot;
  float* sbias = blo + (size_t)NPAD * ktot;
  __shared__ __align__(8) uint64_t bar_full[NST];
  __shared__ __align__(8) uint64_t bar_done[NST];
  __shared__ uint32_t tmem_slot;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  // ---- resident operand: B image (hi / lo), bias
  for (int idx = tid; idx < NPAD * ktot; idx += kTcThreads) {
    const int n = idx / ktot, k = idx - n * ktot;
    float v = 0.f;
    if (n < p.nvalid && k < p.kvalid) v = p.scale * __ldg(p.b + (long)n * p.ldbn + (long)k * p.ldbk);
    float hi, lo;
    hi = rna_tf32(v);
    lo = rna_tf32(v - hi);
    const int o = kmajor_plain_index<NPAD>(n, k);
    bhi[o] = hi;
    blo[o] = lo;
  }
  for (int n = tid; n < NPAD; n += kTcThreads) sbias[n] = (p.bias != nullptr && n < p.nout) ? __ldg(p.bias + n) : 0.f;
  if (tid == 0) {
#pragma unroll
    for (int s = 0; s < NST; ++s) {
      mbar_init(&bar_full[s], 1);
      mbar_init(&bar_done[s], 1);
    }
    mbar_fence_init();
    tma_prefetch_desc(&tm0);
    tma_prefetch_desc(&tm1);
  }
  if (warp == 0) tmem_alloc(&tmem_slot, kTmemCols);
  fence_proxy_async_smem();
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem = tmem_slot;

  const long my_tiles = p.total_tiles > blockIdx.x ? (p.total_tiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
  const long nitems = my_tiles * p.nchunk;

  auto issue_load = [&](long it) {  // thread 0 only
    const long ti = it / p.nchunk;
    const int c = (int)(it - ti * p.nchunk);
    const long tile = blockIdx.x + ti * gridDim.x;
    const int g = (int)(tile / p.tiles_per_slab);
    const int m0 = (int)(tile - (long)g * p.tiles_per_slab) * 128;
    const int src = c / p.chunks_per_src;
    const int row0 = (c - src * p.chunks_per_src) * KC;
    const int s = (int)(it % NST);
    uint8_t* dst = raw + (size_t)s * kChunkBytes;
    mbar_expect_tx(&bar_full[s], kChunkBytes);
    const CUtensorMap* tm = src == 0 ? &tm0 : &tm1;
#pragma unroll
    for (int j = 0; j < 4; ++j) tma_load_3d(dst + j * (KC * 128), tm, m0 + 32 * j, row0, g, &bar_full[s]);
  };

  if (tid == 0) {
    for (long it = 0; it < NST - 1 && it < nitems; ++it) issue_load(it);
  }

  for (long it = 0; it < nitems; ++it) {
    const int s = (int)(it % NST);
    const uint32_t ph = (uint32_t)((it / NST) & 1);
    const long ti = it / p.nchunk;
    const int c = (int)(it - ti * p.nchunk);
    mbar_wait(&bar_full[s], ph);
    if (it >= 2) {  // the lo buffer of item it-2 must have been consumed
      const long j = it - 2;
      mbar_wait(&bar_done[j % NST], (uint32_t)((j / NST) & 1));
    }
    // ---- operand split: hi in place, lo to the side buffer
    {
      float4* r4 = reinterpret_cast<float4*>(raw + (size_t)s * kChunkBytes);
      float4* l4 = reinterpret_cast<float4*>(lob + (size_t)(it & 1) * kChunkBytes);
#pragma unroll
      for (int i = 0; i < kChunkBytes / 16 / kTcThreads; ++i) {
        const int idx = tid + i * kTcThreads;
        const float4 x = r4[idx];
        float4 h, l;
        h.x = rna_tf32(x.x);
        h.y = rna_tf32(x.y);
        h.z = rna_tf32(x.z);
        h.w = rna_tf32(x.w);
        l.x = rna_tf32(x.x - h.x);
        l.y = rna_tf32(x.y - h.y);
        l.z = rna_tf32(x.z - h.z);
        l.w = rna_tf32(x.w - h.w);
        r4[idx] = h;
        l4[idx] = l;
      }
    }
    fence_proxy_async_smem();
    tc_fence_before_sync();
    __syncthreads();
    if (tid == 0) {
      tc_fence_after_sync();
      const uint32_t a_hi = smem_u32(raw + (size_t)s * kChunkBytes);
      const uint32_t a_lo = smem_u32(lob + (size_t)(it & 1) * kChunkBytes);
      const uint32_t b_hi = smem_u32(bhi), b_lo = smem_u32(blo);
#pragma unroll
      for (int g = 0; g < NKG; ++g) {
        const uint32_t boff = (uint32_t)(c * NKG + g) * (NPAD / 8) * 256;
        const uint64_t dah = make_smem_desc(a_hi + g * 1024, KC * 128, 512, kLayoutSw128Base32);
        const uint64_t dal = make_smem_desc(a_lo + g * 1024, KC * 128, 512, kLayoutSw128Base32);
        const uint64_t dbh = make_smem_desc(b_hi + boff, kPlainLbo, kPlainSbo, kLayoutNone);
        const uint64_t dbl = make_smem_desc(b_lo + boff, kPlainLbo, kPlainSbo, kLayoutNone);
        mma_tf32(tmem, dal, dbh, kIdesc, !(c == 0 && g == 0));
        mma_tf32(tmem, dah, dbl, kIdesc, true);
        mma_tf32(tmem, dah, dbh, kIdesc, true);
      }
      mma_commit(&bar_done[s]);
      // refill the stage used by the previous item with the load that is NST-1 items ahead
      const long nxt = it + NST - 1;
      if (nxt < nitems) {
        if (it >= 1) {
          const long j = it - 1;
          mbar_wait(&bar_done[j % NST], (uint32_t)((j / NST) & 1));
        }
        issue_load(nxt);
      }
    }
    if (c == p.nchunk - 1) {
      // ---- epilogue of this tile
      mbar_wait(&bar_done[s], ph);
      tc_fence_after_sync();
      const long tile = blockIdx.x + ti * gridDim.x;
      const int g = (int)(tile / p.tiles_per_slab);
      const long m = (long)(tile - (long)g * p.tiles_per_slab) * 128 + warp * 32 + lane;
      const bool in_range = m < p.mext;
      const bool live = m < p.valid_m;
      float* po = p.out + (long)g * p.gso + m;
#pragma unroll 1
      for (int n0 = 0; n0 < NPAD; n0 += 32) {
        if (n0 >= p.nout) break;
        float v[32];
        tmem_ld32(tmem + ((uint32_t)(warp * 32) << 16) + n0, v);
        if (in_range) {
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            const int n = n0 + j;
            if (n < p.nout) {
              float val = v[j] + sbias[n];
              if (p.act == 1) val = selu_f(val);
              float* q = po + (long)n * p.ldo;
              if (p.epi == 1) {
                if (live) *q += val;
              } else {
                *q = live ? val : 0.f;
              }
            }
          }
        }
      }
      tc_fence_before_sync();  // ordered before the next tile's first MMA by the __syncthreads of the next item
    }
  }
  tc_fence_before_sync();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, kTmemCols);
}

// ------------------------------------------------------------------------------------------------ host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  });
  return fn;
}

int encode_tensor_map(CUtensorMap* out, const float* base, int rank, const uint64_t* dims, const uint64_t* strides,
                      const uint32_t* box, int swizzle) {
  EncodeTiledFn fn = encode_fn();
  HNO_CHECK(fn != nullptr, "cuTensorMapEncodeTiled is not available from the CUDA driver");
  cuuint64_t d[5], s[5];
  cuuint32_t b[5], e[5];
  for (int i = 0; i < rank; ++i) {
    d[i] = dims[i];
    b[i] = box[i];
    e[i] = 1;
    if (i + 1 < rank) s[i] = strides[i];
  }
  const CUtensorMapSwizzle sw = swizzle == 2   ? CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B
                                : swizzle == 1 ? CU_TENSOR_MAP_SWIZZLE_128B
                                               : CU_TENSOR_MAP_SWIZZLE_NONE;
  const CUresult rc = fn(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, (cuuint32_t)rank, const_cast<float*>(base), d, s, b, e,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  HNO_CHECK(rc == CUDA_SUCCESS, "cuTensorMapEncodeTiled failed (CUresult %d)", (int)rc);
  return 0;
}

static std::atomic<int> g_tc_enabled{1};
int tc_set_enabled(int on) { return g_tc_enabled.exchange(on ? 1 : 0); }
bool tc_enabled() { return g_tc_enabled.load() != 0; }

bool tc_stream_eligible(const TcStreamArgs& a) {
  if (!tc_enabled()) return false;
  if (a.nsrc < 1 || a.nsrc > 2) return false;
  for (int i = 0; i < a.nsrc; ++i) {
    if (reinterpret_cast<uintptr_t>(a.a[i]) % 16) return false;
    if (a.lda[i] % 4 || a.gsa[i] % 4) return false;
  }
  if (a.mext < 1 || a.mext >= (1L << 31) || a.G < 1) return false;
  const int kc = a.kc;
  if (kc != 24 && kc != 32 && kc != 8) return false;
  if (a.nout > 256) return false;
  const int npad = a.nout <= 32 ? 32 : (a.nout <= 128 ? 128 : 256);
  const int nchunk = a.chunks_per_src * a.nsrc;
  if ((size_t)2 * npad * nchunk * kc * 4 > 96 * 1024) return false;
  return true;
}

template <int KC, int NPAD, int NST>
static int launch_t(const TcStreamArgs& a, cudaStream_t st) {
  CUtensorMap tm[2];
  for (int i = 0; i < 2; ++i) {
    const int j = i < a.nsrc ? i : 0;
    const uint64_t dims[3] = {(uint64_t)a.mext, (uint64_t)a.rows[j], (uint64_t)a.G};
    const uint64_t strides[2] = {(uint64_t)a.lda[j] * 4, (uint64_t)a.gsa[j] * 4};
    const uint32_t box[3] = {32, (uint32_t)KC, 1};
    if (int rc = encode_tensor_map(&tm[i], a.a[j], 3, dims, strides, box, 2)) return rc;
  }
  TcDev p;
  p.b = a.b;
  p.ldbn = a.ldbn;
  p.ldbk = a.ldbk;
  p.nvalid = a.nout;
  p.kvalid = a.kvalid;
  p.scale = a.scale;
  p.bias = a.bias;
  p.out = a.out;
  p.ldo = a.ldo;
  p.gso = a.gso;
  p.nout = a.nout;
  p.mext = a.mext;
  p.valid_m = a.valid_m;
  p.chunks_per_src = a.chunks_per_src;
  p.nchunk = a.chunks_per_src * a.nsrc;
  p.tiles_per_slab = ceil_div(a.mext, 128);
  p.total_tiles = (long)p.tiles_per_slab * a.G;
  p.act = a.act;
  p.epi = a.epi;
  const size_t smem = TcSmem<KC, NPAD, NST>::bytes(p.nchunk);
  auto kern = k_tc_stream<KC, NPAD, NST>;
  HNO_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int per_sm = (int)((227 * 1024) / (smem + 1024));
  if (per_sm < 1) per_sm = 1;
  if (per_sm > 512 / (NPAD < 32 ? 32 : NPAD)) per_sm = 512 / (NPAD < 32 ? 32 : NPAD);
  if (per_sm > 4) per_sm = 4;
  long grid = (long)sm_count() * per_sm;
  if (grid > p.total_tiles) grid = p.total_tiles;
  kern<<<(int)grid, kTcThreads, smem, st>>>(tm[0], tm[1], p);
  HNO_LAUNCH_CHECK();
  return 0;
}

int tc_stream_launch(const TcStreamArgs& a, cudaStream_t st) {
  HNO_CHECK(tc_stream_eligible(a), "tc_stream: configuration is not eligible for the tensor-core path");
  const int npad = a.nout <= 32 ? 32 : (a.nout <= 128 ? 128 : 256);
#define HNO_TC_CASE(KC_, NP_)                                     \
  if (a.kc == KC_ && npad == NP_) return launch_t<KC_, NP_, 3>(a, st);
  HNO_TC_CASE(24, 32)
  HNO_TC_CASE(32, 32)
  HNO_TC_CASE(8, 32)
  HNO_TC_CASE(24, 128)
  HNO_TC_CASE(32, 128)
  HNO_TC_CASE(8, 128)
  HNO_TC_CASE(24, 256)
  HNO_TC_CASE(32, 256)
  HNO_TC_CASE(8, 256)
#undef HNO_TC_CASE
  set_error("tc_stream: no kernel instance for kc=%d npad=%d", a.kc, npad);
  return -1;
}

}  // namespace hno
