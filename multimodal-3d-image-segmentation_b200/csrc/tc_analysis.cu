// Analysis stages of the truncated DHT (D axis: 121 samples -> 21 cos/sin rows; H axis: 121 -> 29) on tcgen05 with the
// streamed operand fed through TENSOR MEMORY.                       reference: nets/hnosegxs.py:378-410 via nets/dht.py:16-36
//
//   out[g][n][m] = sum_k A[g][k][m] * Bm[n][k]          m contiguous (plane columns), K = axis samples, N = J <= 32 rows
//
// Same contract as tc_stream.cu for the (kc = 16, nout <= 32, one source) instances; different data path.  In the shared-
// memory-operand kernel the streamed operand crosses shared memory five times per byte (TMA write, split read, lo write,
// two MMA operand reads); cycle counters (profiles/r2f_prof.log) show its worker warps busy ~725 cycles per 8 KB chunk with
// two CTAs per SM -- the shared-memory port, not HBM (the D stage ran at 3.0 TB/s).  Here every byte crosses shared memory
// twice and never as an MMA operand:
//   * warp 8 (one thread) keeps an 8-deep ring of plain [16 k][128 m] TMA boxes in flight (512-byte rows move at HBM speed);
//   * worker thread t owns voxel m0 + t = TMEM lane t: 16 conflict-free LDS.32 bring its k-values into registers, it
//     splits them (hi = the fp32 word, lo = x - trunc_tf32(x)) and writes both to one of four A slots in tensor memory
//     (tcgen05.st), releasing the ring stage as soon as the values are in registers;
//   * warp 9 (one thread) issues per 8-wide k-step  A_hi [B_hi | B_lo]  (N = 64) and  A_lo B_hi  (N = 32, first half) with A
//     from TMEM and the resident cos/sin image from shared memory (2 KB per instruction), into one of two accumulators;
//   * warps 4-7 run the epilogue of tile t (add the two halves, re-map m = h W + w -> [h][jd][w] for the D stage, store)
//     while tile t + 1 streams.
// Precision: 3xTF32 as everywhere.
#include "common.cuh"
#include "tc_common.cuh"
#include "tc_stream.h"

#include <stdio.h>
#include <stdlib.h>

namespace hno {

using namespace tc;

constexpr int kAnKC = 16;       // k rows per ring stage = values per thread and chunk (two MMA k-steps)
constexpr int kAnNP = 32;       // padded output rows
constexpr int kAnNB = 64;       // rows of the fused B image [hi | lo] = accumulator columns per buffer
constexpr int kAnSlots = 4;     // A slots in tensor memory: 4 x (16 hi + 16 lo) columns
constexpr int kAnThreads = 320; // warps 0-3 workers, 4-7 epilogue, 8 TMA producer, 9 MMA issuer
constexpr int kAnStageBytes = kAnKC * 512;
constexpr uint32_t kAnACols = kAnSlots * 2 * kAnKC;           // 128
constexpr uint32_t kAnTmemCols = 256;                          // A slots + 2 x 64 accumulator columns

struct AnDev {
  const float* b;
  long ldbn, ldbk;
  int nvalid, kvalid;
  float scale;
  float* out;
  long ldo, gso;
  int nout;
  int mext, valid_m;
  int nchunk;
  int tiles_per_slab, total_tiles;
  int nst;
  int out_rw;
  long out_rp;
};

__device__ __forceinline__ void an_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void an_st16(uint32_t taddr, const uint32_t* r) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};" ::"r"(taddr),
      "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
      "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
      : "memory");
}
__device__ __forceinline__ void an_mma(uint32_t tmem_d, uint32_t tmem_a, uint64_t desc_b, uint32_t idesc, bool accumulate) {
  const uint32_t acc = accumulate ? 1u : 0u;
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}" ::"r"(tmem_d),
      "r"(tmem_a), "l"(desc_b), "r"(idesc), "r"(acc)
      : "memory");
}

constexpr int kAnMaxStages = 10;

__global__ void __launch_bounds__(kAnThreads, 2) k_tc_analysis(const __grid_constant__ CUtensorMap tm, const AnDev p) {
  constexpr uint32_t kIdescB = make_idesc_tf32(128, kAnNB, 0, 0);
  constexpr uint32_t kIdesc = make_idesc_tf32(128, kAnNP, 0, 0);
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  const int NST = p.nst;
  uint8_t* ring = smem;                                               // [NST][16][128] fp32
  float* bimg = reinterpret_cast<float*>(ring + NST * kAnStageBytes);  // [64 rows][ktot] K-major core-matrix image
  const int ktot = p.nchunk * kAnKC;
  __shared__ __align__(8) uint64_t bar_full[kAnMaxStages];   // TMA bytes landed                 (1 arrival + tx)
  __shared__ __align__(8) uint64_t bar_empty[kAnMaxStages];  // stage read into registers       (128 arrivals)
  __shared__ __align__(8) uint64_t bar_ready[kAnSlots];      // A slot written                   (128 arrivals)
  __shared__ __align__(8) uint64_t bar_free[kAnSlots];       // MMAs that read the slot retired  (tcgen05.commit)
  __shared__ __align__(8) uint64_t bar_accfull[2];           // all MMAs of a tile retired       (tcgen05.commit)
  __shared__ __align__(8) uint64_t bar_accfree[2];           // accumulator drained              (128 arrivals)
  __shared__ uint32_t tmem_slot;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  {  // resident operand: rows [0, 32) = hi, [32, 64) = lo; coalesced along whichever index of B is contiguous
    constexpr int kNW = kAnThreads / 32;
    const bool k_contig = p.ldbk == 1;
    const int n_outer = k_contig ? kAnNP : ktot, n_inner = k_contig ? ktot : kAnNP;
    for (int o = warp; o < n_outer; o += kNW)
      for (int i = lane; i < n_inner; i += 32) {
        const int n = k_contig ? o : i, k = k_contig ? i : o;
        float v = 0.f;
        if (n < p.nvalid && k < p.kvalid) v = p.scale * __ldg(p.b + (long)n * p.ldbn + (long)k * p.ldbk);
        const float hi = __uint_as_float((__float_as_uint(v) + 0x1000u) & 0xFFFFE000u);
        bimg[kmajor_plain_index<kAnNB>(n, k)] = hi;
        bimg[kmajor_plain_index<kAnNB>(kAnNP + n, k)] = v - hi;
      }
  }
  if (tid == 0) {
    for (int s = 0; s < kAnMaxStages; ++s) {
      mbar_init(&bar_full[s], 1);
      mbar_init(&bar_empty[s], 128);
    }
    for (int s = 0; s < kAnSlots; ++s) {
      mbar_init(&bar_ready[s], 128);
      mbar_init(&bar_free[s], 1);
    }
    mbar_init(&bar_accfull[0], 1);
    mbar_init(&bar_accfull[1], 1);
    mbar_init(&bar_accfree[0], 128);
    mbar_init(&bar_accfree[1], 128);
    mbar_fence_init();
  }
  if (warp == 8) tmem_alloc(&tmem_slot, kAnTmemCols);
  fence_proxy_async_smem();
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem = tmem_slot;
  asm volatile("griddepcontrol.wait;" ::: "memory");  // the previous kernel's results are read from here on

  const int my_tiles = p.total_tiles > (int)blockIdx.x ? (p.total_tiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
  const int nchunk = p.nchunk;
  const int items = my_tiles * nchunk;

  if (warp == 8) {
    // =============================================================== TMA producer (one thread)
    if (lane == 0) {
      tma_prefetch_desc(&tm);
      int s = 0, it = 0;
      uint32_t ph = 0;
      for (int ti = 0; ti < my_tiles; ++ti) {
        const uint32_t tile = blockIdx.x + (uint32_t)ti * gridDim.x;
        const int g = tile / (uint32_t)p.tiles_per_slab;
        const int m0 = (tile - (uint32_t)g * p.tiles_per_slab) * 128;
        for (int c = 0; c < nchunk; ++c, ++it) {
          if (it >= NST) mbar_wait(&bar_empty[s], ph ^ 1);
          mbar_expect_tx(&bar_full[s], kAnStageBytes);
          tma_load_3d(ring + s * kAnStageBytes, &tm, m0, c * kAnKC, g, &bar_full[s]);
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
      const uint32_t b0 = smem_u32(bimg);
      int it = 0;
      for (int ti = 0; ti < my_tiles; ++ti) {
        const int buf = ti & 1;
        if (ti >= 2) mbar_wait(&bar_accfree[buf], (uint32_t)(((ti >> 1) - 1) & 1));
        tc_fence_after_sync();
        const uint32_t acc = tmem + kAnACols + buf * kAnNB;
        for (int c = 0; c < nchunk; ++c, ++it) {
          const int slot = it % kAnSlots;
          mbar_wait(&bar_ready[slot], (uint32_t)((it / kAnSlots) & 1));
          tc_fence_after_sync();
          const uint32_t a_hi = tmem + slot * (2 * kAnKC), a_lo = a_hi + kAnKC;
#pragma unroll
          for (int g = 0; g < kAnKC / 8; ++g) {
            const uint32_t boff = (uint32_t)(c * (kAnKC / 8) + g) * (kAnNB / 8) * 256;
            const uint64_t db = make_smem_desc(b0 + boff, kPlainLbo, kPlainSbo, kLayoutNone);
            an_mma(acc, a_hi + 8 * g, db, kIdescB, !(c == 0 && g == 0));  // [hi*hi | hi*lo]
            an_mma(acc, a_lo + 8 * g, db, kIdesc, true);                   // lo*hi into the first half
          }
          mma_commit(&bar_free[slot]);
          if (c == nchunk - 1) mma_commit(&bar_accfull[buf]);
        }
      }
    }
    __syncwarp();
  } else if (warp < 4) {
    // =============================================================== workers: ring -> registers -> tensor memory
    const uint32_t lane_base = tmem + ((uint32_t)(warp * 32) << 16);
    int s = 0;
    uint32_t ph = 0;
    for (int it = 0; it < items; ++it) {
      const int slot = it % kAnSlots;
      mbar_wait(&bar_full[s], ph);
      const float* src = reinterpret_cast<const float*>(ring + s * kAnStageBytes) + tid;
      uint32_t hi[kAnKC], lo[kAnKC];
#pragma unroll
      for (int k = 0; k < kAnKC; ++k) {
        const float x = src[k * 128];
        hi[k] = __float_as_uint(x);
        lo[k] = __float_as_uint(tf32_lo(x));
      }
      if (it >= kAnSlots) mbar_wait(&bar_free[slot], (uint32_t)(((it / kAnSlots) - 1) & 1));
      tc_fence_after_sync();
      const uint32_t a0 = lane_base + slot * (2 * kAnKC);
      an_st16(a0, hi);
      an_st16(a0 + kAnKC, lo);
      an_arrive(&bar_empty[s]);  // every value of the stage has been consumed by the stores above
      asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
      tc_fence_before_sync();
      an_arrive(&bar_ready[slot]);
      if (++s == NST) {
        s = 0;
        ph ^= 1;
      }
    }
  } else {
    // =============================================================== epilogue (warps 4-7: TMEM lane quarter warp - 4)
    const int quarter = warp - 4;
    for (int ti = 0; ti < my_tiles; ++ti) {
      const uint32_t tile = blockIdx.x + (uint32_t)ti * gridDim.x;
      const int g = tile / (uint32_t)p.tiles_per_slab;
      const int m = (tile - (uint32_t)g * p.tiles_per_slab) * 128 + quarter * 32 + lane;
      const bool live = m < p.valid_m;
      const bool in_range = p.out_rw > 0 ? (m < p.mext && live) : m < p.mext;  // re-mapped: dead columns have no address
      long moff = m;
      if (p.out_rw > 0) {
        const int hh = m / p.out_rw;
        moff = in_range ? (long)hh * p.out_rp + (m - hh * p.out_rw) : 0;
      }
      float* po = p.out + (long)g * p.gso + moff;
      const int buf = ti & 1;
      mbar_wait(&bar_accfull[buf], (uint32_t)((ti >> 1) & 1));
      tc_fence_after_sync();
      const uint32_t acc = tmem + kAnACols + buf * kAnNB + ((uint32_t)(quarter * 32) << 16);
      float v[32], v2[32];
      tmem_ld32(acc, v);
      tmem_ld32(acc + kAnNP, v2);
      tc_fence_before_sync();
      an_arrive(&bar_accfree[buf]);
      if (in_range) {
#pragma unroll
        for (int j = 0; j < 32; ++j)
          if (j < p.nout) po[(long)j * p.ldo] = live ? v[j] + v2[j] : 0.f;
      }
    }
  }
  tc_fence_before_sync();
  __syncthreads();
  if (warp == 8) tmem_dealloc(tmem, kAnTmemCols);
}

bool tc_analysis_eligible(const TcStreamArgs& a) {
  static const bool on = !(getenv("HNO_TC_ANALYSIS") && atoi(getenv("HNO_TC_ANALYSIS")) == 0);
  if (!on || !tc_enabled()) return false;
  if (a.nsrc != 1 || a.kc != kAnKC || a.nout < 1 || a.nout > kAnNP || a.in_rw != 0 || a.out_rw < 0) return false;
  if (a.epi != 0 || a.act != 0 || a.bias != nullptr) return false;
  if (reinterpret_cast<uintptr_t>(a.a[0]) % 16 || a.lda[0] % 4 || a.gsa[0] % 4 || a.mext % 4) return false;
  if (a.mext < 1 || a.mext >= (1L << 30) || a.G < 1 || (a.mext + 127) / 128 * a.G >= (1L << 30)) return false;
  if ((size_t)kAnNB * a.chunks_per_src * kAnKC * 4 > 64 * 1024) return false;
  return true;
}

int tc_analysis_launch(const TcStreamArgs& a, cudaStream_t st) {
  HNO_CHECK(tc_analysis_eligible(a), "tc_analysis: configuration is not eligible");
  CUtensorMap tm;
  {
    const uint64_t dims[3] = {(uint64_t)a.mext, (uint64_t)a.rows[0], (uint64_t)a.G};
    const uint64_t strides[2] = {(uint64_t)a.lda[0] * 4, (uint64_t)a.gsa[0] * 4};
    const uint32_t box[3] = {128, kAnKC, 1};
    if (int rc = encode_tensor_map(&tm, a.a[0], 3, dims, strides, box, 0)) return rc;
  }
  AnDev p;
  p.b = a.b, p.ldbn = a.ldbn, p.ldbk = a.ldbk;
  p.nvalid = a.nout, p.kvalid = a.kvalid < a.rows[0] ? a.kvalid : a.rows[0];
  p.scale = a.scale;
  p.out = a.out, p.ldo = a.ldo, p.gso = a.gso, p.nout = a.nout;
  p.mext = (int)a.mext;
  p.valid_m = (int)(a.valid_m < a.mext ? a.valid_m : a.mext);
  p.nchunk = a.chunks_per_src;
  p.tiles_per_slab = ceil_div(a.mext, 128);
  p.total_tiles = p.tiles_per_slab * a.G;
  p.out_rw = a.out_rw, p.out_rp = a.out_rp;
  const size_t bimg = (size_t)kAnNB * p.nchunk * kAnKC * 4;
  // two CTAs per SM: 228 KB - 2 x (1 KB reserved + static) -> ring depth from what the B image leaves
  int nst = (int)((110 * 1024 - 1024 - bimg) / kAnStageBytes);
  static const int nst_env = getenv("HNO_TC_ANALYSIS_NST") ? atoi(getenv("HNO_TC_ANALYSIS_NST")) : 0;
  if (nst_env > 0) nst = nst_env;
  if (nst > kAnMaxStages) nst = kAnMaxStages;
  HNO_CHECK(nst >= 2, "tc_analysis: the basis image leaves no room for the ring");
  p.nst = nst;
  const size_t smem = 1024 + (size_t)nst * kAnStageBytes + bimg;
  HNO_CUDA(cudaFuncSetAttribute(k_tc_analysis, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  HNO_CUDA(cudaFuncSetAttribute(k_tc_analysis, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
  long grid = (long)sm_count() * 2;
  if (grid > p.total_tiles) grid = p.total_tiles;
  static const bool pdl = !(getenv("HNO_TC_PDL") && atoi(getenv("HNO_TC_PDL")) == 0);
  if (pdl) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)grid);
    cfg.blockDim = dim3(kAnThreads);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    HNO_CUDA(cudaLaunchKernelEx(&cfg, k_tc_analysis, tm, p));
  } else {
    k_tc_analysis<<<(int)grid, kAnThreads, smem, st>>>(tm, p);
  }
  HNO_LAUNCH_CHECK();
  return 0;
}

}  // namespace hno
