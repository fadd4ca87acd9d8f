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
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <mutex>

namespace hno {

using namespace tc;

// Worker warps (operand split + epilogue): 4 for narrow outputs; 8 for NPAD >= 128, where the epilogue (up to 121 rows
// of SELU + stores per voxel) is the critical path: warps w and w + 4 share a TMEM lane quarter and take alternating
// 32-column blocks of the accumulator.  Two more warps follow: producer (cp.async or TMA) and MMA issuer.
template <int NPAD>
struct TcShape {
  // measured on B200: 8 warps do not speed up the store/SELU epilogue (instruction bound, dhts 0.138 ms either way) and
  // cost the accumulate epilogue its second prefetch block (dhta 0.196 -> 0.292 ms), so every shape runs with 4
#ifndef HNO_TC_WIDE_WORKER_WARPS
#define HNO_TC_WIDE_WORKER_WARPS 8
#endif
  static constexpr int kWorkerWarps = NPAD == 128 ? HNO_TC_WIDE_WORKER_WARPS : 4;
  static constexpr int kWorkers = 32 * kWorkerWarps;
  // Wide outputs (the synthesis stages: up to 121 rows per voxel) leave through shared memory and bulk tensor stores
  // issued by a seventh warp: one STS per element instead of one STG with 64-bit address arithmetic, and the
  // accumulate form becomes a bulk reduce-add (the old values are never loaded by the SM).
  static constexpr bool kTmaOut = NPAD >= 128;
#ifndef HNO_TC_STAGE_BUFS
#define HNO_TC_STAGE_BUFS 3
#endif
  // staging buffers of 32 rows x 128 voxels.  Three, with up to two bulk stores in flight behind the one being filled: the
  // issuer used to wait for every store to finish READING its buffer before it even looked at the next one, so stores
  // were serialised with a bubble and the workers waited ~930 cycles per block for a free buffer (cycle counters,
  // profiles/r2g_prof2.log).  128-row instances afford it (one lo buffer less), 256-row ones keep two.
  static constexpr int kStageBufs = kTmaOut ? (NPAD <= 128 ? HNO_TC_STAGE_BUFS : 2) : 0;
  static constexpr int kStageBytes = 32 * 128 * 4;
  // Narrow outputs (pointwise convolutions, analysis stages): the operand split of tile t+1 and the epilogue of tile t
  // run on DIFFERENT warps (4 split warps + 4 epilogue warps) instead of one after the other on the same four -- the
  // worker warps were the critical path (cycle counters: split 2240 + epilogue 2837 of 6375 cycles per tile of the
  // 48 -> 24 convolution).  HNO_TC_SPLIT_EPI=0 at compile time restores the single worker group.
#ifndef HNO_TC_SPLIT_EPI
#define HNO_TC_SPLIT_EPI 1
#endif
  static constexpr bool kSplitEpi = HNO_TC_SPLIT_EPI && NPAD <= 32;
  static constexpr int kEpiWarps = kSplitEpi ? 4 : 0;
  static constexpr int kHelper0 = kWorkerWarps + kEpiWarps;  // first helper warp: producer, then MMA issuer, then store
  static constexpr int kThreads = kWorkers + 32 * kEpiWarps + 64 + (kTmaOut ? 32 : 0);
};

struct TcDev {
  const float* b;
  long ldbn, ldbk;
  int nvalid, kvalid;
  float scale;
  const float* bias;
  float* out;
  long ldo, gso;
  int nout;
  int mext, valid_m;
  int nchunk, chunks_per_src;
  int tiles_per_slab;
  int total_tiles;
  int act, epi;
  int prefetch_items;  // L2 prefetch distance of the producer, in chunks (0 = off)
  long long* prof;     // debug (HNO_TC_PROF=1): per-CTA cycle counters [grid][8], else null
  int prof_mode;       // HNO_TC_PROF=2: slots 4..7 = staged epilogue breakdown (tmem ld, wait free, compute + STS, fence)
  // streamed operand as raw pointers (LDGSTS loader); the TMA loader uses the tensor maps instead
  const float* a[2];
  long lda[2], gsa[2];
  int rows[2];
  int loader;          // 0 = cp.async (LDGSTS) warp, 1 = TMA (one thread)
  int tma_out;         // epilogue through shared memory + bulk tensor store / reduce-add (NPAD >= 128 instances)
  int out_rw, in_rw;   // row re-mapping of the contiguous axis (tc_stream.h), 0 = off
  long out_rp, in_rp;
};

template <int KC, int NPAD, int NST, int kNLo>
struct TcSmem {
  static constexpr int kChunkBytes = KC * 512;  // 4 blocks of 32 m x KC rows x 4 B
  static size_t bytes(int nchunk) {
    return 1024 /* alignment slack */ + (size_t)(NST + kNLo) * kChunkBytes + (size_t)2 * NPAD * nchunk * KC * 4 +
           NPAD * 4 + (size_t)TcShape<NPAD>::kStageBufs * TcShape<NPAD>::kStageBytes;
  }
};

__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// round-to-nearest TF32 "hi" part with two integer ops (half-ulp add, mask) -- the F2F conversion unit is an
// eighth-rate pipe and would dominate the operand split
__device__ __forceinline__ float rn_tf32_bits(float x) {
  return __uint_as_float((__float_as_uint(x) + 0x1000u) & 0xFFFFE000u);
}

template <int KC, int NPAD, int NST, int kNLo>
__global__ void __launch_bounds__(TcShape<NPAD>::kThreads,
                                  (NPAD <= 32 ? (TcShape<NPAD>::kSplitEpi ? 2 : 3) : (NPAD <= 128 ? 2 : 1))) k_tc_stream(const __grid_constant__ CUtensorMap tm0,
                                                           const __grid_constant__ CUtensorMap tm1,
                                                           const __grid_constant__ CUtensorMap tmo, const TcDev p) {
  constexpr int kChunkBytes = KC * 512;
  constexpr int NKG = KC / 8;
  constexpr int kTcWorkers = TcShape<NPAD>::kWorkers, kTcThreads = TcShape<NPAD>::kThreads;
  constexpr int kWW = TcShape<NPAD>::kHelper0;      // producer = warp kWW, MMA issuer = warp kWW + 1
  constexpr int kHalves = TcShape<NPAD>::kWorkerWarps / 4;  // worker warps per TMEM lane quarter
  constexpr bool kSplitEpi = TcShape<NPAD>::kSplitEpi;
  // kFuseN: A_hi * [B_hi | B_lo] as ONE MMA of N = 2 * NPAD (the two halves are added in the epilogue) plus A_lo * B_hi
  // into the first half: the streamed operand is read from shared memory twice per k-step instead of three times
  // (shared-memory bandwidth, not the tensor pipe, is what bounds this kernel once HBM is fed properly).
#ifndef HNO_TC_FUSEN_KC16
#define HNO_TC_FUSEN_KC16 1
#endif
  // measured slower on the 48-row pointwise convolution (pw48f 0.169 -> 0.200 ms) but the 121-row analysis stages are
  // bound by shared-memory bandwidth (27 KB cross it per 4 KB streamed), where one read less of A per k-step pays
  constexpr bool kFuseN = HNO_TC_FUSEN_KC16 && KC == 16 && NPAD == 32;
  constexpr int NB = kFuseN ? 2 * NPAD : NPAD;  // rows of the resident B image / accumulator columns per buffer
  constexpr uint32_t kIdesc = make_idesc_tf32(128, NPAD, 1, 0);
  constexpr uint32_t kIdesc2 = make_idesc_tf32(128, NB, 1, 0);
  constexpr uint32_t kTmemCols = 2 * NB < 32 ? 32 : 2 * NB;  // two accumulator buffers
  static_assert(KC % 8 == 0 && NPAD % 16 == 0 && NPAD <= 256, "bad tile configuration");
  // Programmatic dependent launch: let the NEXT kernel on the stream begin (its prologue overlaps this kernel's tail) ...
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  // align on the shared-window address so that the compiler keeps the shared address space (LDS / STS)
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* raw = smem;                              // [NST][kChunkBytes]
  uint8_t* lob = raw + NST * kChunkBytes;           // [kNLo][kChunkBytes]
  // loader 2 ("plain TMA"): the ring holds unswizzled [k][128 m] boxes (512-byte rows: the TMA unit moves those at
  // HBM speed, 128-byte-row swizzled boxes cap at 4.2 TB/s and a cp.async warp at ~3 TB/s with 2 CTAs per SM) and the
  // split step writes BOTH operand images (hi, lo) in the swizzled layout the MMA reads.  hib: [kNLo][kChunkBytes],
  // like every operand buffer a multiple of 512 bytes from the 1024-byte aligned base (the swizzle uses address bits).
  uint8_t* hib = lob + kNLo * kChunkBytes;
  float* bhi = reinterpret_cast<float*>(hib + (p.loader == 2 ? kNLo * kChunkBytes : 0));
  const int ktot = p.nchunk * KC;
  float* blo = bhi + NPAD * ktot;
  float* sbias = blo + NPAD * ktot;
  constexpr bool kTmaOut = TcShape<NPAD>::kTmaOut;
  // staging buffers of the bulk-store epilogue: the B image and the bias occupy a multiple of 128 bytes, so these stay
  // 128-byte aligned
  float* stage = sbias + NPAD;
  __shared__ __align__(8) uint64_t bar_read[NST];   // loader 2: ring stage read by the split warps (128 arrivals)
  constexpr int kSB = TcShape<NPAD>::kStageBufs > 0 ? TcShape<NPAD>::kStageBufs : 1;
  __shared__ __align__(8) uint64_t bar_stfull[kSB];   // staging buffer written by the workers   (128 arrivals)
  __shared__ __align__(8) uint64_t bar_stfree[kSB];   // bulk store has read the staging buffer (1 arrival)
  __shared__ __align__(8) uint64_t bar_full[NST];   // TMA bytes landed                     (1 arrival + tx)
  __shared__ __align__(8) uint64_t bar_split[NST];  // hi / lo operands ready                (128 arrivals)
  __shared__ __align__(8) uint64_t bar_done[NST];   // MMAs of the item retired              (tcgen05.commit)
  __shared__ __align__(8) uint64_t bar_accfree[2];  // accumulator buffer drained by epilogue (128 arrivals)
  __shared__ __align__(8) uint64_t bar_accfull[2];  // all MMAs of a tile retired            (tcgen05.commit)
  __shared__ uint32_t tmem_slot;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  // ---- resident operand: B image (hi / lo), bias
  {
    // no divisions, coalesced along whichever index of B is contiguous in memory (every CTA of every launch pays this
    // prologue: 1.5 tiles per CTA in the H stages of the transform)
    constexpr int kNW = kTcThreads / 32;
    const bool k_contig = p.ldbk == 1;
    const int n_outer = k_contig ? NPAD : ktot, n_inner = k_contig ? ktot : NPAD;
    for (int o = tid >> 5; o < n_outer; o += kNW)
      for (int i = tid & 31; i < n_inner; i += 32) {
        const int n = k_contig ? o : i, k = k_contig ? i : o;
        // K axis of the image = chunks_per_src * KC rows per source; a source with fewer rows (12-channel tensors in
        // 16-row chunks) leaves a gap the TMA zero-fills, so image column k maps to column src * rows[0] + r of B
        const int kps = p.chunks_per_src * KC;
        const int src = k >= kps ? 1 : 0, r = k - src * kps;
        const int col = src * p.rows[0] + r;
        float v = 0.f;
        if (n < p.nvalid && r < p.rows[src] && col < p.kvalid) v = p.scale * __ldg(p.b + (long)n * p.ldbn + (long)col * p.ldbk);
        const float hi = rn_tf32_bits(v);
        if (kFuseN) {  // one image: rows [0, NPAD) = hi, rows [NPAD, 2 NPAD) = lo
          bhi[kmajor_plain_index<NB>(n, k)] = hi;
          bhi[kmajor_plain_index<NB>(NPAD + n, k)] = v - hi;
        } else {
          const int idx = kmajor_plain_index<NPAD>(n, k);
          bhi[idx] = hi;
          blo[idx] = v - hi;
        }
      }
  }
  for (int n = tid; n < NPAD; n += kTcThreads) sbias[n] = (p.bias != nullptr && n < p.nout) ? __ldg(p.bias + n) : 0.f;
  if (tid == 0) {
#pragma unroll
    for (int s = 0; s < NST; ++s) {
      mbar_init(&bar_full[s], p.loader != 0 ? 1 : 32);
      mbar_init(&bar_split[s], kTcWorkers);
      mbar_init(&bar_done[s], 1);
      mbar_init(&bar_read[s], kTcWorkers);
    }
    mbar_init(&bar_accfree[0], kTcWorkers);
    mbar_init(&bar_accfree[1], kTcWorkers);
    mbar_init(&bar_accfull[0], 1);
    mbar_init(&bar_accfull[1], 1);
#pragma unroll
    for (int sb = 0; sb < kSB; ++sb) {
      mbar_init(&bar_stfull[sb], kTcWorkers);
      mbar_init(&bar_stfree[sb], 1);
    }
    mbar_fence_init();
  }
  if (warp == kWW) tmem_alloc(&tmem_slot, kTmemCols);
  fence_proxy_async_smem();
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem = tmem_slot;
  // ... and wait here, after the prologue (B image from constants, barriers, TMEM), for the PREVIOUS kernel's results:
  // nothing above reads or writes data another kernel of the step produces.  A no-op without the launch attribute.
  asm volatile("griddepcontrol.wait;" ::: "memory");

  const int my_tiles = p.total_tiles > (int)blockIdx.x ? (p.total_tiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
  const int nchunk = p.nchunk;

  if (warp == kWW && p.loader == 0) {
    // =============================================================== LDGSTS producer (one warp)
    // TMA tile loads of 128-byte-wide boxes (the widest a swizzled MN-major tf32 operand allows) are limited by the TMA
    // unit to one box row per ~8.6 cycles per SM = 4.2 TB/s chip-wide (tools/ubench_tma.cu, profiles/r1b_ubench_tma.log);
    // 16-byte cp.async copies issued by one warp (512 contiguous bytes of one row per instruction) reach the HBM limit.
    // The warp writes the 128B_BASE32B swizzle pattern itself: the 32-byte chunk c of row r lands at chunk c ^ (r & 3).
    const int q = lane & 7, j = lane >> 3;  // 16-byte piece within a 128-byte row, 32-float column block
    int it = 0, s = 0;
    uint32_t ph = 0;
    for (int ti = 0; ti < my_tiles; ++ti) {
      const uint32_t tile = blockIdx.x + (uint32_t)ti * gridDim.x;
      const int g = tile / (uint32_t)p.tiles_per_slab;
      const int col = (tile - (uint32_t)g * p.tiles_per_slab) * 128 + j * 32 + q * 4;
      const bool col_ok = col < p.mext;
      // re-mapped source (in_rw > 0): 8-byte pieces, this lane copies pieces `lane` and `lane + 32` of every row
      long roff[2] = {0, 0};
      bool rok[2] = {false, false};
      uint32_t rdst[2] = {0, 0};
      if (p.in_rw > 0) {
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const int pc = lane + 32 * h;  // piece 0..63 of the 512-byte tile row
          const int m = (tile - (uint32_t)g * p.tiles_per_slab) * 128 + 2 * pc;
          rok[h] = m < p.mext && m < p.valid_m;
          const int hh = m / p.in_rw;
          roff[h] = rok[h] ? (long)hh * p.in_rp + (m - hh * p.in_rw) : 0;
          // 32-float block pc >> 4, 32-byte chunk (pc >> 2) & 3 (XOR-ed with the row below), 8-byte piece pc & 3
          rdst[h] = (uint32_t)(pc >> 4) * (KC * 128) + ((pc & 3) << 3);
        }
      }
      for (int c = 0; c < nchunk; ++c) {
        const int src = c / p.chunks_per_src;
        const int row0 = (c - src * p.chunks_per_src) * KC;
        if (it >= NST) mbar_wait(&bar_done[s], ph ^ 1);  // previous use of this stage fully consumed
        if (p.in_rw > 0) {
          const long ld = p.lda[src];
          const int nrow = p.rows[src] - row0;
          const float* gp = p.a[src] + (long)g * p.gsa[src] + (long)row0 * ld;
          const uint32_t base = smem_u32(raw + s * kChunkBytes);
#pragma unroll 4
          for (int r = 0; r < KC; ++r) {
#pragma unroll
            for (int h = 0; h < 2; ++h) {
              const int pc = lane + 32 * h;
              const uint32_t dst = base + rdst[h] + r * 128 + (((((pc >> 2) & 3) ^ r) & 3) << 5);
              const bool ok = rok[h] && r < nrow;
              const float* sp = ok ? gp + (long)r * ld + roff[h] : p.a[src];
              asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;" ::"r"(dst), "l"(sp), "r"(ok ? 8 : 0) : "memory");
            }
          }
          asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(smem_u32(&bar_full[s])) : "memory");
          ++it;
          if (++s == NST) {
            s = 0;
            ph ^= 1;
          }
          continue;
        }
        const float* gp = p.a[src] + (long)g * p.gsa[src] + (long)row0 * p.lda[src] + (col_ok ? col : 0);
        const long ld = p.lda[src];
        const int nrow = p.rows[src] - row0;  // rows of this chunk that exist (the rest reads as zero)
        const uint32_t dst0 = smem_u32(raw + s * kChunkBytes) + j * (KC * 128) + ((q & 1) << 4);
#pragma unroll 8
        for (int r = 0; r < KC; ++r) {
          const uint32_t dst = dst0 + r * 128 + ((((q >> 1) ^ r) & 3) << 5);
          const bool ok = col_ok && r < nrow;
          const float* sp = ok ? gp + (long)r * ld : p.a[src];
          asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(sp), "r"(ok ? 16 : 0) : "memory");
        }
        asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(smem_u32(&bar_full[s])) : "memory");
        ++it;
        if (++s == NST) {
          s = 0;
          ph ^= 1;
        }
      }
    }
    asm volatile("cp.async.wait_all;" ::: "memory");
  } else if (warp == kWW) {
    // =============================================================== TMA producer (one thread)
    if (lane == 0) {
      tma_prefetch_desc(&tm0);
      tma_prefetch_desc(&tm1);
      // (tile, chunk) cursor; `pf` runs kPrefetch items ahead of `ld` and only warms L2
      struct Cursor {
        int ti, c, src, cs, row0, g, m0;
      };
      auto locate = [&](Cursor& k) {
        const uint32_t tile = blockIdx.x + (uint32_t)k.ti * gridDim.x;
        k.g = tile / (uint32_t)p.tiles_per_slab;
        k.m0 = (tile - (uint32_t)k.g * p.tiles_per_slab) * 128;
      };
      auto advance = [&](Cursor& k) {
        k.row0 += KC;
        if (++k.cs == p.chunks_per_src) {
          k.cs = 0;
          k.row0 = 0;
          ++k.src;
        }
        if (++k.c == nchunk) {
          k.c = 0;
          k.src = 0;
          k.cs = 0;
          k.row0 = 0;
          ++k.ti;
          if (k.ti < my_tiles) locate(k);
        }
      };
      const int kPrefetch = p.prefetch_items;
      Cursor ld{0, 0, 0, 0, 0, 0, 0}, pf{0, 0, 0, 0, 0, 0, 0};
      long long prof_acc[1] = {0};
      const long long t_begin = p.prof ? clock64() : 0;
      if (my_tiles > 0) {
        locate(ld);
        locate(pf);
      }
      const bool plain = p.loader == 2;
      for (int i = 0; i < kPrefetch && pf.ti < my_tiles; ++i) {
        const CUtensorMap* tm = pf.src == 0 ? &tm0 : &tm1;
        if (plain) {
          tma_prefetch_l2_3d(tm, pf.m0, pf.row0, pf.g);
        } else {
#pragma unroll
          for (int j = 0; j < 4; ++j) tma_prefetch_l2_3d(tm, pf.m0 + 32 * j, pf.row0, pf.g);
        }
        advance(pf);
      }
      int it = 0, s = 0;
      uint32_t ph = 0;  // parity of the current use of stage s
      while (ld.ti < my_tiles) {
        if (pf.ti < my_tiles) {
          const CUtensorMap* tm = pf.src == 0 ? &tm0 : &tm1;
          if (plain) {
            tma_prefetch_l2_3d(tm, pf.m0, pf.row0, pf.g);
          } else {
#pragma unroll
            for (int j = 0; j < 4; ++j) tma_prefetch_l2_3d(tm, pf.m0 + 32 * j, pf.row0, pf.g);
          }
          advance(pf);
        }
        if (it >= NST) {  // previous use of this stage fully consumed (loader 2: read by the split warps)
          const long long t0 = p.prof ? clock64() : 0;
          mbar_wait(plain ? &bar_read[s] : &bar_done[s], ph ^ 1);
          if (p.prof) prof_acc[0] += clock64() - t0;
        }
        uint8_t* dst = raw + s * kChunkBytes;
        mbar_expect_tx(&bar_full[s], kChunkBytes);
        const CUtensorMap* tm = ld.src == 0 ? &tm0 : &tm1;
        if (plain) {
          tma_load_3d(dst, tm, ld.m0, ld.row0, ld.g, &bar_full[s]);
        } else {
#pragma unroll
          for (int j = 0; j < 4; ++j) tma_load_3d(dst + j * (KC * 128), tm, ld.m0 + 32 * j, ld.row0, ld.g, &bar_full[s]);
        }
        advance(ld);
        ++it;
        if (++s == NST) {
          s = 0;
          ph ^= 1;
        }
      }
      if (p.prof) {
        p.prof[blockIdx.x * 8 + 0] = prof_acc[0];
        p.prof[blockIdx.x * 8 + 1] = clock64() - t_begin;
      }
    }
    __syncwarp();
  } else if (kTmaOut && warp == kWW + 2) {
    // =============================================================== bulk-store issuer (one thread)
    if (lane == 0 && p.tma_out) {
      tma_prefetch_desc(&tmo);
      int sblk = 0;
      for (int ti = 0; ti < my_tiles; ++ti) {
        const uint32_t tile = blockIdx.x + (uint32_t)ti * gridDim.x;
        const int g = tile / (uint32_t)p.tiles_per_slab;
        const int m0 = (tile - (uint32_t)g * p.tiles_per_slab) * 128;
        for (int n0 = 0; n0 < p.nout; n0 += 32, ++sblk) {
          const int sb = sblk % kSB;
          mbar_wait(&bar_stfull[sb], (uint32_t)((sblk / kSB) & 1));
          const uint32_t src = smem_u32(stage + sb * (32 * 128));
          if (p.epi == 1)
            asm volatile("cp.reduce.async.bulk.tensor.3d.global.shared::cta.add.bulk_group [%0, {%2, %3, %4}], [%1];" ::"l"(&tmo),
                         "r"(src), "r"(m0), "r"(n0), "r"(g)
                         : "memory");
          else
            asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];" ::"l"(&tmo),
                         "r"(src), "r"(m0), "r"(n0), "r"(g)
                         : "memory");
          asm volatile("cp.async.bulk.commit_group;" ::: "memory");
          // all but the newest kSB - 1 stores have finished reading shared memory: the oldest of them frees its buffer
          asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(kSB - 1) : "memory");
          if (sblk >= kSB - 1) mbar_arrive(&bar_stfree[(sblk - (kSB - 1)) % kSB]);
        }
      }
      asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
      asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    }
    __syncwarp();
  } else if (warp == kWW + 1) {
    // =============================================================== MMA issuer (one thread)
    if (lane == 0) {
      int it = 0, s = 0;
      uint32_t ph = 0;
      const uint32_t b_hi = smem_u32(bhi), b_lo = smem_u32(blo);
      long long w_free = 0, w_split = 0;
      for (int ti = 0; ti < my_tiles; ++ti) {
        const int buf = ti & 1;
        long long t0 = p.prof ? clock64() : 0;
        if (ti >= 2) mbar_wait(&bar_accfree[buf], (uint32_t)(((ti >> 1) - 1) & 1));
        if (p.prof) w_free += clock64() - t0;
        const uint32_t acc = tmem + buf * NB;
        for (int c = 0; c < nchunk; ++c) {
          t0 = p.prof ? clock64() : 0;
          mbar_wait(&bar_split[s], ph);
          if (p.prof) w_split += clock64() - t0;
          tc_fence_after_sync();
          const uint32_t a_hi = p.loader == 2 ? smem_u32(hib + (it % kNLo) * kChunkBytes) : smem_u32(raw + s * kChunkBytes);
          const uint32_t a_lo = smem_u32(lob + (it % kNLo) * kChunkBytes);
#pragma unroll
          for (int g = 0; g < NKG; ++g) {
            const uint32_t boff = (uint32_t)(c * NKG + g) * (NB / 8) * 256;
            const uint64_t dah = make_smem_desc(a_hi + g * 1024, KC * 128, 512, kLayoutSw128Base32);
            const uint64_t dal = make_smem_desc(a_lo + g * 1024, KC * 128, 512, kLayoutSw128Base32);
            const uint64_t dbh = make_smem_desc(b_hi + boff, kPlainLbo, kPlainSbo, kLayoutNone);
            if (kFuseN) {
              mma_tf32(acc, dah, dbh, kIdesc2, !(c == 0 && g == 0));  // [hi*hi | hi*lo]
              mma_tf32(acc, dal, dbh, kIdesc, true);                   // lo*hi into the first half
            } else {
              const uint64_t dbl = make_smem_desc(b_lo + boff, kPlainLbo, kPlainSbo, kLayoutNone);
              mma_tf32(acc, dal, dbh, kIdesc, !(c == 0 && g == 0));
              mma_tf32(acc, dah, dbl, kIdesc, true);
              mma_tf32(acc, dah, dbh, kIdesc, true);
            }
          }
          mma_commit(&bar_done[s]);
          if (c == nchunk - 1) mma_commit(&bar_accfull[buf]);
          ++it;
          if (++s == NST) {
            s = 0;
            ph ^= 1;
          }
        }
      }
      if (p.prof) {
        p.prof[blockIdx.x * 8 + 2] = w_free;
        p.prof[blockIdx.x * 8 + 3] = w_split;
      }
    }
    __syncwarp();
  } else {
    // =============================================================== workers: operand split + epilogue
    // The epilogue of tile t runs AFTER the operands of tile t+1 have been split, so the tensor-core round trip of
    // tile t (issue, execute, commit, wake-up: ~1.5 us) is hidden behind useful work instead of being waited for.
    long long w_lo = 0, w_full = 0, w_acc = 0, t_epi = 0, t_split = 0;
    int st_blk = 0;  // running index of the 32-row output blocks this CTA has staged (bulk-store epilogue)
    const long long t_begin_w = p.prof ? clock64() : 0;
    auto epilogue = [&](int ti) {
      const long long te0 = p.prof ? clock64() : 0;
      const uint32_t tile = blockIdx.x + (uint32_t)ti * gridDim.x;
      const int g = tile / (uint32_t)p.tiles_per_slab;
      const int quarter = warp & 3, half = kHalves > 1 ? (warp >> 2) : 0;  // TMEM lane quarter / which 32-column blocks this warp takes
      const int m = (tile - (uint32_t)g * p.tiles_per_slab) * 128 + quarter * 32 + lane;
      const bool live = m < p.valid_m;
      const bool in_range = p.out_rw > 0 ? (m < p.mext && live) : m < p.mext;  // re-mapped: dead columns have no address
      long moff = m;
      if (p.out_rw > 0) {
        const int hh = m / p.out_rw;
        moff = in_range ? (long)hh * p.out_rp + (m - hh * p.out_rw) : 0;
      }
      float* po = p.out + (long)g * p.gso + moff;
      const uint32_t acc = tmem + (ti & 1) * NB + ((uint32_t)(quarter * 32) << 16);
      if (kTmaOut && p.tma_out) {
        // rows leave in blocks of 32 through the staging buffers; the store warp turns each block into one bulk tensor
        // store (or reduce-add) that clips rows >= nout and columns >= mext itself
        mbar_wait(&bar_accfull[ti & 1], (uint32_t)((ti >> 1) & 1));
        if (p.prof && p.prof_mode != 2) w_acc += clock64() - te0;
        tc_fence_after_sync();
        const bool tile_live = (int)((tile - (uint32_t)g * p.tiles_per_slab) * 128 + 128) <= p.valid_m;  // uniform
        const bool has_bias = p.bias != nullptr;
        constexpr int RH = 32 / kHalves;  // rows of a 32-row block per warp (two warps per TMEM lane quarter split it)
        for (int n0 = 0; n0 < p.nout; n0 += 32, ++st_blk) {
          float v[RH];
          long long q0 = p.prof_mode == 2 ? clock64() : 0, q1;
          if constexpr (RH == 32) tmem_ld32(acc + n0, *reinterpret_cast<float(*)[32]>(v));
          else tmem_ld16(acc + n0 + half * RH, *reinterpret_cast<float(*)[16]>(v));
          if (n0 + 32 >= p.nout) {  // last read of the accumulator buffer
            tc_fence_before_sync();
            mbar_arrive(&bar_accfree[ti & 1]);
          }
          if (p.prof_mode == 2) { q1 = clock64(); w_lo += q1 - q0; q0 = q1; }
          const int sb = st_blk % kSB;
          if (st_blk >= kSB) mbar_wait(&bar_stfree[sb], (uint32_t)(((st_blk / kSB) - 1) & 1));
          if (p.prof_mode == 2) { q1 = clock64(); w_full += q1 - q0; q0 = q1; }
          float* so = stage + sb * (32 * 128) + half * RH * 128 + (quarter * 32 + lane);
          const int nb = n0 + half * RH;  // first output row of this warp's share
          if (tile_live && !has_bias) {  // the common block: no predicates at all (rows >= nout are zeros of the
                                         // padded B image and are clipped by the bulk store)
            // the activation test stays OUTSIDE the unrolled loop: with a (uniform) branch per pair the 16 SELU chains of a
            // block ran one after the other (cycle counters: 51 cycles per pair, 6,600 of the tile's 11,400 cycles)
            if (p.act == 1) {
#pragma unroll
              for (int j = 0; j < RH; j += 2) {
                const float2 r = selu2(make_float2(v[j], v[j + 1]));
                so[j * 128] = r.x;
                so[(j + 1) * 128] = r.y;
              }
            } else {
#pragma unroll
              for (int j = 0; j < RH; ++j) so[j * 128] = v[j];
            }
          } else {
#pragma unroll
            for (int j = 0; j < RH; j += 2) {
              if (nb + j >= p.nout) break;  // warp uniform (rows beyond nout are clipped by the store anyway)
              float2 r = make_float2(v[j], v[j + 1]);
              if (has_bias) {
                r.x += sbias[nb + j];
                r.y += sbias[nb + j + 1];
              }
              if (p.act == 1) r = selu2(r);
              if (!live) r = make_float2(0.f, 0.f);
              so[j * 128] = r.x;
              so[(j + 1) * 128] = r.y;
            }
          }
          if (p.prof_mode == 2) { q1 = clock64(); w_acc += q1 - q0; q0 = q1; }
          fence_proxy_async_smem();
          mbar_arrive(&bar_stfull[sb]);
          if (p.prof_mode == 2) { q1 = clock64(); t_epi += q1 - q0; }
        }
        if (p.prof && p.prof_mode != 2) t_epi += clock64() - te0;
        return;
      }
      if (p.epi == 1) {
        // accumulate: out += acc.  The old values of a 32-row block are requested BEFORE the accumulator is waited for
        // and one block ahead of the stores (64 loads in flight per thread instead of 8: the epilogue was a chain of
        // exposed HBM round trips, 185 us against 77 us for the store-only form)
        float old[kHalves == 1 ? 2 : 1][32];
        auto fetch = [&](int n0, float (&o)[32]) {
          if (live && n0 < p.nout) {
            const float* q = po + (long)n0 * p.ldo;
#pragma unroll
            for (int j = 0; j < 32; ++j)
              if (n0 + j < p.nout) o[j] = __ldcs(q + (long)j * p.ldo);
          }
        };
        constexpr int kDepth = kHalves == 1 ? 2 : 1;  // 8 worker warps: registers allow one block in flight per warp
        fetch(32 * half, old[0]);
        mbar_wait(&bar_accfull[ti & 1], (uint32_t)((ti >> 1) & 1));
        if (p.prof) w_acc += clock64() - te0;
        tc_fence_after_sync();
        bool released = false;
#pragma unroll
        for (int b = 0; b < NPAD / 32 / kHalves; ++b) {
          const int n0 = 32 * (b * kHalves + half);
          if (n0 < p.nout) {
            if (kDepth == 2 && b + 1 < NPAD / 32 / kHalves) fetch(n0 + 32 * kHalves, old[(b + 1) % kDepth]);
            float v[32];
            tmem_ld32(acc + n0, v);
            if (n0 + 32 * kHalves >= p.nout || b + 1 == NPAD / 32 / kHalves) {  // this warp's last read of the buffer
              tc_fence_before_sync();
              mbar_arrive(&bar_accfree[ti & 1]);
              released = true;
            }
            if (live) {
              float* q = po + (long)n0 * p.ldo;
#pragma unroll
              for (int j = 0; j < 32; ++j)
                if (n0 + j < p.nout) {
                  const float sum = old[b % kDepth][j] + v[j];
                  __stcs(q + (long)j * p.ldo, p.act == 1 ? selu_f(sum) : sum);
                }
            }
            if (kDepth == 1 && b + 1 < NPAD / 32 / kHalves) fetch(n0 + 32 * kHalves, old[0]);
          }
        }
        if (!released) {  // this warp owns no block below nout
          tc_fence_before_sync();
          mbar_arrive(&bar_accfree[ti & 1]);
        }
        if (p.prof) t_epi += clock64() - te0;
        return;
      }
      mbar_wait(&bar_accfull[ti & 1], (uint32_t)((ti >> 1) & 1));
      if (p.prof) w_acc += clock64() - te0;
      tc_fence_after_sync();
      bool released = false;
#pragma unroll 1
      for (int n0 = 32 * half; n0 < NPAD; n0 += 32 * kHalves) {
        if (n0 >= p.nout) break;
        float v[32];
        tmem_ld32(acc + n0, v);
        if (kFuseN) {
          float v2[32];
          tmem_ld32(acc + NPAD + n0, v2);
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] += v2[j];
        }
        if (n0 + 32 * kHalves >= p.nout || n0 + 32 * kHalves >= NPAD) {  // this warp's last read of the buffer
          tc_fence_before_sync();
          mbar_arrive(&bar_accfree[ti & 1]);
          released = true;
        }
        if (in_range) {
          float* q = po + (long)n0 * p.ldo;
          if (p.epi == 1) {
            if (live) {
#pragma unroll
              for (int j0 = 0; j0 < 32; j0 += 8) {
                if (n0 + j0 >= p.nout) break;  // warp uniform
                const bool full = n0 + j0 + 8 <= p.nout;
                float old[8];
#pragma unroll
                for (int j = 0; j < 8; ++j)
                  if (full || n0 + j0 + j < p.nout) old[j] = q[(long)(j0 + j) * p.ldo];
#pragma unroll
                for (int j = 0; j < 8; ++j)
                  if (full || n0 + j0 + j < p.nout) q[(long)(j0 + j) * p.ldo] = old[j] + v[j0 + j];
              }
            }
          } else {
            // groups of 8 output rows: full groups run without per-row predicates (nout is 24 / 21 / 121 / 4 ...)
#pragma unroll
            for (int j0 = 0; j0 < 32; j0 += 8) {
              if (n0 + j0 >= p.nout) break;  // warp uniform
              float2 r[4];
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                const float2 b2 = *reinterpret_cast<const float2*>(sbias + n0 + j0 + 2 * j);
                r[j] = make_float2(v[j0 + 2 * j] + b2.x, v[j0 + 2 * j + 1] + b2.y);
                if (p.act == 1) r[j] = selu2(r[j]);
                if (!live) r[j] = make_float2(0.f, 0.f);
              }
              if (n0 + j0 + 8 <= p.nout) {
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                  q[(long)(j0 + 2 * j) * p.ldo] = r[j].x;
                  q[(long)(j0 + 2 * j + 1) * p.ldo] = r[j].y;
                }
              } else {
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                  if (n0 + j0 + 2 * j < p.nout) q[(long)(j0 + 2 * j) * p.ldo] = r[j].x;
                  if (n0 + j0 + 2 * j + 1 < p.nout) q[(long)(j0 + 2 * j + 1) * p.ldo] = r[j].y;
                }
              }
            }
          }
        }
      }
      if (!released) {  // this warp owns no 32-column block below nout
        tc_fence_before_sync();
        mbar_arrive(&bar_accfree[ti & 1]);
      }
      if (p.prof) t_epi += clock64() - te0;
    };
    int it = 0, s = 0;
    uint32_t ph = 0;
    const bool epi_warp = kSplitEpi && warp >= TcShape<NPAD>::kWorkerWarps;  // epilogue-only warpgroup
    // one extra pass of the tile loop runs the epilogue of the last tile: ONE inlined copy of the (large) epilogue
    for (int ti = 0; ti <= my_tiles; ++ti) {
      for (int c = 0; c < nchunk; ++c) {
        if (kSplitEpi) {
          if (epi_warp) {
            if (c == 0 && ti < my_tiles) epilogue(ti);
            continue;
          }
        }
        if (ti < my_tiles) {
        long long t0 = p.prof ? clock64() : 0;
        if (it >= kNLo) {  // the lo buffer is free once the MMAs of item it - kNLo have retired
          const int j = it - kNLo;
          mbar_wait(&bar_done[j % NST], (uint32_t)((j / NST) & 1));
        }
        if (p.prof && p.prof_mode != 2) {
          const long long t1 = clock64();
          w_lo += t1 - t0;
          t0 = t1;
        }
        mbar_wait(&bar_full[s], ph);
        if (p.prof && p.prof_mode != 2) {
          const long long t1 = clock64();
          w_full += t1 - t0;
          t0 = t1;
        }
        if (p.loader == 2) {
          // plain ring [k][128 m] -> hi and lo operand images in the 128B_BASE32B swizzle: 32-float block j of row r
          // lives at j * KC * 128 + r * 128, its 32-byte chunk c at chunk c ^ (r & 3).  This thread's float4 sits at
          // column c4 = tid & 31 of rows (tid >> 5) + 4 i, so (r & 3) and with it the whole intra-row offset are fixed.
          const float4* r4 = reinterpret_cast<const float4*>(raw + s * kChunkBytes);
          const int c4 = tid & 31, r0 = tid >> 5;
          const int q = c4 & 7;
          const uint32_t off0 = (uint32_t)(c4 >> 3) * (KC * 128) + r0 * 128 + ((((q >> 1) ^ r0) & 3) << 5) + ((q & 1) << 4);
          uint8_t* hb = hib + (it % kNLo) * kChunkBytes + off0;
          uint8_t* lb = lob + (it % kNLo) * kChunkBytes + off0;
#pragma unroll
          for (int i = 0; i < kChunkBytes / 16 / kTcWorkers; ++i) {
            const float4 x = r4[tid + i * kTcWorkers];
            *reinterpret_cast<float4*>(hb + i * (kTcWorkers * 4)) = x;  // kTcWorkers / 32 rows of 128 bytes per pass
            *reinterpret_cast<float4*>(lb + i * (kTcWorkers * 4)) = make_float4(tf32_lo(x.x), tf32_lo(x.y), tf32_lo(x.z), tf32_lo(x.w));
          }
          mbar_arrive(&bar_read[s]);  // the ring stage may be refilled
        } else {
          // hi operand = the fp32 word as it is (the tensor core ignores the 13 low mantissa bits);
          // lo operand = the exact remainder x - trunc_tf32(x)
          const float4* r4 = reinterpret_cast<const float4*>(raw + s * kChunkBytes);
          float4* l4 = reinterpret_cast<float4*>(lob + (it % kNLo) * kChunkBytes);
#pragma unroll
          for (int i = 0; i < kChunkBytes / 16 / kTcWorkers; ++i) {
            const int idx = tid + i * kTcWorkers;
            const float4 x = r4[idx];
            l4[idx] = make_float4(tf32_lo(x.x), tf32_lo(x.y), tf32_lo(x.z), tf32_lo(x.w));
          }
        }
        fence_proxy_async_smem();
        mbar_arrive(&bar_split[s]);
        if (p.prof) t_split += clock64() - t0;
        ++it;
        if (++s == NST) {
          s = 0;
          ph ^= 1;
        }
        }
        if (!kSplitEpi && c == nchunk - 1 && ti > 0) epilogue(ti - 1);
      }
    }
    if (p.prof && tid == 0) {
      p.prof[blockIdx.x * 8 + 4] = w_lo;
      p.prof[blockIdx.x * 8 + 5] = w_full;
      p.prof[blockIdx.x * 8 + 6] = w_acc;
      p.prof[blockIdx.x * 8 + 7] = t_epi;
      if (p.loader == 0) {
        p.prof[blockIdx.x * 8 + 0] = t_split;
        p.prof[blockIdx.x * 8 + 1] = clock64() - t_begin_w;
      }
    }
  }
  tc_fence_before_sync();
  __syncthreads();
  if (warp == kWW) tmem_dealloc(tmem, kTmemCols);
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
  static const int promo_env = getenv("HNO_TC_L2PROMO") ? atoi(getenv("HNO_TC_L2PROMO")) : 3;
  const CUtensorMapL2promotion promo = promo_env == 0   ? CU_TENSOR_MAP_L2_PROMOTION_NONE
                                       : promo_env == 1 ? CU_TENSOR_MAP_L2_PROMOTION_L2_64B
                                       : promo_env == 2 ? CU_TENSOR_MAP_L2_PROMOTION_L2_128B
                                                        : CU_TENSOR_MAP_L2_PROMOTION_L2_256B;
  const CUresult rc = fn(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, (cuuint32_t)rank, const_cast<float*>(base), d, s, b, e,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, sw, promo, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
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
  if (a.mext % 4) return false;
  if (a.in_rw < 0 || a.out_rw < 0 || a.in_rw % 2 || a.in_rp % 2) return false;
  if (a.in_rw > 0 && a.nsrc != 1) return false;
  if (a.mext < 1 || a.mext >= (1L << 30) || a.G < 1 || (a.mext + 127) / 128 * a.G >= (1L << 30)) return false;
  const int kc = a.kc;
  if (kc != 24 && kc != 32 && kc != 16 && kc != 8) return false;
  if (a.nout > 256) return false;
  const int npad = a.nout <= 32 ? 32 : (a.nout <= 128 ? 128 : 256);
  const int nchunk = a.chunks_per_src * a.nsrc;
  if ((size_t)2 * npad * nchunk * kc * 4 > 96 * 1024) return false;
  return true;
}

template <int KC, int NPAD, int NST, int kNLo>
static int launch_t(const TcStreamArgs& a, cudaStream_t st) {
  const int loader_sel = a.loader;  // resolved by tc_stream_launch
  CUtensorMap tm[2];
  for (int i = 0; i < 2; ++i) {
    const int j = i < a.nsrc ? i : 0;
    const uint64_t dims[3] = {(uint64_t)a.mext, (uint64_t)a.rows[j], (uint64_t)a.G};
    const uint64_t strides[2] = {(uint64_t)a.lda[j] * 4, (uint64_t)a.gsa[j] * 4};
    const uint32_t box[3] = {loader_sel == 2 ? 128u : 32u, (uint32_t)KC, 1};
    if (int rc = encode_tensor_map(&tm[i], a.a[j], 3, dims, strides, box, loader_sel == 2 ? 0 : 2)) return rc;
  }
  TcDev p;
  p.tma_out = 0;
  CUtensorMap tmo = tm[0];
  if (TcShape<NPAD>::kTmaOut) {
    static const bool tma_out_on = !(getenv("HNO_TC_TMA_OUT") && atoi(getenv("HNO_TC_TMA_OUT")) == 0);
    if (tma_out_on && !(a.epi == 1 && a.act == 1) /* a bulk reduce-add cannot apply the SELU */ && a.out_rw == 0 && reinterpret_cast<uintptr_t>(a.out) % 16 == 0 && a.ldo % 4 == 0 && a.gso % 4 == 0 &&
        a.nout >= 1) {
      const uint64_t dims[3] = {(uint64_t)a.mext, (uint64_t)a.nout, (uint64_t)a.G};
      const uint64_t strides[2] = {(uint64_t)a.ldo * 4, (uint64_t)a.gso * 4};
      const uint32_t box[3] = {128, 32, 1};
      if (int rc = encode_tensor_map(&tmo, a.out, 3, dims, strides, box, 0)) return rc;
      p.tma_out = 1;
    }
  }
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
  p.mext = (int)a.mext;
  p.valid_m = (int)(a.valid_m < a.mext ? a.valid_m : a.mext);
  p.chunks_per_src = a.chunks_per_src;
  p.nchunk = a.chunks_per_src * a.nsrc;
  p.tiles_per_slab = ceil_div(a.mext, 128);
  p.total_tiles = p.tiles_per_slab * a.G;
  p.act = a.act;
  p.epi = a.epi;
  for (int i = 0; i < 2; ++i) {
    const int j = i < a.nsrc ? i : 0;
    p.a[i] = a.a[j];
    p.lda[i] = a.lda[j];
    p.gsa[i] = a.gsa[j];
    p.rows[i] = a.rows[j];
  }
  p.loader = loader_sel;
  p.out_rw = a.out_rw;
  p.out_rp = a.out_rp;
  p.in_rw = a.in_rw;
  p.in_rp = a.in_rp;
  {
    // L2 prefetch cursor of the TMA producers: off by default since the rings became deep enough (2 CTAs x 4-6 stages
    // per SM); measured pw48f 0.150 ms at 96 KB ahead vs 0.131 ms without (ncu: 25 % extra DRAM reads with it)
    static const int pf_kb = getenv("HNO_TC_PREFETCH_KB") ? atoi(getenv("HNO_TC_PREFETCH_KB")) : 0;
    p.prefetch_items = pf_kb * 1024 / (KC * 512);
  }
  const size_t smem = TcSmem<KC, NPAD, NST, kNLo>::bytes(p.nchunk) + (p.loader == 2 ? (size_t)kNLo * KC * 512 : 0);
  auto kern = k_tc_stream<KC, NPAD, NST, kNLo>;
  HNO_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  HNO_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
  // CTAs per SM: cudaOccupancyMaxActiveBlocksPerMultiprocessor reports 1 for every kernel that allocates TMEM
  // (measured: also with 0 B of dynamic shared memory), while the hardware does co-schedule them; count resources here:
  // 228 KB of shared memory per SM (1 KB reserved per CTA + 1 KB static), <= 112 registers x 192 threads, 512 TMEM columns.
  int per_sm = (int)(233472 / (smem + 2 * 1024));
  if (per_sm > 3) per_sm = 3;
  if (TcShape<NPAD>::kSplitEpi && per_sm > 2) per_sm = 2;  // 320 threads per CTA, ~100 registers each
  if (per_sm < 1) per_sm = 1;
  if (per_sm > 512 / (2 * NPAD)) per_sm = 512 / (2 * NPAD);
  if (per_sm > 4) per_sm = 4;
  long grid = (long)sm_count() * per_sm;
  if (grid > p.total_tiles) grid = p.total_tiles;
  static const bool prof_on = getenv("HNO_TC_PROF") != nullptr;
  static long long* prof_buf = nullptr;
  p.prof = nullptr;
  p.prof_mode = prof_on ? atoi(getenv("HNO_TC_PROF")) : 0;
  if (prof_on) {
    if (!prof_buf) cudaMalloc(&prof_buf, 4096 * 8 * sizeof(long long));
    cudaMemsetAsync(prof_buf, 0, 4096 * 8 * sizeof(long long), st);
    p.prof = prof_buf;
  }
  static const bool pdl = !(getenv("HNO_TC_PDL") && atoi(getenv("HNO_TC_PDL")) == 0);
  if (pdl) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)grid);
    cfg.blockDim = dim3(TcShape<NPAD>::kThreads);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    HNO_CUDA(cudaLaunchKernelEx(&cfg, kern, tm[0], tm[1], tmo, p));
  } else {
    kern<<<(int)grid, TcShape<NPAD>::kThreads, smem, st>>>(tm[0], tm[1], tmo, p);
  }
  HNO_LAUNCH_CHECK();
  if (prof_on) {  // debug only: synchronous read-back of the per-CTA wait-cycle counters
    static long long host[4096 * 8];
    cudaStreamSynchronize(st);
    cudaMemcpy(host, prof_buf, grid * 8 * sizeof(long long), cudaMemcpyDeviceToHost);
    double s[8] = {0};
    for (long i = 0; i < grid; ++i)
      for (int j = 0; j < 8; ++j) s[j] += (double)host[i * 8 + j];
    const double tiles = (double)p.total_tiles / grid;
    fprintf(stderr,
            "[tc_prof KC=%d NPAD=%d NST=%d grid=%ld tiles/cta=%.1f nchunk=%d] cycles per tile per CTA: total %.0f | producer wait "
            "done (loader 0: worker split) %.0f | mma wait accfree %.0f, wait split %.0f | worker wait lo %.0f, wait full %.0f, wait acc %.0f, epilogue "
            "%.0f\n",
            KC, NPAD, NST, grid, tiles, p.nchunk, s[1] / grid / tiles, s[0] / grid / tiles, s[2] / grid / tiles,
            s[3] / grid / tiles, s[4] / grid / tiles, s[5] / grid / tiles, s[6] / grid / tiles, s[7] / grid / tiles);
  }
  return 0;
}

int tc_stream_launch(const TcStreamArgs& a_in, cudaStream_t st) {
  HNO_CHECK(tc_stream_eligible(a_in), "tc_stream: configuration is not eligible for the tensor-core path");
  TcStreamArgs a = a_in;
  {
    static const int loader = getenv("HNO_TC_LOADER") ? atoi(getenv("HNO_TC_LOADER")) : -1;
    if (loader >= 0) a.loader = loader;
    if (a.in_rw > 0) a.loader = 0;  // only the cp.async loader can gather
  }
  if (tc_analysis_eligible(a)) return tc_analysis_launch(a, st);
  const int npad = a.nout <= 32 ? 32 : (a.nout <= 128 ? 128 : 256);
  // loader 2 keeps a third set of chunk buffers (the hi image): its own stage counts so that 2 CTAs still fit an SM
  if (a.loader == 2 && a.kc == 24 && npad == 32) return launch_t<24, 32, 4, 2>(a, st);
  if (a.loader == 2 && a.kc == 16 && npad == 32) return launch_t<16, 32, 3, 2>(a, st);
#define HNO_TC_CASE(KC_, NP_, NST_, NLO_)                                  \
  if (a.kc == KC_ && npad == NP_) return launch_t<KC_, NP_, NST_, NLO_>(a, st);
  HNO_TC_CASE(24, 32, 5, 3)   // pointwise conv 24(+24) -> <= 32: 2 CTAs / SM of 4 split + 4 epilogue warps
  HNO_TC_CASE(32, 32, 3, 2)
  HNO_TC_CASE(16, 32, 6, 3)   // D / H-axis analysis: 8 KB chunks (TMA loader + L2 prefetch cursor)
  HNO_TC_CASE(8, 32, 4, 2)
  HNO_TC_CASE(24, 128, 2, 1)  // D-axis synthesis (one chunk per tile); smem: 3 staging buffers instead of a second lo buffer
  HNO_TC_CASE(16, 128, 2, 1)  // H-axis synthesis (29 rows = two chunks)
  HNO_TC_CASE(32, 128, 2, 1)
  HNO_TC_CASE(8, 128, 3, 2)
  HNO_TC_CASE(24, 256, 3, 2)  // synthesis onto axes of 129..256 samples (2x super-resolution grids)
  HNO_TC_CASE(16, 256, 3, 2)
  HNO_TC_CASE(32, 256, 3, 2)
  HNO_TC_CASE(8, 256, 3, 2)
#undef HNO_TC_CASE
  set_error("tc_stream: no kernel instance for kc=%d npad=%d", a.kc, npad);
  return -1;
}

}  // namespace hno
