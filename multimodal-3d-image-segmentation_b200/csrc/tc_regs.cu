// Streamed tall-skinny tensor-core contraction, register-fed variant (tcgen05 with the A operand in TENSOR MEMORY).
//
//   out[g][n][m] = EPI( sum_k  A[g][k][m] * Bm[n][k] )        m = contiguous axis (voxels / plane columns)
//
// Same contract as tc_stream.cu (see tc_stream.h), different data path.  Measurements on B200 that led here
// (profiles/r1b_*):
//   * TMA tile loads of 128-byte-wide boxes -- the widest a swizzled MN-major tf32 shared-memory operand allows --
//     are limited by the TMA unit to one box row per ~8.6 cycles per SM = 4.2 TB/s chip-wide (tools/ubench_tma.cu);
//   * with an LDGSTS loader the shared-memory ring version is bounded by shared-memory bandwidth: the streamed
//     operand crosses it 6 times per tile (async write, split read, lo write, 3 MMA reads).
//   * feeding the operand global -> registers directly (2 x 24 registers per thread in flight, 72 KB per SM) is
//     latency bound at 3.4 TB/s: HBM needs > 100 KB in flight per SM (tools/ubench_strided.cu).
// Here the streamed operand crosses shared memory exactly twice and never as an MMA operand: a loader warp streams
// 512-byte rows into a deep cp.async ring (plain [k][128 m] layout, up to 8 x 8 KB per CTA, 3 CTAs per SM, completion
// on mbarriers), every worker thread owns one voxel (= one TMEM lane), reads its KC channel values from the ring
// (conflict-free LDS.32), splits them into the two TF32 terms in registers and writes both into tensor memory with
// tcgen05.st; the MMA warp then issues tcgen05.mma with A from TMEM and the small resident B image (weights /
// cas-basis rows, hi and lo) from shared memory.  A ring stage is released as soon as it has been read into registers.
//
// Precision: 3xTF32 exactly as in tc_stream.cu (hi = the fp32 word, whose 13 low mantissa bits the tensor core
// ignores; lo = x - trunc_tf32(x); products lo*hi + hi*lo + hi*hi accumulated in fp32 in TMEM).
#include "common.cuh"
#include "tc_common.cuh"
#include "tc_stream.h"

#include <stdio.h>
#include <stdlib.h>

namespace hno {

using namespace tc;

constexpr int kRegKC = 24;         // rows (k) per chunk = values per thread and ring stage (three MMA k-steps)
constexpr int kRegWorkers = 128;   // 4 worker warps, thread t <-> TMEM lane t <-> voxel m0 + t
constexpr int kRegThreads = kRegWorkers + 64;  // + warp 4: MMA issuer, warp 5: cp.async loader
constexpr int kRegSlots = 1;                   // A slots in tensor memory (TMEM columns are the scarce resource)
constexpr int kRegACols = kRegSlots * 2 * kRegKC;  // slots x (hi, lo)

struct TcRegDev {
  const float* a[2];
  long lda[2], gsa[2];
  int rows[2];
  int cps;             // chunks per source
  int nchunk;          // cps * nsrc
  const float* b;
  long ldbn, ldbk;
  int nvalid;          // valid rows of B (= nout)
  float scale;
  const float* bias;
  float* out;
  long ldo, gso;
  int nout;
  int mext, valid_m;
  int tiles_per_slab, total_tiles;
  int act, epi;
  int nst;             // ring stages
  long long* prof;     // debug (HNO_TC_PROF=1): per-CTA cycle counters [grid][8]
  int dbg;             // debug knobs (HNO_TC_DBG): 1 no activation, 2 no stores, 4 no accumulator read, 8 no MMA
};

__device__ __forceinline__ void mbar_arrive_cta(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t* r) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"r"(taddr), "r"(r[0]), "r"(r[1]),
               "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
               : "memory");
}
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t* r) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};" ::"r"(taddr),
      "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
      "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
      : "memory");
}
// MMA with A in tensor memory
__device__ __forceinline__ void mma_tf32_ta(uint32_t tmem_d, uint32_t tmem_a, uint64_t desc_b, uint32_t idesc,
                                            bool accumulate) {
  const uint32_t acc = accumulate ? 1u : 0u;
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}" ::"r"(tmem_d),
      "r"(tmem_a), "l"(desc_b), "r"(idesc), "r"(acc)
      : "memory");
}

template <int NPAD>
__global__ void __launch_bounds__(kRegThreads, (NPAD <= 32 ? 3 : 2)) k_tc_regs(const TcRegDev p) {
  constexpr int KC = kRegKC;
  // Every tcgen05.mma costs its issuing thread >= ~102 cycles whatever N <= 128 is (tools/ubench_mma.cu), so the number of
  // MMA instructions per tile is what matters, not their size.  kFuse: A_hi * [B_hi | B_lo] is ONE instruction of
  // N = 2 NPAD (the two halves are added in the epilogue) and A_lo * B_hi goes into the first half: 2 instead of 3
  // instructions per k-step.
  constexpr bool kFuse = NPAD <= 32;
  constexpr int NB = kFuse ? 2 * NPAD : NPAD;  // rows of the B image / accumulator columns
  constexpr uint32_t kIdesc = make_idesc_tf32(128, NPAD, 0, 0);
  constexpr uint32_t kIdescB = make_idesc_tf32(128, NB, 0, 0);
  constexpr uint32_t kDCol = kRegACols;                              // accumulator columns start after the A slots
  constexpr uint32_t kTmemCols = (kRegACols + NB) <= 128 ? 128 : ((kRegACols + NB) <= 256 ? 256 : 512);
  static_assert(NPAD % 32 == 0 && NPAD <= 256, "bad tile configuration");
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  const int ktot = p.nchunk * KC;
  constexpr int kStageBytes = KC * 512;
  constexpr int kMaxStages = 8;
  const int NST = p.nst;
  uint8_t* ring = smem;                             // [NST][KC][128] fp32
  float* bhi = reinterpret_cast<float*>(ring + NST * kStageBytes);
  float* blo = bhi + NPAD * ktot;
  float* sbias = blo + NPAD * ktot;
  __shared__ __align__(8) uint64_t bar_full[kMaxStages];   // stage landed            (32 cp.async arrivals)
  __shared__ __align__(8) uint64_t bar_empty[kMaxStages];  // stage read by the workers (128 arrivals)
  __shared__ __align__(8) uint64_t bar_ready[2];  // A slot written by the 128 workers
  __shared__ __align__(8) uint64_t bar_free[2];   // MMAs that read the slot have retired   (tcgen05.commit)
  __shared__ __align__(8) uint64_t bar_accfull;   // all MMAs of a tile have retired          (tcgen05.commit)
  __shared__ uint32_t tmem_slot;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  // ---- resident operand: B image (hi / lo) in the K-major core-matrix layout, bias
  for (int idx = tid; idx < NPAD * ktot; idx += kRegThreads) {
    const int n = idx / ktot, k = idx - n * ktot;
    const int c = k / KC, r = k - c * KC;
    const int src = c / p.cps;
    const int row = (c - src * p.cps) * KC + r;
    float v = 0.f;
    if (n < p.nvalid && row < p.rows[src])
      v = p.scale * __ldg(p.b + (long)n * p.ldbn + (long)(src * p.rows[0] + row) * p.ldbk);
    const float hi = __uint_as_float((__float_as_uint(v) + 0x1000u) & 0xFFFFE000u);
    if (kFuse) {  // one image of 2 NPAD rows: [0, NPAD) = hi, [NPAD, 2 NPAD) = lo
      bhi[kmajor_plain_index<NB>(n, k)] = hi;
      bhi[kmajor_plain_index<NB>(NPAD + n, k)] = v - hi;
    } else {
      const int o = kmajor_plain_index<NPAD>(n, k);
      bhi[o] = hi;
      blo[o] = v - hi;
    }
  }
  for (int n = tid; n < NPAD; n += kRegThreads) sbias[n] = (p.bias != nullptr && n < p.nout) ? __ldg(p.bias + n) : 0.f;
  if (tid == 0) {
    mbar_init(&bar_ready[0], kRegWorkers);
    mbar_init(&bar_ready[1], kRegWorkers);
    mbar_init(&bar_free[0], 1);
    mbar_init(&bar_free[1], 1);
    mbar_init(&bar_accfull, 1);
    for (int s = 0; s < kMaxStages; ++s) {
      mbar_init(&bar_full[s], 32);
      mbar_init(&bar_empty[s], kRegWorkers);
    }
    mbar_fence_init();
  }
  if (warp == 4) tmem_alloc(&tmem_slot, kTmemCols);
  fence_proxy_async_smem();
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem = tmem_slot;

  const int my_tiles = p.total_tiles > (int)blockIdx.x ? (p.total_tiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
  const int nchunk = p.nchunk;
  const int items = my_tiles * nchunk;

  if (warp == 4) {
    // =============================================================== MMA issuer (one thread)
    if (lane == 0) {
      const uint32_t b_hi = smem_u32(bhi), b_lo = smem_u32(blo);
      const uint32_t acc = tmem + kDCol;
      int c = 0;
      for (int it = 0; it < items; ++it) {
        const int slot = it % kRegSlots;
        mbar_wait(&bar_ready[slot], (uint32_t)((it / kRegSlots) & 1));
        tc_fence_after_sync();
        const uint32_t a_hi = tmem + slot * (2 * KC), a_lo = a_hi + KC;
#pragma unroll
        for (int g = 0; g < KC / 8; ++g) {
          const uint32_t boff = (uint32_t)(c * (KC / 8) + g) * (NB / 8) * 256;
          const uint64_t dbh = make_smem_desc(b_hi + boff, kPlainLbo, kPlainSbo, kLayoutNone);
          if (kFuse) {
            mma_tf32_ta(acc, a_hi + 8 * g, dbh, kIdescB, !(c == 0 && g == 0));  // [hi*hi | hi*lo]
            mma_tf32_ta(acc, a_lo + 8 * g, dbh, kIdesc, true);                   // lo*hi into the first half
          } else {
            const uint64_t dbl = make_smem_desc(b_lo + boff, kPlainLbo, kPlainSbo, kLayoutNone);
            mma_tf32_ta(acc, a_lo + 8 * g, dbh, kIdesc, !(c == 0 && g == 0));
            mma_tf32_ta(acc, a_hi + 8 * g, dbl, kIdesc, true);
            mma_tf32_ta(acc, a_hi + 8 * g, dbh, kIdesc, true);
          }
        }
        mma_commit(&bar_free[slot]);
        if (++c == nchunk) {
          c = 0;
          mma_commit(&bar_accfull);
        }
      }
    }
    __syncwarp();
  } else if (warp == 5) {
    // =============================================================== cp.async loader (one warp): 512-byte rows
    int s = 0;
    uint32_t ph = 0;
    int it = 0;
    for (int ti = 0; ti < my_tiles; ++ti) {
      const uint32_t tile = blockIdx.x + (uint32_t)ti * gridDim.x;
      const int g = tile / (uint32_t)p.tiles_per_slab;
      const int col = (tile - (uint32_t)g * p.tiles_per_slab) * 128 + lane * 4;
      const bool col_ok = col < p.mext;
      for (int c = 0; c < nchunk; ++c) {
        const int src = c / p.cps;
        const int row0 = (c - src * p.cps) * KC;
        if (it >= NST) mbar_wait(&bar_empty[s], ph ^ 1);  // previous contents are in the workers' registers
        const long ld = p.lda[src];
        const float* gp = p.a[src] + (long)g * p.gsa[src] + (long)row0 * ld + (col_ok ? col : 0);
        const int nrow = p.rows[src] - row0;  // rows of this chunk that exist (the rest reads as zero)
        const uint32_t dst0 = smem_u32(ring + s * kStageBytes) + lane * 16;
#pragma unroll 8
        for (int r = 0; r < KC; ++r) {
          const bool ok = col_ok && r < nrow;
          const float* sp = ok ? gp + (long)r * ld : p.a[src];
          asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst0 + r * 512), "l"(sp), "r"(ok ? 16 : 0)
                       : "memory");
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
  } else {
    // =============================================================== workers: ring -> registers -> TMEM, epilogue
    const uint32_t lane_base = tmem + ((uint32_t)(warp * 32) << 16);

    long long pt_acc = 0;
    auto epilogue = [&](int ti) {
      const long long ta = p.prof ? clock64() : 0;
      mbar_wait(&bar_accfull, (uint32_t)(ti & 1));
      if (p.prof) pt_acc += clock64() - ta;
      tc_fence_after_sync();
      const uint32_t tile = blockIdx.x + (uint32_t)ti * gridDim.x;
      const int g = tile / (uint32_t)p.tiles_per_slab;
      const int m = (tile - (uint32_t)g * p.tiles_per_slab) * 128 + tid;
      const bool in_range = m < p.mext;
      const bool live = m < p.valid_m;
      float* po = p.out + (long)g * p.gso + m;
      const uint32_t acc = lane_base + kDCol;
#pragma unroll 1
      for (int n0 = 0; n0 < NPAD; n0 += 32) {
        if (n0 >= p.nout) break;
        float v[32];
        if (!(p.dbg & 4)) tmem_ld32(acc + n0, v);
        if (kFuse) {
          float v2[32];
          tmem_ld32(acc + NPAD + n0, v2);
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] += v2[j];
        }
        if (in_range && !(p.dbg & 2)) {
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
#pragma unroll
            for (int j0 = 0; j0 < 32; j0 += 8) {
              if (n0 + j0 >= p.nout) break;  // warp uniform
              float2 r[4];
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                const float2 b2 = *reinterpret_cast<const float2*>(sbias + n0 + j0 + 2 * j);
                r[j] = make_float2(v[j0 + 2 * j] + b2.x, v[j0 + 2 * j + 1] + b2.y);
                if (p.act == 1 && !(p.dbg & 1)) r[j] = selu2(r[j]);
                if (!live) r[j] = make_float2(0.f, 0.f);
              }
              if (n0 + j0 + 8 <= p.nout) {
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                  __stcs(q + (long)(j0 + 2 * j) * p.ldo, r[j].x);
                  __stcs(q + (long)(j0 + 2 * j + 1) * p.ldo, r[j].y);
                }
              } else {
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                  if (n0 + j0 + 2 * j < p.nout) __stcs(q + (long)(j0 + 2 * j) * p.ldo, r[j].x);
                  if (n0 + j0 + 2 * j + 1 < p.nout) __stcs(q + (long)(j0 + 2 * j + 1) * p.ldo, r[j].y);
                }
              }
            }
          }
        }
      }
    };

    int cs_ti = 0, cs_c = 0;
    int s = 0;
    uint32_t ph = 0;
    long long pt[6] = {0, 0, 0, 0, 0, 0};
    const bool prof = p.prof != nullptr;
    const long long t_begin = prof ? clock64() : 0;
    for (int it = 0; it < items; ++it) {
      const int slot = it % kRegSlots;
      long long t0 = prof ? clock64() : 0, t1;
      mbar_wait(&bar_full[s], ph);
      if (prof) { t1 = clock64(); pt[0] += t1 - t0; t0 = t1; }
      const float* src = reinterpret_cast<const float*>(ring + s * kStageBytes) + tid;
      uint32_t hi[KC], lo[KC];
#pragma unroll
      for (int k = 0; k < KC; ++k) {
        const float x = src[k * 128];
        hi[k] = __float_as_uint(x);
        lo[k] = __float_as_uint(tf32_lo(x));
      }
      if (prof) { t1 = clock64(); pt[1] += t1 - t0; t0 = t1; }
      if (it >= kRegSlots)  // the MMAs of the previous user of the slot have read it
        mbar_wait(&bar_free[slot], (uint32_t)(((it / kRegSlots) - 1) & 1));
      tc_fence_after_sync();
      if (prof) { t1 = clock64(); pt[2] += t1 - t0; t0 = t1; }
      const uint32_t a0 = lane_base + slot * (2 * KC);
      static_assert(KC == 24, "the TMEM stores below are written for 24 values per chunk");
      tmem_st16(a0, hi);
      tmem_st8(a0 + 16, hi + 16);
      tmem_st16(a0 + KC, lo);
      tmem_st8(a0 + KC + 16, lo + 16);
      mbar_arrive_cta(&bar_empty[s]);  // the stage is in registers (the stores above consumed every value)
      asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
      tc_fence_before_sync();
      mbar_arrive_cta(&bar_ready[slot]);
      if (prof) { t1 = clock64(); pt[3] += t1 - t0; t0 = t1; }
      if (++s == NST) {
        s = 0;
        ph ^= 1;
      }
      if (++cs_c == nchunk) {
        cs_c = 0;
        epilogue(cs_ti++);
        if (prof) pt[4] += clock64() - t0;
      }
    }
    if (prof && tid == 0) {
      for (int i = 0; i < 5; ++i) p.prof[blockIdx.x * 8 + i] = pt[i];
      p.prof[blockIdx.x * 8 + 5] = pt_acc;
      p.prof[blockIdx.x * 8 + 6] = clock64() - t_begin;
    }
  }
  tc_fence_before_sync();
  __syncthreads();
  if (warp == 4) tmem_dealloc(tmem, kTmemCols);
}

// ------------------------------------------------------------------------------------------------ host side
bool tc_regs_eligible(const TcStreamArgs& a) {
  if (!tc_enabled()) return false;
  if (a.nsrc < 1 || a.nsrc > 2) return false;
  if (a.nsrc == 2 && a.rows[0] != a.rows[1]) return false;
  if (a.mext < 1 || a.mext >= (1L << 30) || a.G < 1 || (a.mext + 127) / 128 * a.G >= (1L << 30)) return false;
  if (a.mext % 4) return false;
  for (int i = 0; i < a.nsrc; ++i) {  // 16-byte cp.async pieces
    if (reinterpret_cast<uintptr_t>(a.a[i]) % 16) return false;
    if (a.lda[i] % 4 || a.gsa[i] % 4) return false;
  }
  if (a.nout < 1 || a.nout > 128) return false;
  const int npad = a.nout <= 32 ? 32 : 128;
  const int cps = (a.rows[0] + kRegKC - 1) / kRegKC;
  if ((size_t)2 * npad * cps * a.nsrc * kRegKC * 4 > 64 * 1024) return false;
  return true;
}

template <int NPAD>
static int launch_regs(const TcStreamArgs& a, cudaStream_t st) {
  TcRegDev p;
  for (int i = 0; i < 2; ++i) {
    const int j = i < a.nsrc ? i : 0;
    p.a[i] = a.a[j];
    p.lda[i] = a.lda[j];
    p.gsa[i] = a.gsa[j];
    p.rows[i] = a.rows[j];
  }
  p.cps = (a.rows[0] + kRegKC - 1) / kRegKC;
  p.nchunk = p.cps * a.nsrc;
  p.b = a.b;
  p.ldbn = a.ldbn;
  p.ldbk = a.ldbk;
  p.nvalid = a.nout;
  p.scale = a.scale;
  p.bias = a.bias;
  p.out = a.out;
  p.ldo = a.ldo;
  p.gso = a.gso;
  p.nout = a.nout;
  p.mext = (int)a.mext;
  p.valid_m = (int)(a.valid_m < a.mext ? a.valid_m : a.mext);
  p.tiles_per_slab = ceil_div(a.mext, 128);
  p.total_tiles = p.tiles_per_slab * a.G;
  p.act = a.act;
  p.epi = a.epi;
  const size_t fixed = 1024 + (size_t)2 * NPAD * p.nchunk * kRegKC * 4 + NPAD * 4;
  const int want_ctas = NPAD <= 32 ? 3 : 2;
  int nst = (int)((233472 / want_ctas - 2 * 1024 - fixed) / (kRegKC * 512));
  static const int nst_env = getenv("HNO_TC_NST") ? atoi(getenv("HNO_TC_NST")) : 0;
  if (nst_env > 0) nst = nst_env;
  if (nst > 8) nst = 8;
  if (nst < 2) nst = 2;
  p.nst = nst;
  static const int dbg_env = getenv("HNO_TC_DBG") ? atoi(getenv("HNO_TC_DBG")) : 0;
  p.dbg = dbg_env;
  const size_t smem = fixed + (size_t)nst * kRegKC * 512;
  auto kern = k_tc_regs<NPAD>;
  HNO_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  // CTAs per SM (the occupancy API reports 1 for kernels that allocate TMEM): registers 3 x 192 x 112 <= 64 K,
  // TMEM 3 x 128 columns (NPAD 32) / 2 x 256 columns (NPAD 128), shared memory
  int per_sm = NPAD <= 32 ? 3 : 2;
  const int by_smem = (int)(233472 / (smem + 2 * 1024));
  if (per_sm > by_smem) per_sm = by_smem < 1 ? 1 : by_smem;
  static const int per_sm_env = getenv("HNO_TC_CTAS") ? atoi(getenv("HNO_TC_CTAS")) : 0;
  if (per_sm_env > 0) per_sm = per_sm_env;
  long grid = (long)sm_count() * per_sm;
  if (grid > p.total_tiles) grid = p.total_tiles;
  static const bool prof_on = getenv("HNO_TC_PROF") != nullptr;
  static long long* prof_buf = nullptr;
  p.prof = nullptr;
  if (prof_on) {
    if (!prof_buf) cudaMalloc(&prof_buf, 4096 * 8 * sizeof(long long));
    cudaMemsetAsync(prof_buf, 0, 4096 * 8 * sizeof(long long), st);
    p.prof = prof_buf;
  }
  kern<<<(int)grid, kRegThreads, smem, st>>>(p);
  HNO_LAUNCH_CHECK();
  if (prof_on) {  // debug only: synchronous read-back of the per-CTA cycle counters of worker thread 0
    static long long host[4096 * 8];
    cudaStreamSynchronize(st);
    cudaMemcpy(host, prof_buf, grid * 8 * sizeof(long long), cudaMemcpyDeviceToHost);
    double s[8] = {0};
    for (long i = 0; i < grid; ++i)
      for (int j = 0; j < 8; ++j) s[j] += (double)host[i * 8 + j];
    const double d = (double)grid * p.total_tiles / grid;
    fprintf(stderr,
            "[tc_regs NPAD=%d nst=%d grid=%ld tiles/cta=%.1f nchunk=%d] worker cycles per tile: total %.0f | wait data %.0f | "
            "lds+split %.0f | wait slot %.0f | st+signal %.0f | epilogue %.0f (of which wait MMA %.0f)\n",
            NPAD, p.nst, grid, (double)p.total_tiles / grid, p.nchunk, s[6] / d, s[0] / d, s[1] / d, s[2] / d, s[3] / d,
            s[4] / d, s[5] / d);
  }
  return 0;
}

int tc_regs_launch(const TcStreamArgs& a, cudaStream_t st) {
  HNO_CHECK(tc_regs_eligible(a), "tc_regs: configuration is not eligible for the tensor-core path");
  if (a.nout <= 32) return launch_regs<32>(a, st);
  return launch_regs<128>(a, st);
}

}  // namespace hno
