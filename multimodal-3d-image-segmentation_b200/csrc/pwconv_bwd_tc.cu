// Backward of the 24(+24) -> 24 pointwise (1x1x1) convolution + SELU, pipelined (sm_100a).
//
// Replaces the autograd of nets/nets_utils.py:120-174 (ConvNormAct k=1, SELU) for the concat / mapping convolutions of
// nets/hnosegxs.py:254-255, 273-275 -- the kernel with the largest share of an HNOSeg-XS training step.
//   d(pre)[o]  = dy[o] * selu'(y[o])
//   din[i]     = sum_o W[o][i] d(pre)[o]        (x selu'(in1[i]) for the first input when it is itself a SELU output)
//   dW[o][i]  += sum_voxels d(pre)[o] in[i],   db[o] += sum_voxels d(pre)[o]
// The phase-by-phase FFMA kernel (pwconv_kernels.cu) was latency bound at 2.1 TB/s: five CTA-wide barriers per tile and
// no overlap of its loads with its math (ncu: FMA pipe 37 %, DRAM 26 %).  Here
//   * a loader warp streams dy, y, in1, in2 of a 128-voxel tile as four 24-row stages of a deep cp.async ring
//     (completion on mbarriers, rows padded to 132 floats so that the weight-gradient reads are conflict free);
//   * every worker thread owns one voxel (= one TMEM lane): it forms d(pre) in registers, writes the two TF32 terms to
//     tensor memory and the fp32 value to a shared tile;
//   * the input gradient runs on the tensor cores: A = d(pre) from TMEM (M = 128 voxels, K = 24), B = [W^T_hi | W^T_lo]
//     resident in shared memory, 3xTF32 as two instructions per k-step (one instruction A [B_hi | B_lo] of N = 2 NP whose halves are added in the epilogue, plus A_lo B_hi into the first half);
//   * while those MMAs execute, the workers accumulate the weight gradient with packed FFMA2 out of shared memory
//     (exact fp32 products, fp32 partial sums per CTA, fp64 final reduction: k_reduce_partials);
//   * the epilogue reads the accumulator (one lane per voxel), applies selu'(in1) / the += of U-Net skips and writes
//     coalesced 128-byte rows.
#include "common.cuh"
#include "tc_common.cuh"
#include "tc_stream.h"
#include "wgrad.cuh"

#include <stdlib.h>

namespace hno {

using namespace tc;

constexpr int kBtC = 24;                      // channels per source / output channels
constexpr int kBtWorkers = 128;
constexpr int kBtThreads = kBtWorkers + 64;   // + warp 4: MMA issuer, warp 5: loader
constexpr int kBtPitch = 132;                 // floats per ring / tile row (128 voxels + 4: bank spread for the LDS.128 reads)
constexpr int kBtStageFloats = kBtC * kBtPitch;
constexpr int kBtMaxStages = 8;

struct BtDev {
  const float* dy;
  const float* y;
  const float* in1;
  const float* in2;
  const float* w;      // [24][CI]
  float* din1;
  float* din2;
  float* partials;     // [grid][24 * CI + 24]
  long S, P, HW;
  int tiles_per_sample, total_tiles;
  int flags;
  int nst;
  int tma;             // 1: the ring is filled by one bulk tensor copy per stage (box 132 x 24), 0: by the cp.async warp
};

// the four streamed tensors (dy, y, in1, in2) as [B][24][S] tensor maps with a 132-float-wide box: the 4 extra columns
// reproduce the padded row pitch of the ring (bank spread of the weight-gradient reads) and are never used as data
struct BtMaps {
  CUtensorMap m[4];
};

__device__ __forceinline__ void bt_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void bt_st8(uint32_t taddr, const uint32_t* r) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"r"(taddr), "r"(r[0]), "r"(r[1]),
               "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
               : "memory");
}
__device__ __forceinline__ void bt_st16(uint32_t taddr, const uint32_t* r) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};" ::"r"(taddr),
      "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
      "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
      : "memory");
}
__device__ __forceinline__ void bt_ld8(uint32_t taddr, float (&v)[8]) {
  uint32_t r[8];
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 8; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void bt_mma(uint32_t tmem_d, uint32_t tmem_a, uint64_t desc_b, uint32_t idesc, bool accumulate) {
  const uint32_t acc = accumulate ? 1u : 0u;
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}" ::"r"(tmem_d),
      "r"(tmem_a), "l"(desc_b), "r"(idesc), "r"(acc)
      : "memory");
}
__device__ __forceinline__ void bt_worker_sync() { asm volatile("bar.sync 1, 128;" ::: "memory"); }

// CI2 = 0: single input (24 -> 24);  CI2 = 24: virtual concat of two inputs (48 -> 24)
template <int CI2, bool ACC>
__global__ void __launch_bounds__(kBtThreads, 2) k_pwconv_bwd_tc(const BtDev p, const __grid_constant__ BtMaps maps) {
  constexpr int C = kBtC;
  constexpr int CI = C + CI2;
  constexpr int NP = CI == 48 ? 48 : 32;       // MMA N of one half (multiple of 16)
  constexpr int NB = 2 * NP;                    // rows of the fused B image [hi | lo]
  constexpr int NSRC = CI2 > 0 ? 4 : 3;         // ring stages per tile: dy, y, in1 (, in2)
  constexpr uint32_t kIdesc1 = make_idesc_tf32(128, NB, 0, 0);
  constexpr uint32_t kIdesc2 = make_idesc_tf32(128, NP, 0, 0);
  constexpr uint32_t kACol = 0, kDCol = 2 * C;  // TMEM: A = d(pre) hi [0,24) lo [24,48); accumulator [48, 48 + NB)
  constexpr uint32_t kTmemCols = 256;
  constexpr int PSTRIDE = C * CI + C;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  const int NST = p.nst;
  float* ring = reinterpret_cast<float*>(smem);            // [NST][24][132]
  float* sdp = ring + NST * kBtStageFloats;                 // [24][132]   d(pre) of the current tile
  float* bimg = sdp + kBtStageFloats;                       // [NB rows][24 k] K-major core-matrix image of W^T (hi | lo)
  float* scratch = bimg + NB * C;                           // end-of-kernel reduction of the weight-gradient tiles
  __shared__ __align__(8) uint64_t bar_full[kBtMaxStages];  // stage landed              (32 cp.async arrivals)
  __shared__ __align__(8) uint64_t bar_empty[kBtMaxStages]; // stage no longer needed    (128 arrivals)
  __shared__ __align__(8) uint64_t bar_ready;               // d(pre) is in tensor memory (128 arrivals)
  __shared__ __align__(8) uint64_t bar_accfull;             // the MMAs of the tile have retired (tcgen05.commit)
  __shared__ uint32_t tmem_slot;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  // ---- resident operand: B[n][k] = W[k][n] (n = input channel, k = output channel), hi rows then lo rows
  for (int idx = tid; idx < NP * C; idx += kBtThreads) {
    const int n = idx / C, k = idx - n * C;
    const float v = n < CI ? __ldg(p.w + k * CI + n) : 0.f;
    const float hi = __uint_as_float((__float_as_uint(v) + 0x1000u) & 0xFFFFE000u);
    bimg[kmajor_plain_index<NB>(n, k)] = hi;
    bimg[kmajor_plain_index<NB>(NP + n, k)] = v - hi;
  }
  if (tid == 0) {
    for (int s = 0; s < kBtMaxStages; ++s) {
      mbar_init(&bar_full[s], p.tma ? 1 : 32);
      mbar_init(&bar_empty[s], kBtWorkers);
    }
    mbar_init(&bar_ready, kBtWorkers);
    mbar_init(&bar_accfull, 1);
    mbar_fence_init();
  }
  if (warp == 4) tmem_alloc(&tmem_slot, kTmemCols);
  fence_proxy_async_smem();
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem = tmem_slot;
  const int my_tiles = p.total_tiles > (int)blockIdx.x ? (p.total_tiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;

  if (warp == 4) {
    // =============================================================== MMA issuer (one thread): input gradient
    if (lane == 0) {
      const uint32_t b0 = smem_u32(bimg);
      for (int ti = 0; ti < my_tiles; ++ti) {
        mbar_wait(&bar_ready, (uint32_t)(ti & 1));
        tc_fence_after_sync();
#pragma unroll
        for (int g = 0; g < C / 8; ++g) {
          const uint64_t db = make_smem_desc(b0 + g * (NB / 8) * 256, kPlainLbo, kPlainSbo, kLayoutNone);
          bt_mma(tmem + kDCol, tmem + kACol + 8 * g, db, kIdesc1, g != 0);       // [hi*hi | hi*lo]
          bt_mma(tmem + kDCol, tmem + kACol + C + 8 * g, db, kIdesc2, true);      // lo*hi into the first half
        }
        mma_commit(&bar_accfull);
      }
    }
    __syncwarp();
  } else if (warp == 5) {
    // =============================================================== loader (one warp): 512-byte rows -> padded stages
    int s = 0, it = 0;
    uint32_t ph = 0;
    if (p.tma) {
      // one elected thread, one bulk tensor copy per stage: 512-byte rows move at HBM speed through the TMA unit where
      // the cp.async warp is limited to ~3 TB/s with two CTAs per SM (measured on the forward convolution)
      if (lane == 0) {
        for (int q = 0; q < NSRC; ++q) tma_prefetch_desc(&maps.m[q]);
        for (int ti = 0; ti < my_tiles; ++ti) {
          const long tile = blockIdx.x + (long)ti * gridDim.x;
          const int b = (int)(tile / p.tiles_per_sample);
          const int s0 = (int)((tile - (long)b * p.tiles_per_sample) * 128);
#pragma unroll 1
          for (int q = 0; q < NSRC; ++q) {
            if (it >= NST) mbar_wait(&bar_empty[s], ph ^ 1);
            mbar_expect_tx(&bar_full[s], kBtStageFloats * 4);
            tma_load_3d(ring + s * kBtStageFloats, &maps.m[q], s0, 0, b, &bar_full[s]);
            ++it;
            if (++s == NST) {
              s = 0;
              ph ^= 1;
            }
          }
        }
      }
      __syncwarp();
    } else
    for (int ti = 0; ti < my_tiles; ++ti) {
      const long tile = blockIdx.x + (long)ti * gridDim.x;
      const int b = (int)(tile / p.tiles_per_sample);
      const long s0 = (tile - (long)b * p.tiles_per_sample) * 128 + lane * 4;
      const bool ok = s0 < p.S;  // S % 4 == 0: a 16-byte piece is entirely inside or outside
#pragma unroll 1
      for (int q = 0; q < NSRC; ++q) {
        const float* base = q == 0 ? p.dy : (q == 1 ? p.y : (q == 2 ? p.in1 : p.in2));
        const float* gp = base + (long)b * C * p.S + (ok ? s0 : 0);
        if (it >= NST) mbar_wait(&bar_empty[s], ph ^ 1);
        const uint32_t dst0 = smem_u32(ring + s * kBtStageFloats) + lane * 16;
#pragma unroll 8
        for (int r = 0; r < C; ++r)
          asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst0 + r * (kBtPitch * 4)),
                       "l"(gp + (long)r * p.S), "r"(ok ? 16 : 0)
                       : "memory");
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
    // =============================================================== workers
    const uint32_t lane_base = tmem + ((uint32_t)(warp * 32) << 16);
    // weight-gradient ownership (wgrad.cuh tiling for 128 threads): 2 voxel groups x 64 threads, 3 x 3 outputs each
    const int wg = tid >> 6, wl = tid & 63;
    const int ot = wl >> 3, it8 = wl & 7;
    float2 accW[2][3][3];
    // bias gradient: every thread sums d(pre) of ITS voxel over the tiles (12 packed adds per tile); folding it into
    // the weight-gradient loop cost 9 divergent FADDs per 18 FFMA2 on every warp (7 % of the kernel's instructions)
    float2 accB2[C / 2];
#pragma unroll
    for (int o = 0; o < C / 2; ++o) accB2[o] = make_float2(0.f, 0.f);
#pragma unroll
    for (int h = 0; h < 2; ++h)
#pragma unroll
      for (int q = 0; q < 3; ++q)
#pragma unroll
        for (int r = 0; r < 3; ++r) accW[h][q][r] = make_float2(0.f, 0.f);

    auto wgrad = [&](const float* sx, float2 (&acc)[3][3]) {
      const float* dp0 = sdp + (ot * 3) * kBtPitch + wg * 64;
      const float* x0 = sx + (it8 * 3) * kBtPitch + wg * 64;
#pragma unroll 2
      for (int v = 0; v < 64; v += 4) {
        float4 d[3], x[3];
#pragma unroll
        for (int q = 0; q < 3; ++q) d[q] = *reinterpret_cast<const float4*>(dp0 + q * kBtPitch + v);
#pragma unroll
        for (int r = 0; r < 3; ++r) x[r] = *reinterpret_cast<const float4*>(x0 + r * kBtPitch + v);
#pragma unroll
        for (int q = 0; q < 3; ++q)
#pragma unroll
          for (int r = 0; r < 3; ++r) {
            float2 a = acc[q][r];
            a = ffma2(make_float2(d[q].x, d[q].y), make_float2(x[r].x, x[r].y), a);
            a = ffma2(make_float2(d[q].z, d[q].w), make_float2(x[r].z, x[r].w), a);
            acc[q][r] = a;
          }
      }
    };

    int s = 0;
    uint32_t ph = 0;
    auto next_stage = [&]() {
      if (++s == NST) {
        s = 0;
        ph ^= 1;
      }
    };
    for (int ti = 0; ti < my_tiles; ++ti) {
      const long tile = blockIdx.x + (long)ti * gridDim.x;
      const int b = (int)(tile / p.tiles_per_sample);
      const long sv = (tile - (long)b * p.tiles_per_sample) * 128 + tid;  // this thread's voxel
      const bool valid = sv < p.S;
      const bool live = valid && (sv % p.P) < p.HW;
      // ---- d(pre) = dy * selu'(y)
      const int s_dy = s;
      mbar_wait(&bar_full[s], ph);
      next_stage();
      const int s_y = s;
      mbar_wait(&bar_full[s], ph);
      next_stage();
      uint32_t hi[C], lo[C];
      {
        const float* pdy = ring + s_dy * kBtStageFloats + tid;
        const float* py = ring + s_y * kBtStageFloats + tid;
#pragma unroll
        for (int o = 0; o < C; o += 2) {
          const float d0 = live ? pdy[o * kBtPitch] * selu_grad_from_out(py[o * kBtPitch]) : 0.f;
          const float d1 = live ? pdy[(o + 1) * kBtPitch] * selu_grad_from_out(py[(o + 1) * kBtPitch]) : 0.f;
          hi[o] = __float_as_uint(d0);
          lo[o] = __float_as_uint(tf32_lo(d0));
          hi[o + 1] = __float_as_uint(d1);
          lo[o + 1] = __float_as_uint(tf32_lo(d1));
          sdp[o * kBtPitch + tid] = d0;
          sdp[(o + 1) * kBtPitch + tid] = d1;
          accB2[o / 2] = __fadd2_rn(accB2[o / 2], make_float2(d0, d1));
        }
      }
      // previous tile: its MMAs retired before its epilogue ran (the epilogue waited for them), so the A columns are free
      bt_st16(lane_base + kACol, hi);
      bt_st8(lane_base + kACol + 16, hi + 16);
      bt_st16(lane_base + kACol + C, lo);
      bt_st8(lane_base + kACol + C + 16, lo + 16);
      bt_arrive(&bar_empty[s_dy]);
      bt_arrive(&bar_empty[s_y]);
      asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
      tc_fence_before_sync();
      bt_arrive(&bar_ready);
      bt_worker_sync();  // d(pre) tile complete in shared memory
      // ---- weight gradient (FFMA2) while the tensor core computes the input gradient
      const int s_in1 = s;
      mbar_wait(&bar_full[s], ph);
      next_stage();
      wgrad(ring + s_in1 * kBtStageFloats, accW[0]);
      int s_in2 = 0;
      if (CI2 > 0) {
        s_in2 = s;
        mbar_wait(&bar_full[s], ph);
        next_stage();
        wgrad(ring + s_in2 * kBtStageFloats, accW[1]);
      }
      // ---- epilogue: input gradients from the accumulator, one lane = one voxel.
      // Accumulating destinations (U-Net skip gradients): the old values of a block of 8 rows are requested two blocks
      // ahead of their use and the first two before the accumulator is waited for -- loaded inside the store loop they
      // were a chain of exposed HBM round trips (0.99 ms instead of 0.35 ms for the mapping convolutions).
      const long off = (long)b * C * p.S + sv;
      float oldv[3][8];
      auto fetch_old = [&](int i0, float (&o)[8]) {
        float* dst = (i0 < C ? p.din1 : p.din2);
        const bool accumulate = i0 < C ? (p.flags & 1) : (p.flags & 2);
        if (dst != nullptr && valid && accumulate) {
          const float* q = dst + off + (long)(i0 < C ? i0 : i0 - C) * p.S;
#pragma unroll
          for (int j = 0; j < 8; ++j) o[j] = __ldcs(q + (long)j * p.S);
        }
      };
      if (ACC) {
        fetch_old(0, oldv[0]);
        if (CI > 8) fetch_old(8, oldv[1]);
      }
      mbar_wait(&bar_accfull, (uint32_t)(ti & 1));
      tc_fence_after_sync();
      {
        const uint32_t acc = lane_base + kDCol;
        const float* pin1 = ring + s_in1 * kBtStageFloats + tid;
#pragma unroll
        for (int i0 = 0; i0 < CI; i0 += 8) {
          if (ACC && i0 + 16 < CI) fetch_old(i0 + 16, oldv[(i0 / 8 + 2) % 3]);
          float a[8], c[8];
          bt_ld8(acc + i0, a);
          bt_ld8(acc + NP + i0, c);
          float* dst = (i0 < C ? p.din1 : p.din2);
          const bool accumulate = i0 < C ? (p.flags & 1) : (p.flags & 2);
          if (dst != nullptr && valid) {
            float* q = dst + off + (long)(i0 < C ? i0 : i0 - C) * p.S;
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              float v = a[j] + c[j];
              if (i0 < C && (p.flags & 4)) v *= selu_grad_from_out(pin1[(i0 + j) * kBtPitch]);
              if (ACC && accumulate) v += oldv[(i0 / 8) % 3][j];
              q[(long)j * p.S] = v;
            }
          }
        }
      }
      tc_fence_before_sync();
      bt_arrive(&bar_empty[s_in1]);
      if (CI2 > 0) bt_arrive(&bar_empty[s_in2]);
      bt_worker_sync();  // everyone is done with sdp before the next tile overwrites it
    }
    // ---- per-CTA partial sums: reduce the two voxel groups through shared memory, one row per CTA
    {
      float* prow = p.partials + (long)blockIdx.x * PSTRIDE;
#pragma unroll
      for (int h = 0; h < (CI2 > 0 ? 2 : 1); ++h)
#pragma unroll
        for (int q = 0; q < 3; ++q)
#pragma unroll
          for (int r = 0; r < 3; ++r)
            scratch[wg * (C * CI) + (ot * 3 + q) * CI + h * C + it8 * 3 + r] = accW[h][q][r].x + accW[h][q][r].y;
      {  // bias partials: warp sums, one row of C per worker warp
        const int lane = tid & 31;
#pragma unroll
        for (int o = 0; o < C / 2; ++o) {
          const float bx = warp_sum(accB2[o].x), by = warp_sum(accB2[o].y);
          if (lane == 0) {
            scratch[2 * C * CI + warp * C + 2 * o] = bx;
            scratch[2 * C * CI + warp * C + 2 * o + 1] = by;
          }
        }
      }
      bt_worker_sync();
      for (int idx = tid; idx < C * CI; idx += kBtWorkers) prow[idx] = scratch[idx] + scratch[C * CI + idx];
      for (int o = tid; o < C; o += kBtWorkers)
        prow[C * CI + o] = (scratch[2 * C * CI + o] + scratch[2 * C * CI + C + o]) +
                           (scratch[2 * C * CI + 2 * C + o] + scratch[2 * C * CI + 3 * C + o]);
    }
  }
  tc_fence_before_sync();
  __syncthreads();
  if (warp == 4) tmem_dealloc(tmem, kTmemCols);
}

// ------------------------------------------------------------------------------------------------ role-split variant
// Same tile, ring and arithmetic as k_pwconv_bwd_tc, but the per-voxel work (d(pre), TMEM staging, input-gradient epilogue)
// and the weight gradient run on DIFFERENT warps: in the kernel above 128 threads do both one after the other, so an SM holds
// 8 compute warps that are each inside one latency-bound phase (ncu: 18 % warps active, issue 44 %, DRAM 64 %).  Here a CTA
// has 4 voxel warps + 4 weight-gradient warps; d(pre) is double buffered in shared memory and in tensor memory, so the voxel
// warps form d(pre) of tile t+1 while the tensor core works on tile t and the weight-gradient warps are anywhere inside tile
// t or t-1.  Every ring stage is released by both groups (256 arrivals).
constexpr int kB2Threads = 320;  // warps 0-3 voxel, 4 MMA issuer, 5 loader, 6-9 weight gradient

struct B2Cursor {
  int s;
  uint32_t ph;
  __device__ __forceinline__ void advance(int n, int nst) {
    s += n;
    if (s >= nst) {
      s -= nst;
      ph ^= 1;
    }
  }
  __device__ __forceinline__ B2Cursor plus(int n, int nst) const {
    B2Cursor c = *this;
    c.advance(n, nst);
    return c;
  }
};

template <int CI2, bool ACC>
__global__ void __launch_bounds__(kB2Threads, 2) k_pwconv_bwd_split(const BtDev p, const __grid_constant__ BtMaps maps) {
  constexpr int C = kBtC;
  constexpr int CI = C + CI2;
  constexpr int NP = CI == 48 ? 48 : 32;
  constexpr int NB = 2 * NP;
  constexpr int NSRC = CI2 > 0 ? 4 : 3;
  constexpr uint32_t kIdesc1 = make_idesc_tf32(128, NB, 0, 0);
  constexpr uint32_t kIdesc2 = make_idesc_tf32(128, NP, 0, 0);
  constexpr uint32_t kACol = 0, kDCol = 4 * C;  // TMEM: two A buffers of [hi 24 | lo 24], accumulator [96, 96 + NB)
  constexpr uint32_t kTmemCols = 256;
  constexpr int PSTRIDE = C * CI + C;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  const int NST = p.nst;
  float* ring = reinterpret_cast<float*>(smem);            // [NST][24][132]
  float* bimg = ring + NST * kBtStageFloats;                // [NB rows][24 k]
  float* sbias = bimg + NB * C;                             // [4][24]
  __shared__ __align__(8) uint64_t bar_full[kBtMaxStages];
  __shared__ __align__(8) uint64_t bar_empty[kBtMaxStages];  // 256 arrivals: voxel + weight-gradient warps
  __shared__ __align__(8) uint64_t bar_ready;                // d(pre) of the tile is in tensor memory (128)
  __shared__ __align__(8) uint64_t bar_dfree;                // the accumulator has been read back        (128)
  __shared__ __align__(8) uint64_t bar_accfull;              // tcgen05.commit
  __shared__ __align__(8) uint64_t bar_sdpfull[4];           // d(pre) written over the dy stage of tile t (128, voxel warps);
                                                             // the ring bounds the voxel warps' lead to < 3 tiles
  __shared__ uint32_t tmem_slot;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  for (int idx = tid; idx < NP * C; idx += kB2Threads) {
    const int n = idx / C, k = idx - n * C;
    const float v = n < CI ? __ldg(p.w + k * CI + n) : 0.f;
    const float hi = __uint_as_float((__float_as_uint(v) + 0x1000u) & 0xFFFFE000u);
    bimg[kmajor_plain_index<NB>(n, k)] = hi;
    bimg[kmajor_plain_index<NB>(NP + n, k)] = v - hi;
  }
  if (tid == 0) {
    for (int s = 0; s < kBtMaxStages; ++s) {
      mbar_init(&bar_full[s], 1);
      mbar_init(&bar_empty[s], 256);
    }
    mbar_init(&bar_ready, 128);
    mbar_init(&bar_dfree, 128);
    mbar_init(&bar_accfull, 1);
    for (int i = 0; i < 4; ++i) mbar_init(&bar_sdpfull[i], 128);
    mbar_fence_init();
  }
  if (warp == 4) tmem_alloc(&tmem_slot, kTmemCols);
  fence_proxy_async_smem();
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem = tmem_slot;
  const int my_tiles = p.total_tiles > (int)blockIdx.x ? (p.total_tiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;

  if (warp == 4) {
    // =============================================================== MMA issuer
    if (lane == 0) {
      const uint32_t b0 = smem_u32(bimg);
      for (int ti = 0; ti < my_tiles; ++ti) {
        mbar_wait(&bar_ready, (uint32_t)(ti & 1));
        if (ti > 0) mbar_wait(&bar_dfree, (uint32_t)((ti - 1) & 1));
        tc_fence_after_sync();
        const uint32_t acol = tmem + kACol + (uint32_t)(ti & 1) * 2 * C;
#pragma unroll
        for (int g = 0; g < C / 8; ++g) {
          const uint64_t db = make_smem_desc(b0 + g * (NB / 8) * 256, kPlainLbo, kPlainSbo, kLayoutNone);
          bt_mma(tmem + kDCol, acol + 8 * g, db, kIdesc1, g != 0);
          bt_mma(tmem + kDCol, acol + C + 8 * g, db, kIdesc2, true);
        }
        mma_commit(&bar_accfull);
      }
    }
    __syncwarp();
  } else if (warp == 5) {
    // =============================================================== loader: one bulk tensor copy per stage
    if (lane == 0) {
      int s = 0, it = 0;
      uint32_t ph = 0;
      for (int q = 0; q < NSRC; ++q) tma_prefetch_desc(&maps.m[q]);
      for (int ti = 0; ti < my_tiles; ++ti) {
        const unsigned tile = blockIdx.x + (unsigned)ti * gridDim.x;  // 32-bit: total_tiles and S < 2^31 (launcher)
        const int b = (int)(tile / (unsigned)p.tiles_per_sample);
        const int s0 = (int)((tile - (unsigned)b * (unsigned)p.tiles_per_sample) * 128u);
#pragma unroll 1
        for (int q = 0; q < NSRC; ++q) {
          if (it >= NST) {
            mbar_wait(&bar_empty[s], ph ^ 1);
            fence_proxy_async_smem();  // dy stages are rewritten in place (generic proxy) before the next bulk copy lands
          }
          mbar_expect_tx(&bar_full[s], kBtStageFloats * 4);
          tma_load_3d(ring + s * kBtStageFloats, &maps.m[q], s0, 0, b, &bar_full[s]);
          ++it;
          if (++s == NST) {
            s = 0;
            ph ^= 1;
          }
        }
      }
    }
    __syncwarp();
  } else if (warp < 4) {
    // =============================================================== voxel warps: one thread = one voxel = one TMEM lane
    const uint32_t lane_base = tmem + ((uint32_t)(warp * 32) << 16);
    float2 accB2[C / 2];
#pragma unroll
    for (int o = 0; o < C / 2; ++o) accB2[o] = make_float2(0.f, 0.f);

    // 32-bit index arithmetic (total_tiles, S < 2^31 are checked by the launcher): the 64-bit division / modulo of the
    // first version were ~150 instructions per tile and thread
    const unsigned P32 = (unsigned)p.P, HW32 = (unsigned)p.HW;
    auto voxel_of = [&](int ti, int& b, long& sv) {
      const unsigned tile = blockIdx.x + (unsigned)ti * gridDim.x;
      b = (int)(tile / (unsigned)p.tiles_per_sample);
      sv = (long)((tile - (unsigned)b * (unsigned)p.tiles_per_sample) * 128u + (unsigned)tid);
    };
    // d(pre) of tile ti: registers -> tensor memory (two TF32 terms) and, in fp32, over the dy stage it came from (a thread
    // owns its voxel's column, so the overwrite is race free; a separate d(pre) buffer cost two ring stages and the kernel
    // is sensitive to the ring depth: 0.266 ms with 6 stages, 0.301 ms with 5)
    auto dpre = [&](int ti, B2Cursor cur) {
      int b;
      long sv;
      voxel_of(ti, b, sv);
      const bool live = sv < p.S && ((unsigned)sv % P32) < HW32;
      const int buf = ti & 1;
      const B2Cursor cy = cur.plus(1, NST);
      mbar_wait(&bar_full[cur.s], cur.ph);
      mbar_wait(&bar_full[cy.s], cy.ph);
      const float* pdy = ring + cur.s * kBtStageFloats + tid;
      const float* py = ring + cy.s * kBtStageFloats + tid;
      float* ps = ring + cur.s * kBtStageFloats + tid;
      const uint32_t acol = lane_base + kACol + (uint32_t)buf * 2 * C;
#pragma unroll
      for (int o0 = 0; o0 < C; o0 += 8) {
        uint32_t hi[8], lo[8];
#pragma unroll
        for (int j = 0; j < 8; j += 2) {
          const int o = o0 + j;
          const float d0 = live ? pdy[o * kBtPitch] * selu_grad_from_out(py[o * kBtPitch]) : 0.f;
          const float d1 = live ? pdy[(o + 1) * kBtPitch] * selu_grad_from_out(py[(o + 1) * kBtPitch]) : 0.f;
          hi[j] = __float_as_uint(d0);
          lo[j] = __float_as_uint(tf32_lo(d0));
          hi[j + 1] = __float_as_uint(d1);
          lo[j + 1] = __float_as_uint(tf32_lo(d1));
          ps[o * kBtPitch] = d0;
          ps[(o + 1) * kBtPitch] = d1;
          accB2[o / 2] = __fadd2_rn(accB2[o / 2], make_float2(d0, d1));
        }
        bt_st8(acol + o0, hi);
        bt_st8(acol + C + o0, lo);
      }
      fence_proxy_async_smem();
      bt_arrive(&bar_empty[cur.s]);
      bt_arrive(&bar_empty[cy.s]);
      asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
      tc_fence_before_sync();
      bt_arrive(&bar_ready);
      bt_arrive(&bar_sdpfull[ti & 3]);
    };

    B2Cursor cur{0, 0};  // first stage (dy) of the tile whose epilogue runs next
    if (my_tiles > 0) dpre(0, cur);
    for (int ti = 0; ti < my_tiles; ++ti) {
      if (ti + 1 < my_tiles) dpre(ti + 1, cur.plus(NSRC, NST));
      // ---- epilogue of tile ti
      int b;
      long sv;
      voxel_of(ti, b, sv);
      const bool valid = sv < p.S;
      const B2Cursor c1 = cur.plus(2, NST);
      const long off = (long)b * C * p.S + sv;
      float oldv[3][8];
      auto fetch_old = [&](int i0, float (&o)[8]) {
        float* dst = (i0 < C ? p.din1 : p.din2);
        const bool accumulate = i0 < C ? (p.flags & 1) : (p.flags & 2);
        if (dst != nullptr && valid && accumulate) {
          const float* q = dst + off + (long)(i0 < C ? i0 : i0 - C) * p.S;
#pragma unroll
          for (int j = 0; j < 8; ++j) o[j] = __ldcs(q + (long)j * p.S);
        }
      };
      if (ACC) {
        fetch_old(0, oldv[0]);
        if (CI > 8) fetch_old(8, oldv[1]);
      }
      // the voxel warps read in1 only for the selu'(in1) factor and in2 never, but they release every stage like the other
      // group: arrive strictly after the stage has landed, so that an arrival can never be counted in the slot's previous phase
      mbar_wait(&bar_full[c1.s], c1.ph);
      if (CI2 > 0) {
        const B2Cursor c2w = c1.plus(1, NST);
        mbar_wait(&bar_full[c2w.s], c2w.ph);
      }
      mbar_wait(&bar_accfull, (uint32_t)(ti & 1));
      tc_fence_after_sync();
      {
        const uint32_t acc = lane_base + kDCol;
        const float* pin1 = ring + c1.s * kBtStageFloats + tid;
#pragma unroll
        for (int i0 = 0; i0 < CI; i0 += 8) {
          if (ACC && i0 + 16 < CI) fetch_old(i0 + 16, oldv[(i0 / 8 + 2) % 3]);
          float a[8], c[8];
          bt_ld8(acc + i0, a);
          bt_ld8(acc + NP + i0, c);
          if (i0 + 8 >= CI) {  // the accumulator is in registers: the next tile's MMAs may overwrite it
            tc_fence_before_sync();
            bt_arrive(&bar_dfree);
          }
          float* dst = (i0 < C ? p.din1 : p.din2);
          const bool accumulate = i0 < C ? (p.flags & 1) : (p.flags & 2);
          if (dst != nullptr && valid) {
            float* q = dst + off + (long)(i0 < C ? i0 : i0 - C) * p.S;
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              float v = a[j] + c[j];
              if (i0 < C && (p.flags & 4)) v *= selu_grad_from_out(pin1[(i0 + j) * kBtPitch]);
              if (ACC && accumulate) v += oldv[(i0 / 8) % 3][j];
              q[(long)j * p.S] = v;
            }
          }
        }
      }
      bt_arrive(&bar_empty[c1.s]);
      if (CI2 > 0) bt_arrive(&bar_empty[c1.plus(1, NST).s]);
      cur.advance(NSRC, NST);
    }
    // ---- bias partials: warp sums, one row of C per voxel warp
#pragma unroll
    for (int o = 0; o < C / 2; ++o) {
      const float bx = warp_sum(accB2[o].x), by = warp_sum(accB2[o].y);
      if (lane == 0) {
        sbias[warp * C + 2 * o] = bx;
        sbias[warp * C + 2 * o + 1] = by;
      }
    }
    bt_worker_sync();
    float* prow = p.partials + (long)blockIdx.x * PSTRIDE;
    for (int o = tid; o < C; o += 128)
      prow[C * CI + o] = (sbias[o] + sbias[C + o]) + (sbias[2 * C + o] + sbias[3 * C + o]);
  } else {
    // =============================================================== weight-gradient warps
    const int gt = tid - 192;
    const int wg = gt >> 6, wl = gt & 63;
    const int ot = wl >> 3, it8 = wl & 7;
    float2 accW[2][3][3];
#pragma unroll
    for (int h = 0; h < 2; ++h)
#pragma unroll
      for (int q = 0; q < 3; ++q)
#pragma unroll
        for (int r = 0; r < 3; ++r) accW[h][q][r] = make_float2(0.f, 0.f);

    auto wgrad = [&](const float* sd, const float* sx, float2 (&acc)[3][3]) {
      const float* dp0 = sd + (ot * 3) * kBtPitch + wg * 64;
      const float* x0 = sx + (it8 * 3) * kBtPitch + wg * 64;
#pragma unroll 2
      for (int v = 0; v < 64; v += 4) {
        float4 d[3], x[3];
#pragma unroll
        for (int q = 0; q < 3; ++q) d[q] = *reinterpret_cast<const float4*>(dp0 + q * kBtPitch + v);
#pragma unroll
        for (int r = 0; r < 3; ++r) x[r] = *reinterpret_cast<const float4*>(x0 + r * kBtPitch + v);
#pragma unroll
        for (int q = 0; q < 3; ++q)
#pragma unroll
          for (int r = 0; r < 3; ++r) {
            float2 a = acc[q][r];
            a = ffma2(make_float2(d[q].x, d[q].y), make_float2(x[r].x, x[r].y), a);
            a = ffma2(make_float2(d[q].z, d[q].w), make_float2(x[r].z, x[r].w), a);
            acc[q][r] = a;
          }
      }
    };

    // two sources: one pass over the voxels feeds both weight-gradient halves, so the d(pre) rows are read once (the loop is
    // bound by shared-memory wavefronts -- ncu: LSU data pipe 84 % with separate passes -- and this removes a quarter of them)
    auto wgrad2 = [&](const float* sd, const float* sx1, const float* sx2) {
      const float* dp0 = sd + (ot * 3) * kBtPitch + wg * 64;
      const float* x1 = sx1 + (it8 * 3) * kBtPitch + wg * 64;
      const float* x2 = sx2 + (it8 * 3) * kBtPitch + wg * 64;
#pragma unroll 2
      for (int v = 0; v < 64; v += 4) {
        float4 d[3], x[6];
#pragma unroll
        for (int q = 0; q < 3; ++q) d[q] = *reinterpret_cast<const float4*>(dp0 + q * kBtPitch + v);
#pragma unroll
        for (int r = 0; r < 3; ++r) {
          x[r] = *reinterpret_cast<const float4*>(x1 + r * kBtPitch + v);
          x[3 + r] = *reinterpret_cast<const float4*>(x2 + r * kBtPitch + v);
        }
#pragma unroll
        for (int q = 0; q < 3; ++q)
#pragma unroll
          for (int r = 0; r < 6; ++r) {
            float2 a = accW[r / 3][q][r % 3];
            a = ffma2(make_float2(d[q].x, d[q].y), make_float2(x[r].x, x[r].y), a);
            a = ffma2(make_float2(d[q].z, d[q].w), make_float2(x[r].z, x[r].w), a);
            accW[r / 3][q][r % 3] = a;
          }
      }
    };

    B2Cursor cur{0, 0};
    int last_dy = 0;
    for (int ti = 0; ti < my_tiles; ++ti) {
      const float* sd = ring + cur.s * kBtStageFloats;  // d(pre), written over dy by the voxel warps
      mbar_wait(&bar_sdpfull[ti & 3], (uint32_t)((ti >> 2) & 1));
      bt_arrive(&bar_empty[cur.plus(1, NST).s]);  // y was consumed by the voxel warps before they signalled
      const B2Cursor c1 = cur.plus(2, NST);
      mbar_wait(&bar_full[c1.s], c1.ph);
      if (CI2 > 0) {
        const B2Cursor c2 = c1.plus(1, NST);
        mbar_wait(&bar_full[c2.s], c2.ph);
        wgrad2(sd, ring + c1.s * kBtStageFloats, ring + c2.s * kBtStageFloats);
        bt_arrive(&bar_empty[c1.s]);
        bt_arrive(&bar_empty[c2.s]);
      } else {
        wgrad(sd, ring + c1.s * kBtStageFloats, accW[0]);
        bt_arrive(&bar_empty[c1.s]);
      }
      if (ti + 1 < my_tiles) bt_arrive(&bar_empty[cur.s]);  // the last tile's stage becomes the reduction scratch
      last_dy = cur.s;
      cur.advance(NSRC, NST);
    }
    // ---- per-CTA partial sums; the last d(pre) stage is free once every warp of this group has read it
    float* scratch = ring + last_dy * kBtStageFloats;
    float* prow = p.partials + (long)blockIdx.x * PSTRIDE;
    asm volatile("bar.sync 2, 128;" ::: "memory");  // every warp of the group has read its last d(pre) tile
#pragma unroll
    for (int h = 0; h < (CI2 > 0 ? 2 : 1); ++h)
#pragma unroll
      for (int q = 0; q < 3; ++q)
#pragma unroll
        for (int r = 0; r < 3; ++r)
          scratch[wg * (C * CI) + (ot * 3 + q) * CI + h * C + it8 * 3 + r] = accW[h][q][r].x + accW[h][q][r].y;
    asm volatile("bar.sync 2, 128;" ::: "memory");
    for (int idx = gt; idx < C * CI; idx += 128) prow[idx] = scratch[idx] + scratch[C * CI + idx];
  }
  tc_fence_before_sync();
  __syncthreads();
  if (warp == 4) tmem_dealloc(tmem, kTmemCols);
}

// ------------------------------------------------------------------------------------------------ host side
bool pwconv_bwd_tc_eligible(const float* dy, const float* y, const float* in1, const float* in2, const float* din1,
                            const float* din2, int ci1, int ci2, int co, long S, int act, int residual) {
  if (!tc_enabled()) return false;
  static const bool off = getenv("HNO_PWBWD_TC") && atoi(getenv("HNO_PWBWD_TC")) == 0;
  if (off) return false;
  if (ci1 != kBtC || co != kBtC || (ci2 != 0 && ci2 != kBtC) || act != 1 || residual) return false;
  if (S < 4096 || S % 4) return false;
  const void* ptrs[6] = {dy, y, in1, in2, din1, din2};
  for (const void* q : ptrs)
    if (q && reinterpret_cast<uintptr_t>(q) % 16) return false;
  return true;
}

template <int CI2, bool ACC>
static int launch_bt(BtDev p, cudaStream_t st, int* grid_out) {
  constexpr int CI = kBtC + CI2;
  constexpr int NB = 2 * (CI == 48 ? 48 : 32);
  // HNO_PWBWD_SPLIT=0: the single-role kernel (128 workers do the per-voxel work and the weight gradient in turn)
  static const bool split = !(getenv("HNO_PWBWD_SPLIT") && atoi(getenv("HNO_PWBWD_SPLIT")) == 0);
  const size_t fixed = split ? 1024 + (size_t)(NB * kBtC + 4 * kBtC + 64) * sizeof(float)
                             : 1024 + (size_t)(kBtStageFloats + NB * kBtC + 2 * kBtC * CI + 2 * kBtC + 64) * sizeof(float);
  int nst = (int)((233472 / 2 - 2 * 1024 - fixed) / (kBtStageFloats * sizeof(float)));
  static const int nst_env = getenv("HNO_PWBWD_NST") ? atoi(getenv("HNO_PWBWD_NST")) : 0;
  if (nst_env > 0) nst = nst_env;
  if (nst > kBtMaxStages) nst = kBtMaxStages;
  HNO_CHECK(nst >= 5, "pwconv_bwd_tc: not enough shared memory for the ring");
  p.nst = nst;
  const size_t smem = fixed + (size_t)nst * kBtStageFloats * sizeof(float);
  long grid = (long)sm_count() * 2;  // two CTAs per SM (256 TMEM columns and ~100 KB of shared memory each)
  if (grid > p.total_tiles) grid = p.total_tiles;
  *grid_out = (int)grid;
  BtMaps maps;
  {
    static const bool tma_on = !(getenv("HNO_PWBWD_TMA") && atoi(getenv("HNO_PWBWD_TMA")) == 0);
    p.tma = tma_on && p.S < (1L << 31) ? 1 : 0;
    const float* src[4] = {p.dy, p.y, p.in1, CI2 > 0 ? p.in2 : p.in1};
    const int B = p.total_tiles / p.tiles_per_sample;
    for (int q = 0; q < 4; ++q) {
      const uint64_t dims[3] = {(uint64_t)p.S, (uint64_t)kBtC, (uint64_t)B};
      const uint64_t strides[2] = {(uint64_t)p.S * 4, (uint64_t)kBtC * p.S * 4};
      const uint32_t box[3] = {(uint32_t)kBtPitch, (uint32_t)kBtC, 1};
      if (p.tma)
        if (int rc = encode_tensor_map(&maps.m[q], src[q], 3, dims, strides, box, 0)) return rc;
    }
  }
  if (split && p.tma) {
    auto kern = k_pwconv_bwd_split<CI2, ACC>;
    HNO_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kern<<<(int)grid, kB2Threads, smem, st>>>(p, maps);
  } else {
    HNO_CHECK(!split, "pwconv_bwd_tc: the role-split kernel needs the TMA ring (HNO_PWBWD_TMA=0 requires HNO_PWBWD_SPLIT=0)");
    auto kern = k_pwconv_bwd_tc<CI2, ACC>;
    HNO_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kern<<<(int)grid, kBtThreads, smem, st>>>(p, maps);
  }
  HNO_LAUNCH_CHECK();
  return 0;
}

int pwconv_bwd_tc(const float* dy, const float* y, const float* in1, const float* in2, const float* w, float* din1,
                  float* din2, float* dweight, float* dbias, void* ws, int B, int ci2, long S, long P, long HW, int flags,
                  cudaStream_t st) {
  BtDev p;
  p.dy = dy;
  p.y = y;
  p.in1 = in1;
  p.in2 = in2;
  p.w = w;
  p.din1 = din1;
  p.din2 = din2;
  p.partials = reinterpret_cast<float*>(ws);
  p.S = S;
  p.P = P;
  p.HW = HW;
  p.tiles_per_sample = ceil_div(S, 128);
  p.total_tiles = p.tiles_per_sample * B;
  p.flags = flags;
  int grid = 0;
  const bool accum = (flags & 3) != 0;  // an input gradient is accumulated into its destination (U-Net skips)
  int rc;
  if (ci2 > 0)
    rc = accum ? launch_bt<kBtC, true>(p, st, &grid) : launch_bt<kBtC, false>(p, st, &grid);
  else
    rc = accum ? launch_bt<0, true>(p, st, &grid) : launch_bt<0, false>(p, st, &grid);
  if (rc) return rc;
  const int ci = kBtC + ci2;
  return reduce_partials(p.partials, grid, kBtC * ci, kBtC, dweight, dbias, (flags & 8) ? 1 : 0, st);
}

}  // namespace hno
