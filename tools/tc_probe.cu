// tcgen05 / TMA descriptor probe (development tool, not part of the library).
//
// Runs ONE tcgen05.mma tile per configuration with operands fetched by TMA and dumps the raw TMEM accumulator, so
// that the shared-memory descriptor encodings used by csrc/tc_stream.cuh (layout type, LBO/SBO, swizzle pairing
// between the TMA tensor map and the UMMA descriptor) are verified on real hardware instead of taken on faith.
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o tools/tc_probe tools/tc_probe.cu && tools/tc_probe
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <vector>

#define CK(x)                                                                          \
  do {                                                                                 \
    cudaError_t e_ = (x);                                                              \
    if (e_ != cudaSuccess) {                                                           \
      printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); \
      exit(1);                                                                         \
    }                                                                                  \
  } while (0)

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode() {
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult q;
  CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q));
  if (!fn) {
    printf("no cuTensorMapEncodeTiled\n");
    exit(1);
  }
  return (EncodeTiledFn)fn;
}

struct ProbeCfg {
  int nbox;            // TMA boxes to load
  int box_c0[8];       // inner coordinate of each box
  int box_c1[8];       // outer coordinate of each box
  int box_bytes;       // bytes per box
  uint32_t a_lbo, a_sbo, a_layout;  // bytes, bytes, layout type
  int a_kstep;         // bytes added to the A start address per MMA
  int nmma;
  uint32_t idesc;
  int b_bytes;         // bytes of the pre-arranged B image
  uint32_t b_lbo, b_sbo;
  int b_kstep;
  int ncols;           // accumulator columns to dump
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ uint64_t make_desc(uint32_t addr, uint32_t lbo, uint32_t sbo, uint32_t layout) {
  uint64_t d = 0;
  d |= (uint64_t)((addr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;  // descriptor version (Blackwell)
  d |= (uint64_t)(layout & 7) << 61;
  return d;
}

__global__ void __launch_bounds__(128) k_probe(const __grid_constant__ CUtensorMap tmap, ProbeCfg cfg,
                                               const float* __restrict__ bimg, float* __restrict__ out) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint8_t* sA = smem;                  // up to 64 KB
  uint8_t* sB = smem + 64 * 1024;      // up to 32 KB
  __shared__ __align__(8) uint64_t bar_full, bar_mma;
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5;

  for (int i = tid; i < cfg.b_bytes / 4; i += 128) ((float*)sB)[i] = bimg[i];
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar_full)));
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar_mma)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)),
                 "r"(256));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic writes of B -> visible to the MMA
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tmem_base_s;

  if (tid == 0) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&bar_full)),
                 "r"(cfg.nbox * cfg.box_bytes)
                 : "memory");
    for (int b = 0; b < cfg.nbox; ++b) {
      asm volatile(
          "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::
              "r"(smem_u32(sA + (size_t)b * cfg.box_bytes)),
          "l"(&tmap), "r"(cfg.box_c0[b]), "r"(cfg.box_c1[b]), "r"(smem_u32(&bar_full))
          : "memory");
    }
    // wait for the bytes
    uint32_t ok = 0;
    while (!ok) {
      asm volatile(
          "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
          : "=r"(ok)
          : "r"(smem_u32(&bar_full)), "r"(0)
          : "memory");
    }
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    for (int k = 0; k < cfg.nmma; ++k) {
      const uint64_t da = make_desc(smem_u32(sA) + k * cfg.a_kstep, cfg.a_lbo, cfg.a_sbo, cfg.a_layout);
      const uint64_t db = make_desc(smem_u32(sB) + k * cfg.b_kstep, cfg.b_lbo, cfg.b_sbo, 0);
      const uint32_t acc = k > 0 ? 1u : 0u;
      asm volatile(
          "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
          "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem),
          "l"(da), "l"(db), "r"(cfg.idesc), "r"(acc)
          : "memory");
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(
                     smem_u32(&bar_mma))
                 : "memory");
  }
  // everyone waits for the MMA
  {
    uint32_t ok = 0;
    while (!ok) {
      asm volatile(
          "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
          : "=r"(ok)
          : "r"(smem_u32(&bar_mma)), "r"(0)
          : "memory");
    }
  }
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  for (int c0 = 0; c0 < cfg.ncols; c0 += 32) {
    uint32_t r[32];
    const uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16) + c0;
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,"
        "%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
          "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
          "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    for (int j = 0; j < 32; ++j) out[(size_t)tid * cfg.ncols + c0 + j] = __uint_as_float(r[j]);
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(256));
}

static uint32_t make_idesc(int M, int N, int a_mn, int b_mn) {
  uint32_t d = 0;
  d |= 1u << 4;   // D format f32
  d |= 2u << 7;   // A tf32
  d |= 2u << 10;  // B tf32
  d |= (uint32_t)a_mn << 15;
  d |= (uint32_t)b_mn << 16;
  d |= (uint32_t)(N >> 3) << 17;
  d |= (uint32_t)(M >> 4) << 24;
  return d;
}

// B image: K-major, no swizzle.  element (n, k): (k/8)*kstep + ((k%8)/4)*lbo + (n/8)*sbo + (n%8)*16 + (k%4)*4
static void build_b(std::vector<float>& img, const std::vector<float>& B, int N, int K, int lbo, int sbo, int kstep) {
  img.assign((size_t)(K / 8) * kstep / 4, 0.f);
  for (int n = 0; n < N; ++n)
    for (int k = 0; k < K; ++k) {
      size_t off = (size_t)(k / 8) * kstep + ((k % 8) / 4) * lbo + (n / 8) * sbo + (n % 8) * 16 + (k % 4) * 4;
      img[off / 4] = B[(size_t)n * K + k];
    }
}

int main() {
  EncodeTiledFn encode = get_encode();
  int dev = 0;
  CK(cudaSetDevice(dev));
  cudaDeviceProp prop;
  CK(cudaGetDeviceProperties(&prop, dev));
  printf("device %s cc %d.%d\n", prop.name, prop.major, prop.minor);
  CK(cudaFuncSetAttribute(k_probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));

  // global A source: R rows x LD floats; value = row * 128 + col (exact in tf32 for row < 16, col < 128)
  const int R = 64, LD = 4096;
  std::vector<float> hA((size_t)R * LD);
  for (int r = 0; r < R; ++r)
    for (int c = 0; c < LD; ++c) hA[(size_t)r * LD + c] = (float)((r % 16) * 128 + (c % 128)) + (r >= 16 ? 0.f : 0.f);
  float* dA;
  CK(cudaMalloc(&dA, hA.size() * 4));
  CK(cudaMemcpy(dA, hA.data(), hA.size() * 4, cudaMemcpyHostToDevice));
  float *dB, *dOut;
  CK(cudaMalloc(&dB, 64 * 1024));
  CK(cudaMalloc(&dOut, 128 * 256 * 4));

  struct Variant {
    const char* name;
    CUtensorMapSwizzle swz;
    uint32_t layout, sbo;
  };
  // ------------------------------------------------------------------ P1: MN-major A (voxels = M contiguous)
  {
    const int KC = 16, N = 32, M = 128;  // 2 MMAs
    Variant vars[] = {{"MN-major  TMA 128B_ATOM_32B + layout 1 (SW128_BASE32B), SBO 512",
                       CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B, 1, 512},
                      {"MN-major  TMA 128B + layout 2 (SW128), SBO 1024", CU_TENSOR_MAP_SWIZZLE_128B, 2, 1024},
                      {"MN-major  TMA 128B_ATOM_32B + layout 1, SBO 1024", CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B, 1,
                       1024},
                      {"MN-major  TMA 128B + layout 1, SBO 512", CU_TENSOR_MAP_SWIZZLE_128B, 1, 512}};
    for (auto& v : vars) {
      CUtensorMap tm;
      cuuint64_t dims[2] = {(cuuint64_t)LD, (cuuint64_t)KC};
      cuuint64_t strides[1] = {(cuuint64_t)LD * 4};
      cuuint32_t box[2] = {32, (cuuint32_t)KC};
      cuuint32_t es[2] = {1, 1};
      CUresult rc = encode(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, dA, dims, strides, box, es,
                           CU_TENSOR_MAP_INTERLEAVE_NONE, v.swz, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                           CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
      if (rc != CUDA_SUCCESS) {
        printf("[%s] encode failed rc=%d\n", v.name, (int)rc);
        continue;
      }
      ProbeCfg cfg;
      memset(&cfg, 0, sizeof(cfg));
      cfg.nbox = 4;
      for (int j = 0; j < 4; ++j) {
        cfg.box_c0[j] = 32 * j;
        cfg.box_c1[j] = 0;
      }
      cfg.box_bytes = KC * 128;
      cfg.a_lbo = KC * 128;
      cfg.a_sbo = v.sbo;
      cfg.a_layout = v.layout;
      cfg.a_kstep = 1024;
      cfg.nmma = KC / 8;
      cfg.idesc = make_idesc(M, N, 1, 0);
      cfg.b_lbo = 128;
      cfg.b_sbo = 256;
      cfg.b_kstep = (N / 8) * 256;
      cfg.ncols = N;
      std::vector<float> B((size_t)N * KC, 0.f), img;
      for (int n = 0; n < N; ++n) B[(size_t)n * KC + (n % KC)] = 1.f;  // D[m][n] = A(m, n % KC)
      build_b(img, B, N, KC, cfg.b_lbo, cfg.b_sbo, cfg.b_kstep);
      cfg.b_bytes = (int)img.size() * 4;
      CK(cudaMemcpy(dB, img.data(), img.size() * 4, cudaMemcpyHostToDevice));
      CK(cudaMemset(dOut, 0xff, 128 * 256 * 4));
      k_probe<<<1, 128, 100 * 1024>>>(tm, cfg, dB, dOut);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) {
        printf("[%s] kernel failed: %s\n", v.name, cudaGetErrorString(e));
        return 1;
      }
      std::vector<float> out((size_t)128 * N);
      CK(cudaMemcpy(out.data(), dOut, out.size() * 4, cudaMemcpyDeviceToHost));
      int good = 0;
      for (int m = 0; m < 128; ++m)
        for (int n = 0; n < N; ++n) {
          float want = (float)((n % KC) * 128 + m);
          good += out[(size_t)m * N + n] == want;
        }
      printf("[%s] %d / %d correct\n", v.name, good, 128 * N);
      if (good != 128 * N) {
        printf("   first rows (decoded as k:m):\n");
        for (int m = 0; m < 128; m += 9) {
          printf("   m=%3d:", m);
          for (int n = 0; n < 16; ++n) {
            int val = (int)out[(size_t)m * N + n];
            printf(" %2d:%3d", val / 128, val % 128);
          }
          printf("\n");
        }
      }
    }
  }
  // ------------------------------------------------------------------ P2: K-major A (channels = M rows, voxels = K)
  {
    const int M = 64, N = 32, K = 32;  // 4 MMAs, one 128-byte swizzle row per channel
    struct V2 {
      const char* name;
      CUtensorMapSwizzle swz;
      uint32_t layout, sbo;
    } vars[] = {{"K-major   TMA 128B + layout 2 (SW128), SBO 1024, +32 B per MMA", CU_TENSOR_MAP_SWIZZLE_128B, 2, 1024},
                {"K-major   TMA 128B_ATOM_32B + layout 1, SBO 1024, +32 B per MMA", CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B,
                 1, 1024}};
    for (auto& v : vars) {
      CUtensorMap tm;
      cuuint64_t dims[2] = {(cuuint64_t)LD, (cuuint64_t)48};  // 48 valid rows: rows 48..63 are zero-filled
      cuuint64_t strides[1] = {(cuuint64_t)LD * 4};
      cuuint32_t box[2] = {32, 64};
      cuuint32_t es[2] = {1, 1};
      CUresult rc = encode(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, dA, dims, strides, box, es,
                           CU_TENSOR_MAP_INTERLEAVE_NONE, v.swz, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                           CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
      if (rc != CUDA_SUCCESS) {
        printf("[%s] encode failed rc=%d\n", v.name, (int)rc);
        continue;
      }
      ProbeCfg cfg;
      memset(&cfg, 0, sizeof(cfg));
      cfg.nbox = 1;
      cfg.box_c0[0] = 64;  // voxels 64..95
      cfg.box_c1[0] = 0;
      cfg.box_bytes = 64 * 128;
      cfg.a_lbo = 16;
      cfg.a_sbo = v.sbo;
      cfg.a_layout = v.layout;
      cfg.a_kstep = 32;
      cfg.nmma = K / 8;
      cfg.idesc = make_idesc(M, N, 0, 0);
      cfg.b_lbo = 128;
      cfg.b_sbo = 256;
      cfg.b_kstep = (N / 8) * 256;
      cfg.ncols = N;
      std::vector<float> B((size_t)N * K, 0.f), img;
      for (int n = 0; n < N; ++n) B[(size_t)n * K + n] = 1.f;  // D[r][n] = A(r, k = n)
      build_b(img, B, N, K, cfg.b_lbo, cfg.b_sbo, cfg.b_kstep);
      cfg.b_bytes = (int)img.size() * 4;
      CK(cudaMemcpy(dB, img.data(), img.size() * 4, cudaMemcpyHostToDevice));
      CK(cudaMemset(dOut, 0xff, 128 * 256 * 4));
      k_probe<<<1, 128, 100 * 1024>>>(tm, cfg, dB, dOut);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) {
        printf("[%s] kernel failed: %s\n", v.name, cudaGetErrorString(e));
        return 1;
      }
      std::vector<float> out((size_t)128 * N);
      CK(cudaMemcpy(out.data(), dOut, out.size() * 4, cudaMemcpyDeviceToHost));
      // expected: row r (channel), column n: value (r%16)*128 + (64 + n) for r < 48, else 0.
      // M = 64 accumulator rows may sit in lanes {0..15, 32..47, 64..79, 96..111} or 0..63: report both hypotheses
      int good_a = 0, good_b = 0;
      for (int r = 0; r < 64; ++r)
        for (int n = 0; n < N; ++n) {
          float want = r < 48 ? (float)((r % 16) * 128 + 64 + n) : 0.f;
          good_a += out[(size_t)((r / 16) * 32 + r % 16) * N + n] == want;
          good_b += out[(size_t)r * N + n] == want;
        }
      printf("[%s] lanes (r/16)*32+r%%16: %d / %d   lanes r: %d / %d\n", v.name, good_a, 64 * N, good_b, 64 * N);
      if (good_a != 64 * N && good_b != 64 * N) {
        for (int l = 0; l < 128; l += 5) {
          printf("   lane=%3d:", l);
          for (int n = 0; n < 12; ++n) {
            int val = (int)out[(size_t)l * N + n];
            printf(" %2d:%3d", val / 128, val % 128);
          }
          printf("\n");
        }
      }
    }
  }
  printf("probe done\n");
  return 0;
}
