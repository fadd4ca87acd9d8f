// Host interface of the batched TN GEMM (gemm_tc.cu).
#pragma once
#include <cuda_runtime.h>

namespace hno {

// C[b][m][n] = EPI( alpha * sum_k A[b][m][k] * B[b][n][k] ); element strides ld*, batch strides s* (floats).
struct GemmArgs {
  const float* a;   // A[b][m][k]
  long lda, sa;
  const float* b;   // B[b][n][k]
  long ldb, sb;
  float* c;         // C[b][m][n] or null
  long ldc, sc;
  float* ct;        // C^T[b][n][m] or null
  long ldct, sct;
  const float* e;   // epilogue operand E[b][m][n] (epi 2) or null
  long lde, se;
  int batch, M, N, K;
  float alpha;
  int epi;          // 0: alpha * acc   1: selu(alpha * acc)   2: alpha * acc * selu'(.) evaluated from the SELU OUTPUT E
};

int gemm_tn(const GemmArgs& g, cudaStream_t st);

}  // namespace hno
