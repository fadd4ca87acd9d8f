// Host interface of the streamed tensor-core contraction (tc_stream.cu).
#pragma once
#include <cuda_runtime.h>

namespace hno {

// out[g][n][m] (op)= act( scale * sum_k A[g][k][m] * B[n][k] + bias[n] )
//   A: nsrc (1 or 2) tensors [G][rows][mext] (row stride lda, slab stride gsa, floats); the K axis is the virtual
//      concatenation of chunks_per_src * kc rows of each source (rows beyond rows[i] read as zero).
//   B: b[n * ldbn + k * ldbk], n < nout, k < kvalid (zero beyond).
//   out: out[g * gso + n * ldo + m];  columns m >= valid_m are written as 0 (epi 0) or left untouched (epi 1).
struct TcStreamArgs {
  const float* a[2];
  long lda[2], gsa[2];
  int rows[2];
  int nsrc;
  long mext;
  int G;
  int kc;              // rows per chunk: 8, 24 or 32
  int chunks_per_src;
  const float* b;
  long ldbn, ldbk;
  int kvalid;
  float scale;
  const float* bias;   // [nout] or null
  float* out;
  long ldo, gso;
  int nout;
  long valid_m;
  int act;             // 0 none, 1 SELU
  int epi;             // 0 store, 1 accumulate
  int loader;          // producer of the shared-memory ring: 0 = cp.async warp (fastest for 24/48-row operands),
                       // 1 = TMA (fastest for the 121-row D-axis analysis); HNO_TC_LOADER overrides
  // Optional row re-mapping of the contiguous axis (0 = off): column m lives at (m / rw) * rp + m % rw instead of m.
  // Used by the truncated DHT to keep its D-stage intermediate as [h][jd][w] (the H stage then streams it with
  // (jd, w) as the contiguous axis).  out_*: applied by the epilogue (columns >= valid_m are then not written);
  // in_*: applied by the cp.async loader, 8-byte pieces (rw, rp, lda even; columns >= valid_m read as zero).
  int out_rw;
  long out_rp;
  int in_rw;
  long in_rp;
};

bool tc_stream_eligible(const TcStreamArgs& a);
int tc_stream_launch(const TcStreamArgs& a, cudaStream_t st);
// analysis stages of the transform (kc = 16, <= 32 output rows, one source) with the streamed operand fed through tensor
// memory (tc_analysis.cu); tc_stream_launch routes to it when eligible (HNO_TC_ANALYSIS=0 keeps the shared-memory-operand ring)
bool tc_analysis_eligible(const TcStreamArgs& a);
int tc_analysis_launch(const TcStreamArgs& a, cudaStream_t st);
// Global switch (tests / A-B measurements): returns the previous value.
int tc_set_enabled(int on);
bool tc_enabled();

}  // namespace hno
