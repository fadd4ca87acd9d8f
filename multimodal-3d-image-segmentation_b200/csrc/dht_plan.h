// Plan (table blob) for the truncated 3-D discrete Hartley transform.
//
// Reference semantics being replaced (read-only citations):
//   nets/dht.py:16-36          dhtn = Re(fftn) - Im(fftn), 1/N forward, unnormalised inverse
//   nets/hnosegxs.py:378-410   TransformCrop: keep k in [0,m) U [n-m,n) per axis
//   nets/hnosegxs.py:454-494   PadInverse: zero-pad the same corners, inverse dhtn
//
// The transform is evaluated WITHOUT an FFT and WITHOUT complex numbers.  For a
// retained frequency k on an axis of length n let s = k (k <= n/2) or k-n, u = |s|,
// sigma = sign(s).  cas(a+b+c) expands into eight products of cos/sin of the three
// per-axis angles, so
//     Z[kd,kh,kw] = scale * sum_{8 terms} (+/-) sigma.. * T[jd][jh][jw]
//     T[jd][jh][jw] = sum_{d,h,w} x[d,h,w] f_jd(d) f_jh(h) f_jw(w)
// where each row j of an axis is cos(2 pi u i / n) or sin(2 pi u i / n) for a distinct u.
// T is a separable REAL projection (three 1-D contractions); rows are shared by +k and -k,
// and even/odd folding (i <-> n-i) halves the multiplies of the two strided axes.
//
// The blob is position independent (offsets in 4-byte words) so the same bytes are used on
// the host (launch geometry) and on the device (tables).
#pragma once
#include <stdint.h>
#include <stddef.h>

namespace hno {

struct DhtAxis {
  int n;         // axis length
  int L;         // number of retained frequencies (length of the k list)
  int JC;        // cos rows (distinct |s|)
  int JS;        // sin rows (distinct |s| with a non-vanishing sine)
  int J;         // JC + JS
  int nh;        // n / 2
  int JCp;       // JC rounded up to a multiple of 4 (row length of fold_cos)
  int JSp;       // max(JS,1) rounded up to a multiple of 4
  int off_fcos;  // float [(nh+1)][JCp] : cos(2 pi u_j i / n), i = 0..nh   (zero padded)
  int off_fsin;  // float [(nh+1)][JSp] : sin(2 pi u_j i / n)
  int off_full;  // float [J][n]        : unfolded rows (cos rows first, then sin rows)
  int off_kdesc; // int   [L][4]        : {cos row, sin row or -1, sigma, k}
  int off_jdesc; // int   [J][4]        : {index in k list of +u or -1, index of -u or -1, is_sin, u}
  int off_fullT; // float [n4][J4]      : transposed unfolded rows, zero padded (n4, J4 = n, J rounded up to 4)
  int off_fullP; // float [J][n4]       : unfolded rows, zero padded to n4 columns
  int pad[1];
};

struct DhtPlanHeader {
  int magic;        // 'HNOP'
  int version;
  int total_words;  // size of the blob in 4-byte words
  int reserved;
  DhtAxis ax[3];    // 0 = D (slowest), 1 = H, 2 = W (contiguous)
};

constexpr int kDhtPlanMagic = 0x484E4F50;
constexpr int kDhtPlanVersion = 2;

size_t dht_plan_words(const int n[3], const int L[3]);
// Returns 0 on success; fills `blob` (must hold dht_plan_words()*4 bytes).
int dht_plan_fill(void* blob, size_t bytes, const int n[3], const int* const klist[3], const int L[3]);

}  // namespace hno
