// Host-side table builder for the truncated 3-D DHT (no device code, runs without a GPU).
#include "dht_plan.h"
#include "common.cuh"

#include <math.h>
#include <string.h>
#include <algorithm>
#include <vector>

namespace hno {

static inline int round4(int v) { return (v + 3) & ~3; }

size_t dht_plan_words(const int n[3], const int L[3]) {
  size_t words = sizeof(DhtPlanHeader) / 4;
  for (int a = 0; a < 3; ++a) {
    int nh = n[a] / 2;
    int jmax = std::min(L[a], nh + 1);
    int jp = round4(std::max(jmax, 1));
    words += (size_t)(nh + 1) * jp * 2;   // fold cos + fold sin
    words += (size_t)2 * jmax * n[a];     // full rows
    words += (size_t)2 * round4(n[a]) * round4(2 * jmax);  // transposed + padded copies of the full rows
    words += (size_t)4 * L[a];            // kdesc
    words += (size_t)4 * 2 * jmax;        // jdesc
    words += 16;                          // alignment slack
  }
  return words;
}

// cos / sin of 2*pi*(u*i mod n)/n with the product reduced in integers first, so the angle
// handed to libm is always in [0, 2*pi) and accurate to double precision.
static inline void cs(int u, int i, int n, double* c, double* s) {
  long r = ((long)u * (long)i) % n;
  // exact zeros where the analytic value is zero (keeps Nyquist / DC rows clean)
  if (r == 0) { *c = 1.0; *s = 0.0; return; }
  if (2 * r == n) { *c = -1.0; *s = 0.0; return; }
  if (4 * r == n) { *c = 0.0; *s = 1.0; return; }
  if (4 * r == 3L * n) { *c = 0.0; *s = -1.0; return; }
  double ang = 2.0 * M_PI * (double)r / (double)n;
  *c = cos(ang);
  *s = sin(ang);
}

int dht_plan_fill(void* blob, size_t bytes, const int n[3], const int* const klist[3], const int L[3]) {
  HNO_CHECK(blob != nullptr, "dht_plan_fill: null buffer");
  HNO_CHECK(bytes >= dht_plan_words(n, L) * 4, "dht_plan_fill: buffer too small (%zu < %zu)", bytes,
            dht_plan_words(n, L) * 4);
  memset(blob, 0, bytes);
  auto* hdr = reinterpret_cast<DhtPlanHeader*>(blob);
  float* fw = reinterpret_cast<float*>(blob);
  int* iw = reinterpret_cast<int*>(blob);
  hdr->magic = kDhtPlanMagic;
  hdr->version = kDhtPlanVersion;
  size_t cur = sizeof(DhtPlanHeader) / 4;
  auto align4 = [&]() { cur = (cur + 3) & ~(size_t)3; };

  for (int a = 0; a < 3; ++a) {
    const int na = n[a], La = L[a];
    HNO_CHECK(na >= 1 && La >= 1, "dht_plan_fill: axis %d has n=%d L=%d", a, na, La);
    DhtAxis& ax = hdr->ax[a];
    ax.n = na;
    ax.L = La;
    ax.nh = na / 2;
    // distinct |s|
    std::vector<int> us;
    std::vector<int> sig(La), uu(La);
    for (int t = 0; t < La; ++t) {
      int k = klist[a][t];
      HNO_CHECK(k >= 0 && k < na, "dht_plan_fill: axis %d frequency %d out of range [0,%d)", a, k, na);
      int s = (2 * k <= na) ? k : k - na;
      uu[t] = s < 0 ? -s : s;
      sig[t] = s > 0 ? 1 : (s < 0 ? -1 : 0);
      if (2 * uu[t] == na) sig[t] = 0;  // Nyquist: sine vanishes identically
      us.push_back(uu[t]);
    }
    for (int t = 0; t < La; ++t)
      for (int t2 = 0; t2 < t; ++t2)
        HNO_CHECK(klist[a][t] != klist[a][t2], "dht_plan_fill: axis %d frequency %d listed twice", a, klist[a][t]);
    std::sort(us.begin(), us.end());
    us.erase(std::unique(us.begin(), us.end()), us.end());
    std::vector<int> su;  // |s| values with a live sine row
    for (int u : us)
      if (u != 0 && 2 * u != na) su.push_back(u);
    ax.JC = (int)us.size();
    ax.JS = (int)su.size();
    ax.J = ax.JC + ax.JS;
    ax.JCp = round4(ax.JC);
    ax.JSp = round4(std::max(ax.JS, 1));

    align4();
    ax.off_fcos = (int)cur;
    for (int i = 0; i <= ax.nh; ++i)
      for (int j = 0; j < ax.JC; ++j) {
        double c, s;
        cs(us[j], i, na, &c, &s);
        fw[cur + (size_t)i * ax.JCp + j] = (float)c;
      }
    cur += (size_t)(ax.nh + 1) * ax.JCp;
    align4();
    ax.off_fsin = (int)cur;
    for (int i = 0; i <= ax.nh; ++i)
      for (int j = 0; j < ax.JS; ++j) {
        double c, s;
        cs(su[j], i, na, &c, &s);
        fw[cur + (size_t)i * ax.JSp + j] = (float)s;
      }
    cur += (size_t)(ax.nh + 1) * ax.JSp;
    align4();
    ax.off_full = (int)cur;
    for (int j = 0; j < ax.J; ++j)
      for (int i = 0; i < na; ++i) {
        double c, s;
        cs(j < ax.JC ? us[j] : su[j - ax.JC], i, na, &c, &s);
        fw[cur + (size_t)j * na + i] = (float)(j < ax.JC ? c : s);
      }
    cur += (size_t)ax.J * na;
    align4();
    {
      const int n4 = round4(na), J4 = round4(ax.J);
      const size_t full = (size_t)ax.off_full;
      ax.off_fullT = (int)cur;
      for (int i = 0; i < na; ++i)
        for (int j = 0; j < ax.J; ++j) fw[cur + (size_t)i * J4 + j] = fw[full + (size_t)j * na + i];
      cur += (size_t)n4 * J4;
      ax.off_fullP = (int)cur;
      for (int j = 0; j < ax.J; ++j)
        for (int i = 0; i < na; ++i) fw[cur + (size_t)j * n4 + i] = fw[full + (size_t)j * na + i];
      cur += (size_t)ax.J * n4;
    }
    align4();
    ax.off_kdesc = (int)cur;
    for (int t = 0; t < La; ++t) {
      int cj = (int)(std::lower_bound(us.begin(), us.end(), uu[t]) - us.begin());
      int sj = -1;
      if (sig[t] != 0) sj = ax.JC + (int)(std::lower_bound(su.begin(), su.end(), uu[t]) - su.begin());
      iw[cur + 4 * t + 0] = cj;
      iw[cur + 4 * t + 1] = sj;
      iw[cur + 4 * t + 2] = sig[t];
      iw[cur + 4 * t + 3] = klist[a][t];
    }
    cur += (size_t)4 * La;
    align4();
    ax.off_jdesc = (int)cur;
    for (int j = 0; j < ax.J; ++j) {
      int u = j < ax.JC ? us[j] : su[j - ax.JC];
      int kpos = -1, kneg = -1;
      for (int t = 0; t < La; ++t) {
        if (uu[t] != u) continue;
        int k = klist[a][t];
        if (k == u) kpos = t;        // s = +u (also DC and Nyquist)
        else kneg = t;               // s = -u
      }
      iw[cur + 4 * j + 0] = kpos;
      iw[cur + 4 * j + 1] = kneg;
      iw[cur + 4 * j + 2] = j < ax.JC ? 0 : 1;
      iw[cur + 4 * j + 3] = u;
    }
    cur += (size_t)4 * ax.J;
  }
  align4();
  hdr->total_words = (int)cur;
  HNO_CHECK(cur * 4 <= bytes, "dht_plan_fill: internal overflow");
  return 0;
}

}  // namespace hno
