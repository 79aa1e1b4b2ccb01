#include <cstdio>
#include <cstdlib>
#include <vector>
#include <string>
#include <limits>
#include <random>
#include "../../rapmap_b200/csrc/orphan_recovery.cuh"
using namespace rapmap_b200;
static int dp(const std::string& q, const std::string& t, int k, int& firstEnd) {
  int m = q.size(), n = t.size(); firstEnd = -1;
  if (m <= 0 || n <= 0) return -1;
  std::vector<int> prev(m + 1), cur(m + 1);
  for (int i = 0; i <= m; ++i) prev[i] = i;
  int best = std::numeric_limits<int>::max();
  for (int j = 1; j <= n; ++j) {
    cur[0] = 0;
    for (int i = 1; i <= m; ++i) {
      int d = prev[i-1] + (q[i-1] != t[j-1]);
      d = std::min(d, prev[i] + 1); d = std::min(d, cur[i-1] + 1);
      cur[i] = d;
    }
    if (cur[m] < best) { best = cur[m]; firstEnd = j - 1; }
    prev.swap(cur);
  }
  if (best > k) { firstEnd = -1; return -1; }
  return best;
}
int main() {
  std::mt19937_64 rng(12345);
  const char qa[] = "ACGTACGTACGTNacgtRU";
  const char ta[] = "ACGTACGTACGTACGTACGN$";
  long bad = 0, tot = 0, found = 0;
  for (int it = 0; it < 40000; ++it) {
    int m = 1 + rng() % (it % 50 == 0 ? 700 : 180);
    int n = 1 + rng() % 1000;
    bool rc = rng() & 1;
    std::string read(m, 'A'), t(n, 'A');
    for (auto& c : t) c = ta[rng() % (it % 7 == 0 ? 21 : 16)];
    // plant a noisy copy of the (possibly rc) query inside the window half of the time
    for (auto& c : read) c = qa[rng() % (it % 5 == 0 ? 19 : 12)];
    std::string q(m, 'A');
    for (int i = 0; i < m; ++i) q[i] = (char)orphanQueryChar((const uint8_t*)read.data(), m, rc, i);
    if ((rng() & 1) && n > m) {
      int at = rng() % (n - m + 1);
      for (int i = 0; i < m; ++i) t[at + i] = q[i];
      int edits = rng() % (m / 3 + 1);
      for (int e = 0; e < edits; ++e) { int p = at + rng() % m; t[p] = "ACGT"[rng() % 4]; }
      if (rng() % 3 == 0 && m > 4) { t.erase(at + rng() % m, 1); t.push_back('C'); }
    }
    int k = m / 4;
    int e1, e2;
    int d1 = dp(q, t, k, e1);
    int d2 = semiGlobalMyers((const uint8_t*)read.data(), m, rc, (const uint8_t*)t.data(), n, k, e2);
    ++tot; if (d1 >= 0) ++found;
    if (d1 != d2 || e1 != e2) { if (bad < 5) std::printf("MISMATCH m=%d n=%d rc=%d dp=(%d,%d) myers=(%d,%d)\n", m, n, (int)rc, d1, e1, d2, e2); ++bad; }
  }
  std::printf("cases %ld, with a hit %ld, mismatches %ld\n", tot, found, bad);
  return bad != 0;
}
