// Host build of the pair DP (rapmap_b200/csrc/ksw_pair.cuh is host-callable) for the CPU tests.
//  * ksw_pair_host_score(): C entry point for ctypes (tests compare it with the oracle's restatement of ksw_extz2_sse41).
//  * with -DWITH_REF and the reference's src/ksw2pp/ksw2_extz2_sse.c compiled alongside (from where it lies under
//    /root/reference, never copied): main() compares random and adversarial job pairs with the reference's own SSE function.
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <random>
#include <string>
#include <vector>

#include "../../rapmap_b200/csrc/ksw_pair.cuh"

using namespace rapmap_b200::kswpair;

static uint8_t nt4(uint8_t c) {  // seq_nt4_table_loc (reference src/ksw2pp/KSW2Aligner.cpp:61-72)
  if (c < 4) return c;
  switch (c | 0x20) {
    case 'a': return 0;
    case 'c': return 1;
    case 'g': return 2;
    case 't': return 3;
    default: return 4;
  }
}

// strip of one pair, word stride 1 (layout: ksw_pair.cuh)
static void fillStrip(std::vector<uint32_t>& W, const uint8_t* q0, const uint8_t* q1, int qlen, const uint8_t* t0, const uint8_t* t1, int tlen) {
  const int TL = stripTL(tlen), cells = stripCells(qlen, tlen);
  W.assign(static_cast<size_t>(cells + 1) / 2 + 1, 0u);
  uint8_t* B = reinterpret_cast<uint8_t*>(W.data());
  for (int t = 0; t < tlen; ++t) { B[2 * t] = nt4(t0[t]); B[2 * t + 1] = nt4(t1[t]); }
  for (int i = 0; i < qlen; ++i) { B[2 * (TL + i)] = nt4(q0[qlen - 1 - i]); B[2 * (TL + i) + 1] = nt4(q1[qlen - 1 - i]); }
}

// returns: 1 scored, 0 pair must go to the byte-exact kernel, -1 geometry / scores outside the pair kernel
extern "C" int ksw_pair_host_score(const uint8_t* q0, const uint8_t* q1, int qlen, const uint8_t* t0, const uint8_t* t1, int tlen, int mat0, int mat1, int q, int e,
                                   int w, int maxCells, int32_t* out) {
  const Consts C = makeConsts(mat0, mat1, 0, q, e, w);
  if (!geomOk(C, qlen, tlen, maxCells)) return -1;
  std::vector<uint32_t> W;
  fillStrip(W, q0, q1, qlen, t0, t1, tlen);
  return pairDP<1>(W.data(), qlen, tlen, C, out[0], out[1]) ? 1 : 0;
}

#ifdef WITH_REF
#include "ksw2pp/ksw2.h"
extern "C" void ksw_extz2_sse(void* km, int qlen, const uint8_t* query, int tlen, const uint8_t* target, int8_t m, const int8_t* mat, int8_t q, int8_t e, int w, int zdrop,
                              int end_bonus, int flag, ksw_extz_t* ez);

static int32_t refScore(const std::string& qs, const std::string& ts, int mat0, int mat1, int q, int e, int w) {
  int8_t mat[25];
  for (int i = 0; i < 4; ++i) {
    for (int j = 0; j < 4; ++j) mat[i * 5 + j] = static_cast<int8_t>(i == j ? mat0 : mat1);
    mat[i * 5 + 4] = 0;
  }
  for (int j = 0; j < 5; ++j) mat[20 + j] = 0;
  std::vector<uint8_t> qc(qs.size()), tc(ts.size());
  for (size_t i = 0; i < qs.size(); ++i) qc[i] = nt4(qs[i]);
  for (size_t i = 0; i < ts.size(); ++i) tc[i] = nt4(ts[i]);
  ksw_extz_t ez;
  std::memset(&ez, 0, sizeof(ez));
  ez.max_q = ez.max_t = ez.mqe_t = ez.mte_q = -1;
  ez.max = 0; ez.mqe = ez.mte = KSW_NEG_INF;
  ksw_extz2_sse(nullptr, static_cast<int>(qc.size()), qc.data(), static_cast<int>(tc.size()), tc.data(), 5, mat, static_cast<int8_t>(q), static_cast<int8_t>(e), w, -1, 10,
                KSW_EZ_SCORE_ONLY, &ez);
  return ez.mqe > ez.mte ? ez.mqe : ez.mte;
}

int main(int argc, char** argv) {
  const long iters = argc > 1 ? std::atol(argv[1]) : 200000;
  std::mt19937_64 rng(20261017);
  auto rnd = [&](int n) { return static_cast<int>(rng() % static_cast<uint64_t>(n)); };
  const char acgt[] = "ACGT";
  long scored = 0, fallback = 0, skipped = 0, bad = 0;
  for (long it = 0; it < iters; ++it) {
    int mat0 = 2, mat1 = -4, q = 5, e = 3, w = 15;
    if (it % 4 == 1) { w = 1 + rnd(15); }
    if (it % 4 == 2) { mat0 = 1 + rnd(6); mat1 = -rnd(9); q = rnd(13); e = rnd(7); w = 1 + rnd(15); }
    if (it % 64 == 3) { mat0 = 1 + rnd(40); mat1 = -rnd(40); q = rnd(30); e = rnd(20); }   // mostly refused (M + q > 96) or exotic
    int qlen = 20 + rnd(140);
    if (it % 16 == 5) qlen = 1 + rnd(40);
    int tlen = qlen + 20;
    if (it % 3 == 0) tlen = qlen - 12 + rnd(45);
    if (it % 5 == 4) { qlen = 90 + rnd(60); tlen = qlen - 10 - rnd(30); }   // window cut by the transcript end
    if (tlen < 64) tlen = 64 + rnd(8);
    std::string t[2], qs[2];
    for (int j = 0; j < 2; ++j) {
      t[j].resize(tlen);
      for (auto& c : t[j]) c = acgt[rnd(4)];
      if (it % 11 == 0) for (int k = 0; k < 3; ++k) t[j][rnd(tlen)] = 'N';
      if (it % 13 == 0) { const int p = rnd(tlen), l = rnd(30); for (int k = p; k < p + l && k < tlen; ++k) t[j][k] = 'A'; }
      // query: a noisy copy of the window start (substitutions, insertions, deletions), or unrelated
      std::string s;
      const int mode = rnd(8);
      const int sub = mode == 0 ? 0 : (mode < 5 ? 30 : 6), indel = mode < 3 ? 0 : (mode < 6 ? 40 : 8);
      int shift = (it % 7 == 0) ? rnd(25) : 0;   // alignment starts off the main diagonal
      for (int p = shift; static_cast<int>(s.size()) < qlen;) {
        if (mode == 7 || p >= tlen) { s.push_back(acgt[rnd(4)]); ++p; continue; }
        if (indel && rnd(indel) == 0) { if (rnd(2)) { s.push_back(acgt[rnd(4)]); } else { ++p; } continue; }
        char c = t[j][p++];
        if (sub && rnd(sub) == 0) c = acgt[rnd(4)];
        s.push_back(c);
      }
      if (it % 17 == 0) s[rnd(qlen)] = 'N';
      if (it % 19 == 0) s[rnd(qlen)] = 'r';
      qs[j] = s;
    }
    int32_t out[2] = {0, 0};
    const int rc = ksw_pair_host_score(reinterpret_cast<const uint8_t*>(qs[0].data()), reinterpret_cast<const uint8_t*>(qs[1].data()), qlen,
                                       reinterpret_cast<const uint8_t*>(t[0].data()), reinterpret_cast<const uint8_t*>(t[1].data()), tlen, mat0, mat1, q, e, w, 384, out);
    if (rc < 0) { ++skipped; continue; }
    if (rc == 0) { ++fallback; continue; }
    ++scored;
    for (int j = 0; j < 2; ++j) {
      const int32_t ref = refScore(qs[j], t[j], mat0, mat1, q, e, w);
      if (ref != out[j]) {
        if (++bad <= 10)
          std::printf("MISMATCH it=%ld job=%d qlen=%d tlen=%d mat=%d/%d q=%d e=%d w=%d: pair %d ref %d\n", it, j, qlen, tlen, mat0, mat1, q, e, w, out[j], ref);
      }
    }
  }
  std::printf("pairs scored %ld, handed to the exact kernel %ld, outside the pair kernel %ld, mismatches %ld\n", scored, fallback, skipped, bad);
  return bad ? 1 : 0;
}
#endif
