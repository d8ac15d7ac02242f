// tests/cpp/fir_ovs_check.cu -- CPU check of the overlap-save FIR path (no kernel launch, no GPU).
// The thread phases of csrc/fir_ovs.cuh are run as loops over the 256 thread ids of a CTA, block after block, with the
// tables and spectra the host runtime prepares (fir_ovs_tables / fir_ovs_spectrum from libb200dsp.so), and the rounded
// results are compared with a direct 64-bit integer convolution: indexing of the three radix-16 passes, the history /
// stream-edge handling, the real-pair packing, and the distance of every FP64 result from an integer against the a-priori
// error bound.  Built with nvcc (host code only) by tests/test_fir_ovs.py.
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <string>
#include <vector>

#include "fir_ovs.cuh"
#include "kernels.h"

using namespace b2d;
using namespace b2d::ovs;

static int g_bad = 0;

static uint64_t g_state = 0x9E3779B97F4A7C15ULL;
static uint32_t rnd() { g_state ^= g_state << 13; g_state ^= g_state >> 7; g_state ^= g_state << 17; return (uint32_t)(g_state >> 16); }

// mode 0: uniform full range; 1: every sample / tap at the largest magnitude with signs that line up (largest sums);
// 2: pure tone at full scale (all energy in one bin)
static void fill(std::vector<int> &v, int lo, int hi, int mode, int phase) {
  for (size_t i = 0; i < v.size(); i++) {
    if (mode == 0) v[i] = lo + (int)(rnd() % (uint32_t)(hi - lo + 1));
    else if (mode == 1) v[i] = ((i + phase) & 1) ? lo : (lo < 0 ? lo : hi);
    else v[i] = (int)lrint((lo < 0 ? -(double)lo - 1 : (double)hi) * std::cos(0.7853981633974483 * (double)(i + phase)));
  }
}

// np = 2: one interleaved IQ pair; np = 1: C planar real channels
static void run_case(const char *name, int np, uint32_t C, int n_taps, size_t n, int xs, int mode, bool same_taps) {
  const int T = n_taps - 1;
  const int lo = xs ? -32768 : 0, hi = xs ? 32767 : 65535;
  std::vector<std::vector<int>> x(C, std::vector<int>(n)), tl(C, std::vector<int>(T)), h(C, std::vector<int>(n_taps));
  for (uint32_t c = 0; c < C; c++) {
    fill(x[c], lo, hi, mode, (int)c);
    fill(tl[c], lo, hi, mode, (int)c + 3);
    if (c && same_taps) h[c] = h[0];
    else fill(h[c], -32768, 32767, mode == 2 ? 1 : mode, 0);
  }
  // containers as the kernel sees them
  std::vector<uint16_t> xin(C * n), tail((size_t)C * T + 1);
  for (uint32_t c = 0; c < C; c++) {
    for (size_t i = 0; i < n; i++) xin[np == 2 ? i * 2 + c : (size_t)c * n + i] = (uint16_t)x[c][i];
    for (int i = 0; i < T; i++) tail[(size_t)c * T + i] = (uint16_t)tl[c][i];
  }
  std::vector<double2> tw1(6 * 256), tw2(6 * 16), hs((size_t)C * kN);
  fir_ovs_tables(tw1.data(), tw2.data());
  double l1max = 0;
  for (uint32_t c = 0; c < C; c++) {
    std::vector<int64_t> e(h[c].begin(), h[c].end());
    fir_ovs_spectrum(e.data(), n_taps, hs.data() + (size_t)c * kN);
    double l1 = 0;
    for (int v : h[c]) l1 += std::fabs((double)v);
    l1max = std::fmax(l1max, l1);
  }
  Fmt fin{16, 1, xs, B2D_TRN, B2D_WRAP};
  const double bound = fir_ovs_error_bound(fin, l1max);

  Args a;
  a.x = xin.data(); a.y = nullptr; a.tail = tail.data(); a.tw = nullptr; a.hs = hs.data();
  a.n = n; a.T = T; a.D = fir_ovs_discard(n_taps); a.L = kN - a.D; a.C = C; a.xs = xs; a.lsh = 0;
  a.resid = nullptr;
  std::vector<std::vector<long long>> got(C, std::vector<long long>(n, -1));
  std::vector<double2> sm(kSmElems);
  double resid = 0;
  size_t n_interior = 0;
  const size_t blocks = (n + a.L - 1) / a.L;
  const size_t ctas = np == 2 ? blocks : (blocks + 1) / 2;
  for (uint32_t c0 = 0; c0 < (np == 2 ? 1u : C); c0++)
    for (size_t blk = 0; blk < ctas; blk++) {
      const bool in = np == 2 ? block_interior<2>(a, (long long)blk) : block_interior<1>(a, (long long)blk);
      n_interior += in;
      for (int t = 0; t < kThreads; t++) {
        if (np == 2) { if (in) phase_a<2, true>(a, tw1.data(), c0, (long long)blk, t, sm.data()); else phase_a<2, false>(a, tw1.data(), c0, (long long)blk, t, sm.data()); }
        else { if (in) phase_a<1, true>(a, tw1.data(), c0, (long long)blk, t, sm.data()); else phase_a<1, false>(a, tw1.data(), c0, (long long)blk, t, sm.data()); }
      }
      for (int t = 0; t < kThreads; t++) phase_b(tw2.data(), t, sm.data());
      for (int t = 0; t < kThreads; t++) phase_c(hs.data() + (size_t)(np == 2 ? 0 : c0) * kN, t, sm.data());
      for (int t = 0; t < kThreads; t++) phase_d(tw2.data(), t, sm.data());
      for (int t = 0; t < kThreads; t++) {
        double2 v[16];
        phase_e(tw1.data(), t, sm.data(), v);
        for (int k = a.D >> 8; k < 16; k++) {
          for (int e = 0; e < 2; e++) {
            const double d = e ? v[k].y : v[k].x;
            const long long g = (np == 2 ? (long long)blk : 2 * (long long)blk + e) * a.L - a.D + t + 256 * k;
            if (g < 0 || (size_t)g >= n) continue;
            resid = std::fmax(resid, std::fabs(d - std::rint(d)));
            got[np == 2 ? e : c0][g] = llrint(d);
          }
        }
      }
    }
  // direct convolution
  size_t bad = 0;
  for (uint32_t c = 0; c < C; c++)
    for (size_t i = 0; i < n; i++) {
      long long s = 0;
      for (int k = 0; k < n_taps; k++) {
        const long long idx = (long long)i - k;
        const int xv = idx >= 0 ? x[c][idx] : (T + idx >= 0 ? tl[c][T + idx] : 0);
        s += (long long)xv * h[np == 2 ? 0 : c][k];
      }
      if (s != got[c][i]) { if (bad < 5) std::printf("  mismatch c %u i %zu: want %lld got %lld\n", c, i, s, got[c][i]); bad++; }
    }
  std::printf("%-28s np %d C %u taps %4d n %6zu (%zu interior blocks): mismatches %zu, max |v - rint(v)| %.3e, a-priori bound %.3e%s\n", name, np, C, n_taps, n, n_interior, bad,
              resid, bound, bound < 0.5 ? "" : "  (predicate would refuse)");
  if (bad || resid > bound || resid > 0.05) g_bad++;
}

// `dump FILE`: the twiddle tables and the spectrum of a fixed tap set as raw doubles, for tests/test_fir_ovs.py to compare with
// a multi-precision evaluation (the error analysis assumes correctly rounded tables and |dH_k| <= 1.1 u ||h||_1 / 4096).
static int dump(const char *path) {
  std::vector<double2> tw1(6 * 256), tw2(6 * 16), hs(kN);
  fir_ovs_tables(tw1.data(), tw2.data());
  std::vector<int64_t> h(1000);
  for (size_t i = 0; i < h.size(); i++) h[i] = (int64_t)(rnd() % 65536) - 32768;
  fir_ovs_spectrum(h.data(), (int)h.size(), hs.data());
  FILE *f = std::fopen(path, "wb");
  if (!f) return 2;
  const int64_t n = (int64_t)h.size();
  std::fwrite(&n, sizeof(n), 1, f);
  std::fwrite(h.data(), sizeof(int64_t), h.size(), f);
  std::fwrite(tw1.data(), sizeof(double2), tw1.size(), f);
  std::fwrite(tw2.data(), sizeof(double2), tw2.size(), f);
  std::fwrite(hs.data(), sizeof(double2), hs.size(), f);
  std::fclose(f);
  return 0;
}

// `stress N`: N single blocks of an IQ pair with every sample and every one of 256 taps at +-full scale, signs drawn at random
// (the largest ||x||_2 ||h||_1 the format allows): the largest distance of a result from an integer, against the bound 0.21.
static int stress(int blocks) {
  const int n_taps = 256, T = n_taps - 1;
  std::vector<double2> tw1(6 * 256), tw2(6 * 16), hs(kN), sm(kSmElems);
  fir_ovs_tables(tw1.data(), tw2.data());
  double worst = 0;
  size_t wrong = 0;
  for (int b = 0; b < blocks; b++) {
    std::vector<int64_t> h(n_taps);
    for (auto &v : h) v = (rnd() & 1) ? -32768 : 32767;
    fir_ovs_spectrum(h.data(), n_taps, hs.data());
    const size_t n = 3840;
    std::vector<uint16_t> xin(2 * n), tail(2 * T);
    for (auto &v : xin) v = (uint16_t)((rnd() & 1) ? -32768 : 32767);
    for (auto &v : tail) v = (uint16_t)((rnd() & 1) ? -32768 : 32767);
    Args a;
    a.x = xin.data(); a.y = nullptr; a.tail = tail.data(); a.tw = nullptr; a.hs = hs.data();
    a.n = n; a.T = T; a.D = 256; a.L = kN - 256; a.C = 2; a.xs = 1; a.lsh = 0; a.resid = nullptr;
    for (int t = 0; t < kThreads; t++) phase_a<2, false>(a, tw1.data(), 0, 0, t, sm.data());
    for (int t = 0; t < kThreads; t++) phase_b(tw2.data(), t, sm.data());
    for (int t = 0; t < kThreads; t++) phase_c(hs.data(), t, sm.data());
    for (int t = 0; t < kThreads; t++) phase_d(tw2.data(), t, sm.data());
    for (int t = 0; t < kThreads; t++) {
      double2 v[16];
      phase_e(tw1.data(), t, sm.data(), v);
      for (int k = 1; k < 16; k++)
        for (int e = 0; e < 2; e++) {
          const double d = e ? v[k].y : v[k].x;
          worst = std::fmax(worst, std::fabs(d - std::rint(d)));
          const long long g = t + 256 * k - 256;                // output index; exact sum over the history and this block
          if (b % 16 == 0 && (t + k) % 37 == 0) {
            long long sum = 0;
            for (int i = 0; i < n_taps; i++) {
              const long long idx = g - i;
              const int xv = (int)(int16_t)(idx >= 0 ? xin[2 * idx + e] : tail[(size_t)e * T + (size_t)(T + idx)]);
              sum += (long long)xv * h[i];
            }
            if (sum != llrint(d)) wrong++;
          }
        }
    }
  }
  std::printf("stress: %d blocks of full-scale +-1 patterns, largest |v - rint(v)| %.3e, spot-checked sums wrong: %zu\n", blocks, worst, wrong);
  return (worst > 1e-3 || wrong) ? 1 : 0;
}

int main(int argc, char **argv) {
  if (argc == 3 && std::string(argv[1]) == "dump") return dump(argv[2]);
  if (argc == 3 && std::string(argv[1]) == "stress") return stress(std::atoi(argv[2]));
  run_case("iq256 random", 2, 2, 256, 4 * 3840 + 77, 1, 0, true);
  run_case("iq256 unsigned", 2, 2, 256, 3 * 3840 + 1, 0, 0, true);
  run_case("iq256 largest sums", 2, 2, 256, 3840 + 5, 1, 1, true);
  run_case("iq256 tone", 2, 2, 256, 3840 + 300, 1, 2, true);
  run_case("iq97 unsigned short call", 2, 2, 97, 300, 0, 0, true);
  run_case("real1024 x3 random", 1, 3, 1024, 7 * 3072 + 11, 1, 0, false);
  run_case("real300 unsigned", 1, 2, 300, 6 * 3584 + 5, 0, 0, false);
  run_case("real1024 largest sums", 1, 1, 1024, 2 * 3072, 1, 1, false);
  run_case("real2049", 1, 1, 2049, 2048 + 100, 1, 0, false);
  run_case("real128 one block", 1, 2, 128, 1000, 1, 0, false);
  std::printf("bad=%d\n", g_bad);
  return g_bad != 0;
}
