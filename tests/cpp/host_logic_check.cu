// tests/cpp/host_logic_check.cu -- CPU checks of the engine's host-side logic (no kernel launch, no GPU):
//   1. the fixed-point core shared by host and device code (csrc/common.cuh: convert / macc / tap_term) against the
//      assignment and `+=` of the AC-datatypes shim the oracle is built on, over all 8 quantisation x 4 overflow modes;
//   2. the coefficient packers (fir_q15_pack: effective direct-form taps of the folded architectures, reversed, two
//      byte planes; upfir_q15_pack: per-phase reversed composite taps, two or three byte planes) against an
//      independent expansion, and cic_history_len against DESIGN.md section 3.
// Built with nvcc (host code only) and linked with libb200dsp.so by tests/test_abi.py.
#include <ac_fixed.h>
#include <ac_int.h>

#include <cstdio>
#include <cstdlib>
#include <vector>

#include "kernels.h"

using namespace b2d;

static int g_bad = 0;
static long g_checks = 0;
#define EXPECT(cond, ...)                                        \
  do {                                                           \
    g_checks++;                                                  \
    if (!(cond)) {                                               \
      if (g_bad < 20) { std::printf("FAIL %s:%d: ", __FILE__, __LINE__); std::printf(__VA_ARGS__); std::printf("\n"); } \
      g_bad++;                                                   \
    }                                                            \
  } while (0)

static long long rnd64() { return ((long long)rand() << 42) ^ ((long long)rand() << 21) ^ rand(); }

// ---- 1. convert / macc / tap_term vs the shim
template <class SRC, class DST>
static void check_convert() {
  const Fmt fd{DST::width, DST::i_width, DST::sign ? 1 : 0, (int)DST::q_mode, (int)DST::o_mode};
  for (int i = 0; i < 4000; i++) {
    long long r = rnd64();
    if (i < 8) r = (i & 1) ? -(long long)(i >> 1) - 1 : (i >> 1);          // small values around zero
    else if (i < 16) r = (long long)((1ULL << (SRC::width - 1)) - 1ULL) - (i - 8);          // near the top of the range
    const SRC s = ac_shim::from_raw<SRC>(r);
    const DST d = s;                                                        // the shim's assignment: Q then O
    const long long want = ac_shim::to_raw(d);
    const long long got = convert((i128)ac_shim::to_raw(s), SRC::width - SRC::i_width, fd);
    EXPECT(want == got, "convert <%d,%d,%d> -> <%d,%d,%d,Q%d,O%d> raw %lld: shim %lld engine %lld", SRC::width, SRC::i_width,
           (int)SRC::sign, DST::width, DST::i_width, (int)DST::sign, (int)DST::q_mode, (int)DST::o_mode, ac_shim::to_raw(s), want, got);
  }
}

template <class A, class B, class ACC>
static void check_macc() {
  const Fmt fa{ACC::width, ACC::i_width, ACC::sign ? 1 : 0, (int)ACC::q_mode, (int)ACC::o_mode};
  const int Fp = (A::width - A::i_width) + (B::width - B::i_width);
  ACC acc = ac_shim::from_raw<ACC>(0);
  long long eacc = 0, emod = 0;
  const int s = Fp - fa.F();
  for (int i = 0; i < 3000; i++) {
    const A a = ac_shim::from_raw<A>(rnd64());
    const B b = ac_shim::from_raw<B>(rnd64());
    acc += a * b;                                                           // the reference's tap: exact product, += re-quantises
    const i128 p = (i128)ac_shim::to_raw(a) * (i128)ac_shim::to_raw(b);
    eacc = macc(eacc, fa, p, Fp);
    EXPECT(ac_shim::to_raw(acc) == eacc, "macc step %d: shim %lld engine %lld", i, ac_shim::to_raw(acc), eacc);
    if (fa.O == B2D_WRAP && (fa.Q == B2D_TRN || fa.Q == B2D_RND)) {         // the order-free form the fast kernels use
      emod += tap_term(p, s, fa.Q);
      EXPECT(wrap_bits(emod, fa.W, fa.S) == eacc, "tap_term step %d", i);
    }
  }
}

template <ac_q_mode Q, ac_o_mode O>
static void check_modes() {
  check_convert<ac_fixed<40, 8, true>, ac_fixed<16, 1, true, Q, O> >();
  check_convert<ac_fixed<33, 3, true>, ac_fixed<20, 5, false, Q, O> >();
  check_convert<ac_fixed<16, 1, true>, ac_fixed<40, 8, true, Q, O> >();
  check_convert<ac_fixed<64, 32, true>, ac_fixed<48, 32, true, Q, O> >();
  check_convert<ac_fixed<31, 9, false>, ac_fixed<12, 2, true, Q, O> >();
  check_macc<ac_fixed<16, 1, true>, ac_fixed<16, 1, true>, ac_fixed<40, 8, true, Q, O> >();
  check_macc<ac_fixed<28, 6, true>, ac_fixed<23, 7, true>, ac_fixed<64, 32, true, Q, O> >();   // prog bench: per-tap truncation
  check_macc<ac_fixed<16, 8, true>, ac_fixed<12, 2, false>, ac_fixed<24, 12, true, Q, O> >();
}

template <ac_q_mode Q>
static void check_q() {
  check_modes<Q, AC_WRAP>();
  check_modes<Q, AC_SAT>();
  check_modes<Q, AC_SAT_ZERO>();
  check_modes<Q, AC_SAT_SYM>();
}

// ---- 2. packers
static std::vector<int64_t> effective_taps(const std::vector<int64_t> &c, int N, int ft) {
  // ac_fir_load_coeffs.h:231-259 / ac_fir_reg_share.h:151-205: taps i and N-1-i share h[i] (negated for _ANTI), the centre
  // tap of an odd fold is unpaired
  std::vector<int64_t> e(N, 0);
  const bool even = ft == B2D_FOLD_EVEN || ft == B2D_FOLD_EVEN_ANTI, odd = ft == B2D_FOLD_ODD || ft == B2D_FOLD_ODD_ANTI;
  const int64_t sg = (ft == B2D_FOLD_EVEN_ANTI || ft == B2D_FOLD_ODD_ANTI) ? -1 : 1;
  for (int i = 0; i < N; i++) {
    if (!even && !odd) { e[i] = c[i]; continue; }
    const int half = even ? N / 2 : (N - 1) / 2 + 1;
    if (i < half) e[i] = c[i];
    else e[i] = sg * c[N - 1 - i];
    if (even && (N & 1) && i == N / 2) e[i] = 0;                             // FOLD_EVEN on an odd count never reads the middle tap
  }
  return e;
}

static void check_fir_q15_pack() {
  const Fmt fc{16, 1, 1, B2D_TRN, B2D_WRAP};
  const int fts[] = {B2D_SHIFT_REG, B2D_ROTATE_SHIFT, B2D_C_BUFF, B2D_TRANSPOSED, B2D_FOLD_EVEN, B2D_FOLD_ODD, B2D_FOLD_EVEN_ANTI, B2D_FOLD_ODD_ANTI};
  const int Ns[] = {1, 2, 16, 27, 29, 63, 256, 1024};
  for (int ft : fts)
    for (int N : Ns) {
      if ((ft == B2D_FOLD_EVEN || ft == B2D_FOLD_EVEN_ANTI) && (N & 1)) continue;
      if ((ft == B2D_FOLD_ODD || ft == B2D_FOLD_ODD_ANTI) && !(N & 1)) continue;
      std::vector<int64_t> c(N);
      const bool anti = ft == B2D_FOLD_EVEN_ANTI || ft == B2D_FOLD_ODD_ANTI;
      for (auto &v : c) v = anti ? (rand() % 32767) - 16383 : (rand() % 65536) - 32768;   // _ANTI: negated taps must stay 16-bit
      const int words = fir_q15_pk_words(N, ft);
      EXPECT(words * 2 >= N && words * 2 < N + 16 && (words * 2) % 16 == 0, "pk_words(%d) = %d", N, words);
      std::vector<uint32_t> pk(words, 0xDEADBEEF);
      fir_q15_pack(fc, c.data(), N, ft, pk.data(), words);
      const std::vector<int64_t> e = effective_taps(c, N, ft);
      for (int k = 0; k < 2 * words; k++) {
        const uint32_t w = pk[k / 2];
        const int el = k & 1;
        const int64_t lo = (w >> (8 * el)) & 0xFF, hi = (int8_t)((w >> (16 + 8 * el)) & 0xFF);
        const int64_t want = k < N ? e[N - 1 - k] : 0;                       // reversed: sample and tap index advance together
        EXPECT(hi * 256 + lo == want, "fir_q15_pack ft %d N %d k %d: %lld vs %lld", ft, N, k, (long long)(hi * 256 + lo), (long long)want);
      }
    }
}

// fir_q24: effective direct-form taps, reversed, two signed 16-bit taps per word (the DP2A 16-bit lanes)
static void check_fir_q24_pack() {
  const int fts[] = {B2D_SHIFT_REG, B2D_FOLD_EVEN, B2D_FOLD_ODD, B2D_FOLD_EVEN_ANTI, B2D_FOLD_ODD_ANTI};
  for (int ft : fts)
    for (int N : {1, 2, 15, 16, 17, 63, 64, 200}) {
      if ((ft == B2D_FOLD_EVEN || ft == B2D_FOLD_EVEN_ANTI) && (N & 1)) continue;
      if ((ft == B2D_FOLD_ODD || ft == B2D_FOLD_ODD_ANTI) && !(N & 1)) continue;
      const bool anti = ft == B2D_FOLD_EVEN_ANTI || ft == B2D_FOLD_ODD_ANTI;
      std::vector<int64_t> c(N);
      for (auto &v : c) v = anti ? (rand() % 32767) - 16383 : (rand() % 65536) - 32768;
      const int words = fir_q24_pk_words(N);
      EXPECT(2 * words >= N && (2 * words) % 16 == 0 && 2 * words < N + 16, "fir_q24_pk_words(%d) = %d", N, words);
      std::vector<uint32_t> pk(words, 0xDEADBEEF);
      fir_q24_pack(c.data(), N, ft, pk.data(), words);
      const std::vector<int64_t> e = effective_taps(c, N, ft);
      for (int k = 0; k < 2 * words; k++) {
        const int64_t got = (int16_t)((pk[k / 2] >> (16 * (k & 1))) & 0xFFFF);
        EXPECT(got == (k < N ? e[N - 1 - k] : 0), "fir_q24_pack ft %d N %d k %d", ft, N, k);
      }
    }
  const Fmt q15{16, 1, 1, B2D_TRN, B2D_WRAP}, s20{20, 5, 1, B2D_TRN, B2D_WRAP}, acc40{40, 8, 1, B2D_TRN, B2D_WRAP}, u24{24, 4, 0, B2D_TRN, B2D_WRAP};
  const Fmt s24{24, 4, 1, B2D_TRN, B2D_WRAP}, c17{17, 1, 1, B2D_TRN, B2D_WRAP}, sat{40, 8, 1, B2D_TRN, B2D_SAT};
  EXPECT(fir_q24_supported(s20, q15, acc40, acc40, 63, B2D_SHIFT_REG), "BASELINE configs[4] second stage");
  const Fmt acc48{48, 12, 1, B2D_TRN, B2D_WRAP};
  EXPECT(fir_q24_supported(s24, q15, acc48, acc48, 2048, B2D_TRANSPOSED), "24-bit signed samples");
  EXPECT(!fir_q24_supported(s24, q15, acc40, acc40, 2048, B2D_TRANSPOSED), "per-tap truncation (s > 0) is not an exact shift");
  EXPECT(!fir_q24_supported(q15, q15, acc40, acc40, 63, B2D_SHIFT_REG), "16-bit samples belong to fir_q15");
  EXPECT(!fir_q24_supported(u24, q15, acc40, acc40, 63, B2D_SHIFT_REG), "unsigned 24-bit samples need 25 bits");
  EXPECT(!fir_q24_supported(s20, c17, acc40, acc40, 63, B2D_SHIFT_REG), "17-bit taps do not fit the 16-bit lane");
  EXPECT(!fir_q24_supported(s20, q15, sat, acc40, 63, B2D_SHIFT_REG), "saturating accumulators are order-dependent");
  EXPECT(!fir_q24_supported(s20, q15, acc40, acc40, 64, B2D_FOLD_EVEN_ANTI), "negated 16-bit taps need 17 bits");
}

static void check_fir_wide_and_polydec_pack() {
  const int fts[] = {B2D_SHIFT_REG, B2D_FOLD_EVEN, B2D_FOLD_ODD, B2D_FOLD_EVEN_ANTI, B2D_FOLD_ODD_ANTI};
  for (int ft : fts)
    for (int N : {2, 8, 27, 63, 64}) {
      if ((ft == B2D_FOLD_EVEN || ft == B2D_FOLD_EVEN_ANTI) && (N & 1)) continue;
      if ((ft == B2D_FOLD_ODD || ft == B2D_FOLD_ODD_ANTI) && !(N & 1)) continue;
      std::vector<int64_t> c(N);
      for (auto &v : c) v = (rand() % 2000001) - 1000000;
      const int words = fir_wide_words(N);
      EXPECT(words >= N && words % 8 == 0 && words < N + 8, "fir_wide_words(%d) = %d", N, words);
      for (int mode = 0; mode < 3; mode++) {                                 // 0: folds expanded; 1, 2: the kernel folds itself, raw taps
        std::vector<int32_t> out(words, 0x5A5A5A5A);
        fir_wide_pack(c.data(), N, ft, mode, out.data(), words);
        const std::vector<int64_t> e = mode == 0 ? effective_taps(c, N, ft) : c;
        for (int i = 0; i < words; i++) EXPECT(out[i] == (i < N ? (int32_t)e[i] : 0), "fir_wide_pack ft %d N %d mode %d i %d", ft, N, mode, i);
      }
    }
  // polyphase decimator: phase-major planes, zero padded per phase (ac_poly_dec.h:110-126 reads coeffs[tp + NTAPS*df])
  for (int NT : {1, 7, 32})
    for (int DF : {2, 3, 8}) {
      std::vector<int64_t> c((size_t)NT * DF);
      for (auto &v : c) v = (rand() % 65536) - 32768;
      const int ntpad = polydec_words(NT);
      std::vector<int32_t> out((size_t)ntpad * DF, 0x5A5A5A5A);
      polydec_pack(c.data(), NT, DF, out.data());
      for (int r = 0; r < DF; r++)
        for (int tp = 0; tp < ntpad; tp++)
          EXPECT(out[(size_t)r * ntpad + tp] == (tp < NT ? (int32_t)c[tp + NT * r] : 0), "polydec_pack NT %d DF %d r %d tp %d", NT, DF, r, tp);
      const Fmt fc{16, 1, 1, B2D_TRN, B2D_WRAP};
      const int pkw = polydec_q15_words(NT, DF) / DF;
      std::vector<uint32_t> pk((size_t)pkw * DF, 0xDEADBEEF);
      polydec_q15_pack(fc, c.data(), NT, DF, pk.data());
      for (int r = 0; r < DF; r++)
        for (int k = 0; k < 2 * pkw; k++) {
          const uint32_t w = pk[(size_t)r * pkw + k / 2];
          const int el = k & 1;
          const int64_t got = (int64_t)((w >> (8 * el)) & 0xFF) + 256 * (int64_t)(int8_t)((w >> (16 + 8 * el)) & 0xFF);
          EXPECT(got == (k < NT ? c[(size_t)NT * r + NT - 1 - k] : 0), "polydec_q15_pack NT %d DF %d r %d k %d", NT, DF, r, k);
        }
    }
}

static void check_upfir_pack() {
  const int Rs[] = {2, 4, 8};
  const int Ts[] = {1, 5, 18, 72, 73, 200};
  for (int planes = 2; planes <= 3; planes++)
    for (int R : Rs)
      for (int T : Ts) {
        const int bits = planes == 3 ? 23 : 15;
        std::vector<int64_t> c(T);
        for (auto &v : c) v = (rnd64() & ((1LL << (bits + 1)) - 1)) - (1LL << bits);
        EXPECT(upfir_q15_geometry(R, T, bits + 1), "geometry R %d T %d", R, T);
        EXPECT(upfir_q15_planes(bits + 1) == planes, "planes for %d bits", bits + 1);
        const int words = upfir_q15_words(R, T, planes), wpp = planes == 3 ? 2 : 1;
        const int Tc = (T + R - 1) / R, TP = (Tc + 1) / 2;
        EXPECT(words == R * TP * wpp, "words");
        std::vector<uint32_t> pk(words, 0xDEADBEEF);
        upfir_q15_pack(c.data(), T, R, planes, pk.data());
        for (int ph = 0; ph < R; ph++)
          for (int k = 0; k < 2 * TP; k++) {
            const uint32_t wa = pk[(ph * TP + k / 2) * wpp], wb = planes == 3 ? pk[(ph * TP + k / 2) * wpp + 1] : 0;
            const int el = k & 1;
            const int64_t b0 = (wa >> (8 * el)) & 0xFF, b1 = (wa >> (16 + 8 * el)) & 0xFF, b2 = (int8_t)((wb >> (8 * el)) & 0xFF);
            const int64_t got = planes == 3 ? b0 + 256 * b1 + 65536 * b2 : b0 + 256 * (int64_t)(int8_t)b1;
            const int m = 2 * TP - 1 - k, idx = ph + R * m;                   // per phase reversed, zero padded on the oldest sample
            const int64_t want = (m < Tc && idx < T) ? c[idx] : 0;
            EXPECT(got == want, "upfir_q15_pack planes %d R %d T %d ph %d k %d: %lld vs %lld", planes, R, T, ph, k, (long long)got, (long long)want);
          }
      }
  EXPECT(!upfir_q15_geometry(3, 63, 24) && !upfir_q15_geometry(4, 63, 25) && !upfir_q15_geometry(4, 4 * 128 + 1, 24), "geometry limits");
}

// The 64-bit instantiation of the fixed-point primitives equals the 128-bit one wherever fits_i64 admits it.
static void check_i64_primitives() {
  const int Qs[] = {B2D_TRN, B2D_RND, B2D_TRN_ZERO, B2D_RND_ZERO, B2D_RND_INF, B2D_RND_MIN_INF, B2D_RND_CONV, B2D_RND_CONV_ODD};
  const int Os[] = {B2D_WRAP, B2D_SAT, B2D_SAT_ZERO, B2D_SAT_SYM};
  int admitted = 0;
  for (int t = 0; t < 4000; t++) {
    Fmt acc{8 + rand() % 50, 0, rand() % 4 ? 1 : 0, Qs[rand() % 8], Os[rand() % 4]};
    acc.I = acc.W - (rand() % 40);
    Fmt out{4 + rand() % 60, 0, rand() % 2, Qs[rand() % 8], Os[rand() % 4]};
    out.I = out.W - (acc.F() - (rand() % 12) + 4);
    const int Wp = 6 + rand() % 50, Fp = acc.F() + (rand() % 24) - 8;
    if (!fits_i64(acc, out, Wp, Fp)) continue;
    admitted++;
    for (int k = 0; k < 24; k++) {
      const long long a0 = wrap_bits(((long long)rand() << 32) ^ ((long long)rand() << 11) ^ rand(), acc.W, acc.S);
      long long p = wrap_bits(((long long)rand() << 32) ^ ((long long)rand() << 9) ^ rand(), Wp, 1);
      if (k == 0) p = -(1LL << (Wp - 1));
      if (k == 1) p = (1LL << (Wp - 1)) - 1;
      const int64_t r128 = macc_t<i128>(a0, acc, (i128)p, Fp), r64 = macc_t<int64_t>(a0, acc, (int64_t)p, Fp);
      EXPECT(r128 == r64, "macc_t<int64_t> W %d I %d S %d Q %d O %d Wp %d Fp %d", acc.W, acc.I, acc.S, acc.Q, acc.O, Wp, Fp);
      const int64_t o128 = convert_t<i128>((i128)r128, acc.F(), out), o64 = convert_t<int64_t>(r128, acc.F(), out);
      EXPECT(o128 == o64, "convert_t<int64_t> acc F %d -> out W %d I %d S %d Q %d O %d", acc.F(), out.W, out.I, out.S, out.Q, out.O);
    }
  }
  EXPECT(admitted > 500, "fits_i64 admitted only %d of 4000 draws", admitted);
}

static void check_support_tables() {
  const Fmt q15{16, 1, 1, B2D_TRN, B2D_WRAP}, acc40{40, 8, 1, B2D_TRN, B2D_WRAP};
  Fmt sat = acc40; sat.O = B2D_SAT;
  Fmt conv = acc40; conv.Q = B2D_RND_CONV;
  Fmt acc24{24, 8, 1, B2D_TRN, B2D_WRAP};                                     // F_acc = 16 < 30: per-tap truncation, not the q15 path
  EXPECT(fir_q15_supported(q15, q15, acc40, acc40, 256, B2D_SHIFT_REG), "cfg 2 must take fir_q15");
  EXPECT(fir_q15_supported(q15, q15, acc40, acc40, 1024, B2D_TRANSPOSED), "cfg 4 must take fir_q15");
  EXPECT(!fir_q15_supported(q15, q15, sat, sat, 256, B2D_SHIFT_REG), "saturating accumulators are order-dependent");
  EXPECT(!fir_q15_supported(q15, q15, conv, conv, 256, B2D_SHIFT_REG), "parity-dependent rounding is order-dependent");
  EXPECT(!fir_q15_supported(q15, q15, acc24, acc24, 256, B2D_SHIFT_REG), "per-tap truncation is not an exact shift");
  EXPECT(!fir_q15_supported(q15, q15, acc40, acc40, 4096, B2D_SHIFT_REG), "tap limit");
  EXPECT(!fir_q15_supported(q15, q15, acc40, acc40, 256, B2D_FOLD_EVEN_ANTI), "negated 16-bit taps need 17 bits");
  EXPECT(fir_q15_supported(q15, q15, acc40, acc40, 255, B2D_FOLD_ODD), "exact ACC_TYPE pre-add");
  EXPECT(cic_history_len(0, 8, 1, 4) == 4 * 1 * 8 + 4 - 1, "decimator history N*M*R + N - 1");
  EXPECT(cic_history_len(1, 4, 1, 3) == 3 * 1 + 3 + 2, "interpolator history N*M + N + 2");
}

int main() {
  srand(20260101);
  check_q<AC_TRN>(); check_q<AC_RND>(); check_q<AC_TRN_ZERO>(); check_q<AC_RND_ZERO>();
  check_q<AC_RND_INF>(); check_q<AC_RND_MIN_INF>(); check_q<AC_RND_CONV>(); check_q<AC_RND_CONV_ODD>();
  check_fir_q15_pack();
  check_upfir_pack();
  check_fir_wide_and_polydec_pack();
  check_fir_q24_pack();
  check_i64_primitives();
  check_support_tables();
  std::printf("checks=%ld bad=%d\n", g_checks, g_bad);
  return g_bad != 0;
}
