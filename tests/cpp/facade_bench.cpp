// tests/cpp/facade_bench.cpp -- drives the B200 engine through the header facade (include/b200dsp/ac_dsp/*.h) the way
// the reference's own benches drive the reference classes (tests/rtest_ac_*.cpp of hlslibs/ac_dsp): ac_fixed values
// queued on ac_channel FIFOs, run(), outputs drained from the output channel.  The configurations are those benches'
// (formats, taps, ftype, R/M/N) plus the BASELINE.json ones; stimulus and expected outputs come from the committed
// fixtures in tests/golden/ (raw integers as text, written and compared by tests/test_facade.py).
//
//   facade_bench <case> <in.txt> <coef.txt|-> <out.txt> [chunk]
//
// `chunk` > 0 feeds the input in several run() calls of that many samples (state must carry across calls).
#include <ac_fixed.h>
#include <ac_channel.h>
#include <ac_dsp/ac_fir_const_coeffs.h>
#include <ac_dsp/ac_fir_load_coeffs.h>
#include <ac_dsp/ac_fir_prog_coeffs.h>
#include <ac_dsp/ac_cic_dec_full.h>
#include <ac_dsp/ac_cic_intr_full.h>
#include <ac_dsp/ac_fir_reg_share.h>
#include <ac_dsp/ac_poly_dec.h>
#include <ac_dsp/ac_intg_dump.h>
#include <ac_dsp/ac_mv_avg.h>

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <iostream>
#include <string>
#include <vector>

static std::vector<long long> read_ints(const char *path) {
  std::vector<long long> v;
  std::ifstream f(path);
  long long x;
  while (f >> x) v.push_back(x);
  return v;
}

template <class T>
static T from_raw(long long r) { return b200dsp::fixed_traits<T>::from_raw(r); }
template <class T>
static long long to_raw(const T &t) { return b200dsp::fixed_traits<T>::to_raw(t); }

template <class T>
static void drain_to(ac_channel<T> &ch, std::vector<long long> &out) {
  while (ch.available(1)) out.push_back(to_raw(ch.read()));
}

// the wrapper idiom of the reference's constant-coefficient bench: the taps are a member of the DERIVED class
template <class IN, class OUT, class COEFF, class ACC, unsigned N, FTYPE ft>
class const_wrapper : public ac_fir_const_coeffs<IN, OUT, COEFF, ACC, N, ft> {
public:
  COEFF coeffs[N];
  explicit const_wrapper(const std::vector<long long> &c) : ac_fir_const_coeffs<IN, OUT, COEFF, ACC, N, ft>(coeffs) {
    for (unsigned i = 0; i < N; i++) coeffs[i] = from_raw<COEFF>(c[i]);
  }
};

template <class IN, class OUT, class COEFF, class ACC, unsigned N, FTYPE ft>
static int run_const(const std::vector<long long> &x, const std::vector<long long> &c, size_t chunk, std::vector<long long> &y) {
  if (c.size() != N) return 2;
  const_wrapper<IN, OUT, COEFF, ACC, N, ft> filter(c);
  ac_channel<IN> in;
  ac_channel<OUT> out;
  for (size_t i = 0; i < x.size(); i++) {
    in.write(from_raw<IN>(x[i]));
    if (chunk && (i + 1) % chunk == 0) filter.run(in, out);
  }
  filter.run(in, out);
  drain_to(out, y);
  return 0;
}

template <class IN, class OUT, class COEFF, class ACC, unsigned N, FTYPE ft>
static int run_load(const std::vector<long long> &x, const std::vector<long long> &c, size_t chunk, std::vector<long long> &y) {
  if (c.size() != N) return 2;
  ac_fir_load_coeffs<IN, OUT, COEFF, ACC, N, ft> filter;
  ac_channel<IN> in;
  ac_channel<OUT> out;
  ac_channel<COEFF> coeffs_ch;
  ac_channel<bool> ld;
  // an under-filled load request must be dropped silently (reference ac_fir_load_coeffs.h:328)
  coeffs_ch.write(from_raw<COEFF>(c[0]));
  ld.write(true);
  filter.run(in, coeffs_ch, out, ld);
  if (coeffs_ch.debug_size() != 1) return 3;
  for (unsigned i = 1; i < N; i++) coeffs_ch.write(from_raw<COEFF>(c[i]));
  ld.write(true);
  filter.run(in, coeffs_ch, out, ld);   // load phase, no samples queued
  ld.write(false);
  for (size_t i = 0; i < x.size(); i++) {
    in.write(from_raw<IN>(x[i]));
    if (chunk && (i + 1) % chunk == 0) filter.run(in, coeffs_ch, out, ld);
  }
  filter.run(in, coeffs_ch, out, ld);
  drain_to(out, y);
  return 0;
}

template <class IN, class OUT, class COEFF, class ACC, int N, FTYPE ft>
static int run_prog(const std::vector<long long> &x, const std::vector<long long> &c, size_t chunk, std::vector<long long> &y) {
  if ((int)c.size() != N) return 2;
  ac_fir_prog_coeffs<IN, OUT, COEFF, ACC, N, ft> filter;
  COEFF coeffs[N];
  for (int i = 0; i < N; i++) coeffs[i] = from_raw<COEFF>(c[i]);
  ac_channel<IN> in;
  ac_channel<OUT> out;
  if (chunk) {  // the reference bench's pattern: one queued sample, one run() (rtest_ac_fir_prog_coeffs.cpp:109-113)
    for (size_t i = 0; i < x.size(); i++) {
      in.write(from_raw<IN>(x[i]));
      filter.run(in, out, coeffs);
      if (out.debug_size() != i + 1) return 3;
    }
  } else {
    for (size_t i = 0; i < x.size(); i++) in.write(from_raw<IN>(x[i]));
    filter.run(in, out, coeffs);                 // consumes exactly one sample
    if (out.debug_size() != (x.empty() ? 0u : 1u)) return 3;
    filter.run_block(in, out, coeffs);           // extension: the rest in one go
  }
  drain_to(out, y);
  return 0;
}

// ac_fir_reg_share: caller-owned delay line, scalar run(), coefficient RAM per call; out.txt = outputs followed by the
// final ac_firProgCoeffs_delay_line value
template <int N, class IN, class OUT, class COEFF, class ACC, int MWW, int BS, int BO, FTYPE ft>
static int run_reg_share(const std::vector<long long> &x, const std::vector<long long> &ram, std::vector<long long> &y) {
  IN reg[N];
  for (int i = 0; i < N; i++) reg[i] = 0;
  ac_fir_reg_share<N, IN, OUT, COEFF, ACC, MWW, BS, BO, ft> filter(reg);
  std::vector<COEFF> coeffs(ram.size() < (size_t)N ? (size_t)N : ram.size());
  for (size_t i = 0; i < ram.size(); i++) coeffs[i] = from_raw<COEFF>(ram[i]);
  for (size_t k = 0; k < x.size(); k++) {
    IN in = from_raw<IN>(x[k]);
    OUT out;
    filter.run(in, coeffs.data(), out);
    y.push_back(to_raw(out));
    if (to_raw(reg[0]) != x[k]) return 3;      // the caller's delay line is kept up to date
  }
  OUT dl;
  filter.ac_firProgCoeffs_delay_line(dl);
  y.push_back(to_raw(dl));
  return 0;
}

// ac_poly_dec: coefficients as one struct on a channel, whole groups of DF samples per run()
template <class IN, class COEFF, class ACC, class OUT, int NT, int DF>
static int run_poly_dec(const std::vector<long long> &x, const std::vector<long long> &c, size_t chunk, std::vector<long long> &y) {
  struct Str { COEFF coeffs[NT * DF]; };
  if ((int)c.size() != NT * DF) return 2;
  ac_poly_dec<IN, COEFF, Str, ACC, OUT, NT, DF> filter;
  ac_channel<IN> in;
  ac_channel<OUT> out;
  ac_channel<Str> coeffs_st;
  Str junk, good;
  for (int i = 0; i < NT * DF; i++) { junk.coeffs[i] = from_raw<COEFF>(~c[i]); good.coeffs[i] = from_raw<COEFF>(c[i]); }
  coeffs_st.write(junk);
  coeffs_st.write(good);                      // the last struct queued wins
  for (size_t i = 0; i < x.size(); i++) {
    in.write(from_raw<IN>(x[i]));
    if (chunk && (i + 1) % chunk == 0) {
      filter.run(in, out, coeffs_st);
      if (in.debug_size() >= (unsigned)DF) return 3;   // only an incomplete group may stay queued
    }
  }
  filter.run(in, out, coeffs_st);
  drain_to(out, y);
  return 0;
}

// ac_intg_dump: samples interleaved over CHN channels, one n_sample token per frame (passed in place of the taps)
template <class IN, class ACC, class OUT, int NS, int CHN>
static int run_intg_dump(const std::vector<long long> &x, const std::vector<long long> &tokens, std::vector<long long> &y) {
  typedef ac_int<16, false> N_TYPE;
  ac_intg_dump<IN, ACC, OUT, N_TYPE, NS, CHN> filter;
  ac_channel<IN> in;
  ac_channel<OUT> out;
  ac_channel<N_TYPE> n_sample;
  size_t k = 0;
  for (size_t f = 0; f < tokens.size(); f++) {          // two frames per run() call: the sums carry across calls
    const long long n = tokens[f];
    const size_t take = (size_t)((n >= 1 && n <= NS) ? n : NS) * CHN;
    n_sample.write(N_TYPE(n));
    for (size_t i = 0; i < take; i++) in.write(from_raw<IN>(x[k++]));
    if (f % 2 == 1) filter.run(in, out, n_sample);
  }
  filter.run(in, out, n_sample);
  if (k != x.size()) return 3;
  drain_to(out, y);
  return 0;
}

// ac_mv_avg: the wrapper idiom of the manual (pdf p.30: the weights are a member of the DERIVED class), bursts of `chunk`
// samples, the n_sample token written before the call
template <int MAXS, int TAPS, ac_window_mode WT, class IN, class OUT, class ACC, class COEFF>
class mv_wrapper : public ac_mv_avg<MAXS, TAPS, WT, IN, OUT, ACC, COEFF, ac_int<32, false> > {
public:
  COEFF coeffs[TAPS];
  explicit mv_wrapper(const std::vector<long long> &c) : ac_mv_avg<MAXS, TAPS, WT, IN, OUT, ACC, COEFF, ac_int<32, false> >(coeffs) {
    for (int i = 0; i < TAPS; i++) coeffs[i] = from_raw<COEFF>(c[i]);
  }
};
template <int MAXS, int TAPS, ac_window_mode WT, class IN, class OUT, class ACC, class COEFF>
static int run_mv_avg(const std::vector<long long> &x, const std::vector<long long> &c, size_t n_sample, std::vector<long long> &y) {
  if (c.size() != TAPS || n_sample == 0) return 2;
  mv_wrapper<MAXS, TAPS, WT, IN, OUT, ACC, COEFF> filter(c);
  ac_channel<IN> in;
  ac_channel<OUT> out;
  ac_channel<ac_int<32, false> > ns;
  for (size_t i = 0; i < x.size(); i++) in.write(from_raw<IN>(x[i]));
  ns.write(ac_int<32, false>((unsigned)n_sample));
  filter.run(in, out, ns);
  drain_to(out, y);
  return 0;
}

template <class FILTER, class IN, class OUT>
static int run_cic(const std::vector<long long> &x, size_t chunk, std::vector<long long> &y) {
  FILTER filter;
  ac_channel<IN> in;
  ac_channel<OUT> out;
  for (size_t i = 0; i < x.size(); i++) {
    in.write(from_raw<IN>(x[i]));
    if (chunk && (i + 1) % chunk == 0) filter.run(in, out);
  }
  filter.run(in, out);
  drain_to(out, y);
  return 0;
}

int main(int argc, char **argv) {
  if (argc < 5) {
    std::fprintf(stderr, "usage: %s <case> <in.txt> <coef.txt|-> <out.txt> [chunk]\n", argv[0]);
    return 64;
  }
  const std::string name = argv[1];
  const std::vector<long long> x = read_ints(argv[2]);
  const std::vector<long long> c = std::strcmp(argv[3], "-") ? read_ints(argv[3]) : std::vector<long long>();
  const size_t chunk = argc > 5 ? (size_t)std::atoll(argv[5]) : 0;
  std::vector<long long> y;
  int rc = 64;
  try {
    // ---- the reference benches' configurations (tests/rtest_ac_fir_*_coeffs.cpp, tests/ac_cic_*_full_param.h)
    typedef ac_fixed<64, 32, true, AC_TRN, AC_WRAP> acc64;
    if (name == "bench_const")
      rc = run_const<ac_fixed<16, 8, true, AC_TRN, AC_WRAP>, acc64, ac_fixed<32, 16, true, AC_TRN, AC_WRAP>, acc64, 29, FOLD_ODD>(x, c, chunk, y);
    else if (name == "bench_load")
      rc = run_load<ac_fixed<32, 16, true, AC_TRN, AC_WRAP>, acc64, ac_fixed<32, 16, true, AC_TRN, AC_WRAP>, acc64, 27, FOLD_ODD>(x, c, chunk, y);
    else if (name == "bench_prog")
      rc = run_prog<ac_fixed<28, 6, true, AC_TRN, AC_WRAP>, acc64, ac_fixed<23, 7, true, AC_TRN, AC_WRAP>, acc64, 27, FOLD_ODD>(x, c, chunk, y);
    else if (name == "bench_cic_dec")
      rc = run_cic<ac_cic_dec_full<ac_fixed<32, 16, true>, ac_fixed<48, 32, true>, 7, 2, 4>, ac_fixed<32, 16, true>, ac_fixed<48, 32, true> >(x, chunk, y);
    else if (name == "bench_cic_intr")
      rc = run_cic<ac_cic_intr_full<ac_fixed<32, 16, true>, ac_fixed<49, 33, true>, 7, 2, 5>, ac_fixed<32, 16, true>, ac_fixed<49, 33, true> >(x, chunk, y);
    // ---- BASELINE.json configurations
    else if (name == "q15_const16")
      rc = run_const<ac_fixed<16, 1, true>, ac_fixed<40, 8, true>, ac_fixed<16, 1, true>, ac_fixed<40, 8, true>, 16, SHIFT_REG>(x, c, chunk, y);
    else if (name == "q15_load256")
      rc = run_load<ac_fixed<16, 1, true>, ac_fixed<40, 8, true>, ac_fixed<16, 1, true>, ac_fixed<40, 8, true>, 256, SHIFT_REG>(x, c, chunk, y);
    else if (name == "q15_prog1024")
      rc = run_prog<ac_fixed<16, 1, true>, ac_fixed<40, 8, true>, ac_fixed<16, 1, true>, ac_fixed<40, 8, true>, 1024, TRANSPOSED>(x, c, chunk, y);
    else if (name == "q15_cic_dec")
      rc = run_cic<ac_cic_dec_full<ac_fixed<16, 1, true>, ac_fixed<28, 13, true>, 8, 1, 4>, ac_fixed<16, 1, true>, ac_fixed<28, 13, true> >(x, chunk, y);
    else if (name == "q15_cic_intr")
      rc = run_cic<ac_cic_intr_full<ac_fixed<16, 1, true>, ac_fixed<20, 5, true>, 4, 1, 3>, ac_fixed<16, 1, true>, ac_fixed<20, 5, true> >(x, chunk, y);
    // ---- ac_fir_reg_share (oracle/ref_configs.py RS_CONFIGS 2, 4, 7, 9)
    else if (name == "rs2")
      rc = run_reg_share<16, ac_fixed<16, 1, true>, ac_fixed<40, 8, true>, ac_fixed<16, 1, true>, ac_fixed<40, 8, true>, 1, 1, 0, FOLD_EVEN_ANTI>(x, c, y);
    else if (name == "rs4")
      rc = run_reg_share<15, ac_fixed<16, 1, true>, ac_fixed<40, 8, true>, ac_fixed<16, 1, true>, ac_fixed<40, 8, true>, 1, 1, 0, FOLD_ODD_ANTI>(x, c, y);
    else if (name == "rs7")
      rc = run_reg_share<24, ac_fixed<16, 1, true>, ac_fixed<40, 8, true>, ac_fixed<16, 1, true>, ac_fixed<40, 8, true>, 8, 4, 4, FOLD_EVEN_ANTI>(x, c, y);
    else if (name == "rs9")
      rc = run_reg_share<15, ac_fixed<16, 1, true>, ac_fixed<16, 1, true>, ac_fixed<16, 1, true>, ac_fixed<24, 4, true>, 1, 1, 0, FOLD_ODD_ANTI>(x, c, y);
    // ---- ac_poly_dec (oracle/ref_configs.py PD_CONFIGS 2, 6)
    else if (name == "pd2")
      rc = run_poly_dec<ac_fixed<16, 1, true>, ac_fixed<16, 1, true>, ac_fixed<40, 8, true>, ac_fixed<40, 8, true>, 32, 8>(x, c, chunk, y);
    else if (name == "pd6")
      rc = run_poly_dec<ac_fixed<16, 1, true>, ac_fixed<16, 1, true>, ac_fixed<24, 4, true>, ac_fixed<16, 1, true>, 16, 4>(x, c, chunk, y);
    // ---- ac_intg_dump (oracle/ref_configs.py ID_CONFIGS 0, 3)
    else if (name == "id0")
      rc = run_intg_dump<ac_fixed<16, 1, true>, ac_fixed<32, 17, true>, ac_fixed<32, 17, true>, 64, 4>(x, c, y);
    else if (name == "id3")
      rc = run_intg_dump<ac_fixed<16, 1, true>, ac_fixed<24, 12, true>, ac_fixed<16, 8, true>, 16, 3>(x, c, y);
    // ---- ac_mv_avg (oracle/ref_configs.py MV_CONFIGS 0, 3, 8); `chunk` carries n_sample
    else if (name == "mv0")
      rc = run_mv_avg<1024, 7, AC_CLIP, ac_fixed<16, 2, true>, ac_fixed<16, 2, true>, ac_fixed<16, 2, true>, ac_fixed<16, 2, true> >(x, c, chunk, y);
    else if (name == "mv3")
      rc = run_mv_avg<4096, 31, AC_MIRROR, ac_fixed<16, 1, true>, ac_fixed<40, 8, true>, ac_fixed<40, 8, true>, ac_fixed<16, 1, true> >(x, c, chunk, y);
    else if (name == "mv8")
      rc = run_mv_avg<64, 5, AC_WIN, ac_fixed<32, 16, true>, ac_fixed<64, 32, true>, ac_fixed<64, 32, true>, ac_fixed<32, 16, true> >(x, c, chunk, y);
    else
      std::fprintf(stderr, "unknown case %s\n", name.c_str());
  } catch (const b200dsp::engine_error &e) {
    std::fprintf(stderr, "engine_error %d: %s\n", e.status(), e.what());
    return 70;
  }
  if (rc) return rc;
  std::ofstream o(argv[4]);
  for (size_t i = 0; i < y.size(); i++) o << y[i] << "\n";
  return 0;
}
