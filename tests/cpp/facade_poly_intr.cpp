// tests/cpp/facade_poly_intr.cpp -- drives ac_poly_intr through the header facade the way reference user code would: one
// run() call per read_ctrl token (a load of the control / coefficient structures, or one sample).  Separate from
// facade_bench.cpp because ac_poly_intr.h declares its own FTYPE enum (ac_poly_intr.h:95), which cannot share a
// translation unit with the FIR headers' -- in the reference as here.
//
//   facade_poly_intr <case> <in.txt> <ctl.txt> <out.txt>
//
// ctl.txt: c1[COEFFSZ] c2[COEFFSZ] sign[IF] corr[IF] half -- the first set is loaded before the stream, the second
// (with the complemented signs and the reversed corr table) after `half` samples, as tests/golden/make_golden.py did.
#include <ac_fixed.h>
#include <ac_int.h>
#include <ac_channel.h>
#include <ac_dsp/ac_poly_intr.h>

#include <cstdio>
#include <fstream>
#include <string>
#include <vector>

static std::vector<long long> read_ints(const char *path) {
  std::vector<long long> v;
  std::ifstream f(path);
  long long x;
  while (f >> x) v.push_back(x);
  return v;
}

template <class IN, class COEFF, class ACC, class OUT, int NT, int CSZ, int IF, FTYPE ft>
static int run_case(const std::vector<long long> &x, const std::vector<long long> &ctl, std::vector<long long> &y) {
  struct Ctrl { bool sign[IF]; ac_int<8, false> corr[IF]; };
  struct Coef { COEFF coeffs[CSZ]; };
  if ((int)ctl.size() != 2 * CSZ + 2 * IF + 1) return 2;
  ac_poly_intr<IN, COEFF, ACC, OUT, Ctrl, Coef, NT, CSZ, IF, ft> filter;
  ac_channel<IN> in;
  ac_channel<OUT> out;
  ac_channel<Ctrl> ctrl_st;
  ac_channel<Coef> coeffs_st;
  ac_channel<bool> rd;
  const size_t half = (size_t)ctl[2 * CSZ + 2 * IF];
  for (size_t i = 0; i <= x.size(); i++) {
    if (i == 0 || i == half) {
      Ctrl t;
      Coef k;
      for (int c = 0; c < CSZ; c++) k.coeffs[c] = b200dsp::fixed_traits<COEFF>::from_raw(ctl[(i ? CSZ : 0) + c]);
      for (int j = 0; j < IF; j++) {
        const long long s = ctl[2 * CSZ + j];
        t.sign[j] = i ? !s : (s != 0);
        t.corr[j] = ac_int<8, false>((int)ctl[2 * CSZ + IF + (i ? IF - 1 - j : j)]);
      }
      ctrl_st.write(t);
      coeffs_st.write(k);
      rd.write(true);
      filter.run(in, out, ctrl_st, coeffs_st, rd);
    }
    if (i == x.size()) break;
    in.write(b200dsp::fixed_traits<IN>::from_raw(x[i]));
    rd.write(false);
    filter.run(in, out, ctrl_st, coeffs_st, rd);
    while (out.available(1)) y.push_back(b200dsp::fixed_traits<OUT>::to_raw(out.read()));
  }
  return 0;
}

int main(int argc, char **argv) {
  if (argc < 5) { std::fprintf(stderr, "usage: facade_poly_intr <case> <in.txt> <ctl.txt> <out.txt>\n"); return 64; }
  const std::string name = argv[1];
  const std::vector<long long> x = read_ints(argv[2]), ctl = read_ints(argv[3]);
  std::vector<long long> y;
  int rc = 1;
  typedef ac_fixed<16, 1, true> Q15;
  typedef ac_fixed<40, 8, true> ACC40;
  try {
    // oracle/ref_configs.py PI_CONFIGS 0, 1, 2, 9
    if (name == "pi0") rc = run_case<Q15, Q15, ACC40, ACC40, 8, 16, 4, FOLD_EVEN>(x, ctl, y);
    else if (name == "pi1") rc = run_case<Q15, Q15, ACC40, ACC40, 7, 16, 4, FOLD_ODD>(x, ctl, y);
    else if (name == "pi2") rc = run_case<Q15, Q15, ACC40, ACC40, 16, 64, 4, FOLD_ANTI>(x, ctl, y);
    else if (name == "pi9") rc = run_case<Q15, Q15, ac_fixed<24, 4, true>, Q15, 8, 16, 4, FOLD_EVEN>(x, ctl, y);
    else std::fprintf(stderr, "unknown case %s\n", name.c_str());
  } catch (const b200dsp::engine_error &e) {
    std::fprintf(stderr, "engine_error %d: %s\n", e.status(), e.what());
    return 70;
  }
  if (rc) return rc;
  std::ofstream o(argv[4]);
  for (size_t i = 0; i < y.size(); i++) o << y[i] << "\n";
  return 0;
}
