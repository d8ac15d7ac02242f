// oracle/ref_driver_mv.cpp -- TEST INFRASTRUCTURE (Oracle A), not product code.
//
// Instantiates the UNMODIFIED reference class template ac_mv_avg (include/ac_dsp/ac_mv_avg.h:140-204, found by
// -I$AC_DSP_REF/include; nothing is copied) over the clean-room datatype shim and the RESTATED ac_window_1d_flag
// (oracle/ac_shim/ac_window.h: parity unpinned, see its header) for the configurations in oracle/ref_configs.py
// (MV_CONFIGS).  One call = queue the samples and the n_sample token, run(), drain.
#include <ac_fixed.h>
#include <ac_int.h>
#include <ac_channel.h>
#include <ac_dsp/ac_mv_avg.h>

#include <vector>

namespace {

struct MvBase {
  virtual ~MvBase() {}
  virtual long run(const long long *in, long n, long long n_sample, long long *out) = 0;
};

#define A4(W, I, S, Q, O) ac_fixed<W, I, S, Q, O>

template <int MAXS, int TAPS, ac_window_mode WT, class IN, class OUT, class ACC, class COEFF>
struct Mv : MvBase {
  typedef ac_int<32, false> S_TYPE;
  std::vector<COEFF> coeffs;
  ac_mv_avg<MAXS, TAPS, WT, IN, OUT, ACC, COEFF, S_TYPE> f;
  ac_channel<IN> in_ch;
  ac_channel<OUT> out_ch;
  ac_channel<S_TYPE> n_ch;
  explicit Mv(const long long *c) : coeffs(make(c)), f(coeffs.data()) {}
  static std::vector<COEFF> make(const long long *c) {
    std::vector<COEFF> v(TAPS);
    for (int i = 0; i < TAPS; i++) v[i] = ac_shim::from_raw<COEFF>(c[i]);
    return v;
  }
  long run(const long long *in, long n, long long n_sample, long long *out) {
    for (long i = 0; i < n; i++) in_ch.write(ac_shim::from_raw<IN>(in[i]));
    n_ch.write(S_TYPE(n_sample));
    f.run(in_ch, out_ch, n_ch);
    long k = 0;
    while (out_ch.available(1)) out[k++] = ac_shim::to_raw(out_ch.read());
    return k;
  }
};

}  // namespace

extern "C" {
void *acref_mv_create(int cfg, const long long *coeffs) {
  switch (cfg) {
#define X(id, MAXS, TAPS, WT, iW, iI, iS, iQ, iO, oW, oI, oS, oQ, oO, aW, aI, aS, aQ, aO, cW, cI, cS, cQ, cO) \
  case id: return new Mv<MAXS, TAPS, WT, A4(iW, iI, iS, iQ, iO), A4(oW, oI, oS, oQ, oO), A4(aW, aI, aS, aQ, aO), A4(cW, cI, cS, cQ, cO)>(coeffs);
#include "_ref/cfgs_mv.inc"
#undef X
  }
  return 0;
}
long acref_mv_run(void *h, const long long *in, long n, long long n_sample, long long *out) { return ((MvBase *)h)->run(in, n, n_sample, out); }
void acref_mv_destroy(void *h) { delete (MvBase *)h; }
}
