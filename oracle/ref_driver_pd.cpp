// oracle/ref_driver_pd.cpp -- TEST INFRASTRUCTURE (Oracle A), not product code.
//
// Instantiates the UNMODIFIED reference class template ac_poly_dec (include/ac_dsp/ac_poly_dec.h:87-137, found by
// -I$AC_DSP_REF/include; nothing is copied) over the clean-room datatype shim for the configurations in
// oracle/ref_configs.py (PD_CONFIGS) behind a tiny C interface on raw integers.  Coefficients travel as the reference
// wants them: one struct holding coeffs[NTAPS * DF] on a channel; the last struct queued wins (:101-106).
#include <ac_fixed.h>
#include <ac_int.h>
#include <ac_channel.h>
#include <ac_dsp/ac_poly_dec.h>

namespace {

struct PdBase {
  virtual ~PdBase() {}
  virtual void load(const long long *c) = 0;
  virtual long run(const long long *in, long n, long long *out) = 0;
};

#define A4(W, I, S, Q, O) ac_fixed<W, I, S, Q, O>

template <class IN, class COEFF, class ACC, class OUT, int NT, int DF>
struct Pd : PdBase {
  struct Str { COEFF coeffs[NT * DF]; };
  ac_poly_dec<IN, COEFF, Str, ACC, OUT, NT, DF> f;
  ac_channel<IN> in_ch;
  ac_channel<OUT> out_ch;
  ac_channel<Str> c_ch;
  void load(const long long *c) {
    Str s;
    for (int i = 0; i < NT * DF; i++) s.coeffs[i] = ac_shim::from_raw<COEFF>(c[i]);
    c_ch.write(s);
  }
  long run(const long long *in, long n, long long *out) {
    for (long i = 0; i < n; i++) in_ch.write(ac_shim::from_raw<IN>(in[i]));
    f.run(in_ch, out_ch, c_ch);     // consumes whole groups of DF samples; the rest stays queued (:107-109)
    long k = 0;
    while (out_ch.available(1)) out[k++] = ac_shim::to_raw(out_ch.read());
    return k;
  }
};

}  // namespace

extern "C" {
void *acref_pd_create(int cfg) {
  switch (cfg) {
#define X(id, iW, iI, iS, iQ, iO, cW, cI, cS, cQ, cO, aW, aI, aS, aQ, aO, oW, oI, oS, oQ, oO, NT, DF) \
  case id: return new Pd<A4(iW, iI, iS, iQ, iO), A4(cW, cI, cS, cQ, cO), A4(aW, aI, aS, aQ, aO), A4(oW, oI, oS, oQ, oO), NT, DF>();
#include "_ref/cfgs_pd.inc"
#undef X
  }
  return 0;
}
void acref_pd_load(void *h, const long long *c) { ((PdBase *)h)->load(c); }
long acref_pd_run(void *h, const long long *in, long n, long long *out) { return ((PdBase *)h)->run(in, n, out); }
void acref_pd_destroy(void *h) { delete (PdBase *)h; }
}
