// oracle/ref_driver_pi.cpp -- TEST INFRASTRUCTURE (Oracle A), not product code.
//
// Instantiates the UNMODIFIED reference class template ac_poly_intr (include/ac_dsp/ac_poly_intr.h:261-312, found by
// -I$AC_DSP_REF/include; nothing is copied) over the clean-room datatype shim for the configurations in
// oracle/ref_configs.py (PI_CONFIGS).  One reference run() call either loads the control / coefficient structures
// (read_ctrl token true, :286-288) or consumes one sample (:289-306); the driver issues one call per token.
#include <ac_fixed.h>
#include <ac_int.h>
#include <ac_channel.h>
#include <ac_dsp/ac_poly_intr.h>

namespace {

struct PiBase {
  virtual ~PiBase() {}
  virtual void load(const long long *c, const long long *sign, const long long *corr) = 0;
  virtual long run(const long long *in, long n, long long *out) = 0;
};

#define A4(W, I, S, Q, O) ac_fixed<W, I, S, Q, O>

template <class IN, class COEFF, class ACC, class OUT, int NT, int CSZ, int IF, FTYPE ft>
struct Pi : PiBase {
  struct Ctrl { bool sign[IF]; ac_int<8, false> corr[IF]; };
  struct Coef { COEFF coeffs[CSZ]; };
  ac_poly_intr<IN, COEFF, ACC, OUT, Ctrl, Coef, NT, CSZ, IF, ft> f;
  ac_channel<IN> in_ch;
  ac_channel<OUT> out_ch;
  ac_channel<Ctrl> ctrl_ch;
  ac_channel<Coef> coef_ch;
  ac_channel<bool> rd_ch;
  void load(const long long *c, const long long *sign, const long long *corr) {
    Ctrl t;
    Coef k;
    for (int i = 0; i < IF; i++) { t.sign[i] = sign[i] != 0; t.corr[i] = ac_int<8, false>(corr[i]); }
    for (int i = 0; i < CSZ; i++) k.coeffs[i] = ac_shim::from_raw<COEFF>(c[i]);
    ctrl_ch.write(t);
    coef_ch.write(k);
    rd_ch.write(true);
    f.run(in_ch, out_ch, ctrl_ch, coef_ch, rd_ch);
  }
  long run(const long long *in, long n, long long *out) {
    long k = 0;
    for (long i = 0; i < n; i++) {
      in_ch.write(ac_shim::from_raw<IN>(in[i]));
      rd_ch.write(false);
      f.run(in_ch, out_ch, ctrl_ch, coef_ch, rd_ch);
      while (out_ch.available(1)) out[k++] = ac_shim::to_raw(out_ch.read());
    }
    return k;
  }
};

}  // namespace

extern "C" {
void *acref_pi_create(int cfg) {
  switch (cfg) {
#define X(id, iW, iI, iS, iQ, iO, cW, cI, cS, cQ, cO, aW, aI, aS, aQ, aO, oW, oI, oS, oQ, oO, NT, CSZ, IF, FT) \
  case id: return new Pi<A4(iW, iI, iS, iQ, iO), A4(cW, cI, cS, cQ, cO), A4(aW, aI, aS, aQ, aO), A4(oW, oI, oS, oQ, oO), NT, CSZ, IF, FT>();
#include "_ref/cfgs_pi.inc"
#undef X
  }
  return 0;
}
void acref_pi_load(void *h, const long long *c, const long long *sign, const long long *corr) { ((PiBase *)h)->load(c, sign, corr); }
long acref_pi_run(void *h, const long long *in, long n, long long *out) { return ((PiBase *)h)->run(in, n, out); }
void acref_pi_destroy(void *h) { delete (PiBase *)h; }
}
