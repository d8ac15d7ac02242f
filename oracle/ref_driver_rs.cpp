// oracle/ref_driver_rs.cpp -- TEST INFRASTRUCTURE (Oracle A), not product code.
//
// Instantiates the UNMODIFIED reference class template ac_fir_reg_share (found by -I$AC_DSP_REF/include; nothing is
// copied) over the clean-room datatype shim for the configurations in oracle/ref_configs.py (RS_CONFIGS) behind a tiny
// C interface on raw integers.  The reference object works on a caller-owned delay line and takes ONE sample per
// run() call together with the coefficient RAM (include/ac_dsp/ac_fir_reg_share.h:257-303); this driver owns the delay
// line (zero-initialised, like the registers of the other FIR classes) and loops over the samples.
#include <ac_fixed.h>
#include <ac_int.h>
#include <ac_channel.h>
#include <ac_dsp/ac_fir_reg_share.h>

#include <vector>

namespace {

struct RsBase {
  virtual ~RsBase() {}
  virtual long run(const long long *in, long n, const long long *ram, long long *out) = 0;
  virtual long long delay_out() = 0;
  virtual int ram_words() = 0;
};

#define A4(W, I, S, Q, O) ac_fixed<W, I, S, Q, O>

template <int NT, class IN, class OUT, class COEFF, class ACC, int MWW, int BS, int BO, FTYPE FT, int RAM>
struct Rs : RsBase {
  IN reg[NT];
  COEFF ram_c[RAM];
  ac_fir_reg_share<NT, IN, OUT, COEFF, ACC, MWW, BS, BO, FT> f;
  Rs() : f(reg) { for (int i = 0; i < NT; i++) reg[i] = 0; }
  long run(const long long *in, long n, const long long *ram, long long *out) {
    for (int i = 0; i < RAM; i++) ram_c[i] = ac_shim::from_raw<COEFF>(ram[i]);
    for (long k = 0; k < n; k++) {
      IN x = ac_shim::from_raw<IN>(in[k]);
      OUT y;
      f.run(x, ram_c, y);
      out[k] = ac_shim::to_raw(y);
    }
    return n;
  }
  long long delay_out() { OUT y; f.ac_firProgCoeffs_delay_line(y); return ac_shim::to_raw(y); }
  int ram_words() { return RAM; }
};

}  // namespace

extern "C" {

void *acref_rs_create(int cfg) {
  switch (cfg) {
#define X(id, NT, iW, iI, iS, iQ, iO, oW, oI, oS, oQ, oO, cW, cI, cS, cQ, cO, aW, aI, aS, aQ, aO, MWW, BS, BO, FT, RAM) \
  case id: return new Rs<NT, A4(iW, iI, iS, iQ, iO), A4(oW, oI, oS, oQ, oO), A4(cW, cI, cS, cQ, cO), A4(aW, aI, aS, aQ, aO), MWW, BS, BO, FT, RAM>();
#include "_ref/cfgs_rs.inc"
#undef X
  }
  return 0;
}
long acref_rs_run(void *h, const long long *in, long n, const long long *ram, long long *out) { return ((RsBase *)h)->run(in, n, ram, out); }
long long acref_rs_delay_out(void *h) { return ((RsBase *)h)->delay_out(); }
int acref_rs_ram_words(void *h) { return ((RsBase *)h)->ram_words(); }
void acref_rs_destroy(void *h) { delete (RsBase *)h; }
}
