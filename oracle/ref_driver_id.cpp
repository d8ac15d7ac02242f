// oracle/ref_driver_id.cpp -- TEST INFRASTRUCTURE (Oracle A), not product code.
//
// Instantiates the UNMODIFIED reference class template ac_intg_dump (include/ac_dsp/ac_intg_dump.h:113-151, found by
// -I$AC_DSP_REF/include; nothing is copied) over the clean-room datatype shim for the configurations in
// oracle/ref_configs.py (ID_CONFIGS).  One call = queue the samples and the n_sample tokens, run(), drain.
#include <ac_fixed.h>
#include <ac_int.h>
#include <ac_channel.h>
#include <ac_dsp/ac_intg_dump.h>

namespace {

struct IdBase {
  virtual ~IdBase() {}
  virtual long run(const long long *in, long n, const long long *ns, long nf, long long *out) = 0;
};

#define A4(W, I, S, Q, O) ac_fixed<W, I, S, Q, O>

template <class IN, class ACC, class OUT, int NS, int CHN>
struct Id : IdBase {
  typedef ac_int<16, false> N_TYPE;
  ac_intg_dump<IN, ACC, OUT, N_TYPE, NS, CHN> f;
  ac_channel<IN> in_ch;
  ac_channel<OUT> out_ch;
  ac_channel<N_TYPE> n_ch;
  long run(const long long *in, long n, const long long *ns, long nf, long long *out) {
    for (long i = 0; i < n; i++) in_ch.write(ac_shim::from_raw<IN>(in[i]));
    for (long i = 0; i < nf; i++) n_ch.write(N_TYPE((long long)ns[i]));
    f.run(in_ch, out_ch, n_ch);
    long k = 0;
    while (out_ch.available(1)) out[k++] = ac_shim::to_raw(out_ch.read());
    return k;
  }
};

}  // namespace

extern "C" {
void *acref_id_create(int cfg) {
  switch (cfg) {
#define X(id, iW, iI, iS, iQ, iO, aW, aI, aS, aQ, aO, oW, oI, oS, oQ, oO, NS, CHN) \
  case id: return new Id<A4(iW, iI, iS, iQ, iO), A4(aW, aI, aS, aQ, aO), A4(oW, oI, oS, oQ, oO), NS, CHN>();
#include "_ref/cfgs_id.inc"
#undef X
  }
  return 0;
}
long acref_id_run(void *h, const long long *in, long n, const long long *ns, long nf, long long *out) { return ((IdBase *)h)->run(in, n, ns, nf, out); }
void acref_id_destroy(void *h) { delete (IdBase *)h; }
}
