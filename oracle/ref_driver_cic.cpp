// oracle/ref_driver_cic.cpp -- TEST INFRASTRUCTURE (Oracle A), not product code.
//
// Instantiates the UNMODIFIED reference CIC class templates over the clean-room shim for
// every configuration in oracle/ref_configs.py.  Compiled twice (-DACREF_CIC_DEC /
// -DACREF_CIC_INTR) because ac_cic_dec_full.h and ac_cic_intr_full.h both define an
// unguarded `template<int,int> struct power` and cannot share a translation unit.
// run() is called exactly like tests/rtest_ac_cic_{dec,intr}_full.cpp do: queue every
// input, one run(), drain the output channel.  The object persists between calls, so
// repeated acref_cic_run() calls reproduce the reference's streaming behaviour.
#include <ac_fixed.h>
#include <ac_int.h>
#include <ac_channel.h>
#if defined(ACREF_CIC_DEC)
#include <ac_dsp/ac_cic_dec_full.h>
#define CIC_CLASS ac_cic_dec_full
#define CREATE_FN acref_cic_dec_create
#define INC_FILE "_ref/cfgs_cic_dec.inc"
#else
#include <ac_dsp/ac_cic_intr_full.h>
#define CIC_CLASS ac_cic_intr_full
#define CREATE_FN acref_cic_intr_create
#define INC_FILE "_ref/cfgs_cic_intr.inc"
#endif

#include <chrono>

#include "ref_driver_cic.h"

namespace {
template <class IN, class OUT, unsigned R, unsigned M, unsigned N>
struct CicImpl : acref::CicBase {
  CIC_CLASS<IN, OUT, R, M, N> f;
  ac_channel<IN> in_ch;
  ac_channel<OUT> out_ch;
  long run(const long long *in, long n, long long *out) {
    for (long i = 0; i < n; i++) in_ch.write(ac_shim::from_raw<IN>(in[i]));
    std::chrono::steady_clock::time_point t0 = std::chrono::steady_clock::now();
    f.run(in_ch, out_ch);
    run_seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    long k = 0;
    while (out_ch.available(1)) out[k++] = ac_shim::to_raw(out_ch.read());
    return k;
  }
};
}  // namespace

extern "C" void *CREATE_FN(int cfg) {
  switch (cfg) {
#define X(ID, R, M, N, iW, iI, iS, iQ, iO, oW, oI, oS, oQ, oO) \
  case ID:                                                     \
    return new CicImpl<ac_fixed<iW, iI, iS, iQ, iO>, ac_fixed<oW, oI, oS, oQ, oO>, R, M, N>();
#include INC_FILE
#undef X
  }
  return 0;
}

#if defined(ACREF_CIC_DEC)
extern "C" long acref_cic_run(void *h, const long long *in, long n, long long *out) {
  return ((acref::CicBase *)h)->run(in, n, out);
}
extern "C" double acref_cic_last_seconds(void *h) { return ((acref::CicBase *)h)->run_seconds; }
extern "C" void acref_cic_destroy(void *h) { delete (acref::CicBase *)h; }
#endif
