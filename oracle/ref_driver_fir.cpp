// oracle/ref_driver_fir.cpp -- TEST INFRASTRUCTURE (Oracle A), not product code.
//
// Instantiates the UNMODIFIED reference FIR class templates (found by
// -I$AC_DSP_REF/include, default /root/reference/include; nothing is copied) over the
// clean-room datatype shim in oracle/ac_shim, for every configuration listed in
// oracle/ref_configs.py, and exposes them through a tiny C interface that moves raw
// two's-complement integers.  Built into oracle/_ref/libacdsp_ref.so by oracle/Makefile.
//
// Driving pattern per class follows the reference benches:
//   const : coefficient pointer at construction, all samples queued, one run()
//           (tests/rtest_ac_fir_const_coeffs.cpp:93-108,160)
//   load  : run() with ld=true and N_TAPS coefficients queued, then run() on samples
//           (tests/rtest_ac_fir_load_coeffs.cpp:135-147)
//   prog  : one run() per sample with the coefficient array as an argument
//           (tests/rtest_ac_fir_prog_coeffs.cpp:109-113)
#include <ac_fixed.h>
#include <ac_int.h>
#include <ac_channel.h>
#include <ac_dsp/ac_fir_const_coeffs.h>
#include <ac_dsp/ac_fir_load_coeffs.h>
#include <ac_dsp/ac_fir_prog_coeffs.h>

#include <chrono>
#include <vector>

namespace {

typedef std::chrono::steady_clock Clock;
static inline double secs(Clock::time_point a, Clock::time_point b) { return std::chrono::duration<double>(b - a).count(); }

struct FirBase {
  double run_seconds;  // time spent inside the reference's run() during the last acref_fir_run (channel fill / drain excluded)
  FirBase() : run_seconds(0) {}
  virtual ~FirBase() {}
  virtual int load(const long long *c) = 0;
  virtual long run(const long long *in, long n, long long *out) = 0;
};

template <class IN, class OUT, class COEFF, class ACC, unsigned NT, FTYPE FT>
struct FirConst : FirBase {
  COEFF coeffs[NT];
  ac_fir_const_coeffs<IN, OUT, COEFF, ACC, NT, FT> *f;
  ac_channel<IN> in_ch;
  ac_channel<OUT> out_ch;
  FirConst() : f(0) {}
  ~FirConst() { delete f; }
  int load(const long long *c) {
    if (f) return -1;  // constant coefficients: fixed at construction
    for (unsigned i = 0; i < NT; i++) coeffs[i] = ac_shim::from_raw<COEFF>(c[i]);
    f = new ac_fir_const_coeffs<IN, OUT, COEFF, ACC, NT, FT>(coeffs);
    return 0;
  }
  long run(const long long *in, long n, long long *out) {
    if (!f) return -1;
    for (long i = 0; i < n; i++) in_ch.write(ac_shim::from_raw<IN>(in[i]));
    Clock::time_point t0 = Clock::now();
    f->run(in_ch, out_ch);
    run_seconds = secs(t0, Clock::now());
    long k = 0;
    while (out_ch.available(1)) out[k++] = ac_shim::to_raw(out_ch.read());
    return k;
  }
};

template <class IN, class OUT, class COEFF, class ACC, unsigned NT, FTYPE FT>
struct FirLoad : FirBase {
  ac_fir_load_coeffs<IN, OUT, COEFF, ACC, NT, FT> f;
  ac_channel<IN> in_ch;
  ac_channel<OUT> out_ch;
  ac_channel<COEFF> c_ch;
  ac_channel<bool> ld;
  int load(const long long *c) {
    for (unsigned i = 0; i < NT; i++) c_ch.write(ac_shim::from_raw<COEFF>(c[i]));
    ld.write(true);
    f.run(in_ch, c_ch, out_ch, ld);
    return 0;
  }
  long run(const long long *in, long n, long long *out) {
    for (long i = 0; i < n; i++) in_ch.write(ac_shim::from_raw<IN>(in[i]));
    ld.write(false);
    Clock::time_point t0 = Clock::now();
    f.run(in_ch, c_ch, out_ch, ld);
    run_seconds = secs(t0, Clock::now());
    long k = 0;
    while (out_ch.available(1)) out[k++] = ac_shim::to_raw(out_ch.read());
    return k;
  }
};

template <class IN, class OUT, class COEFF, class ACC, unsigned NT, FTYPE FT>
struct FirProg : FirBase {
  ac_fir_prog_coeffs<IN, OUT, COEFF, ACC, (int)NT, FT> f;
  COEFF coeffs[NT];
  ac_channel<IN> in_ch;
  ac_channel<OUT> out_ch;
  int load(const long long *c) {
    for (unsigned i = 0; i < NT; i++) coeffs[i] = ac_shim::from_raw<COEFF>(c[i]);
    return 0;
  }
  long run(const long long *in, long n, long long *out) {
    long k = 0;
    Clock::time_point t0 = Clock::now();  // one run() per sample, as the reference bench does: the whole loop is the run
    for (long i = 0; i < n; i++) {
      in_ch.write(ac_shim::from_raw<IN>(in[i]));
      f.run(in_ch, out_ch, coeffs);
      while (out_ch.available(1)) out[k++] = ac_shim::to_raw(out_ch.read());
    }
    run_seconds = secs(t0, Clock::now());
    return k;
  }
};

template <class IN, class OUT, class COEFF, class ACC, unsigned NT, FTYPE FT>
FirBase *make_cls(int cls) {
  switch (cls) {
    case 0: return new FirConst<IN, OUT, COEFF, ACC, NT, FT>();
    case 1: return new FirLoad<IN, OUT, COEFF, ACC, NT, FT>();
    case 2: return new FirProg<IN, OUT, COEFF, ACC, NT, FT>();
  }
  return 0;
}

template <class IN, class OUT, class COEFF, class ACC, unsigned NT>
FirBase *make_ft(int cls, int ftype) {
  switch (ftype) {
    case SHIFT_REG: return make_cls<IN, OUT, COEFF, ACC, NT, SHIFT_REG>(cls);
    case ROTATE_SHIFT: return make_cls<IN, OUT, COEFF, ACC, NT, ROTATE_SHIFT>(cls);
    case C_BUFF: return make_cls<IN, OUT, COEFF, ACC, NT, C_BUFF>(cls);
    case FOLD_EVEN: return make_cls<IN, OUT, COEFF, ACC, NT, FOLD_EVEN>(cls);
    case FOLD_ODD: return make_cls<IN, OUT, COEFF, ACC, NT, FOLD_ODD>(cls);
    case TRANSPOSED: return make_cls<IN, OUT, COEFF, ACC, NT, TRANSPOSED>(cls);
  }
  return 0;  // FOLD_*_ANTI: the reference classes leave core_out unwritten; not instantiated
}

}  // namespace

extern "C" {

void *acref_fir_create(int cfg, int cls, int ftype) {
  switch (cfg) {
#define X(ID, iW, iI, iS, iQ, iO, cW, cI, cS, cQ, cO, aW, aI, aS, aQ, aO, oW, oI, oS, oQ, oO, NT)        \
  case ID:                                                                                              \
    return make_ft<ac_fixed<iW, iI, iS, iQ, iO>, ac_fixed<oW, oI, oS, oQ, oO>, ac_fixed<cW, cI, cS, cQ, cO>, \
                   ac_fixed<aW, aI, aS, aQ, aO>, NT>(cls, ftype);
#include "_ref/cfgs_fir.inc"
#undef X
  }
  return 0;
}
int acref_fir_load(void *h, const long long *c) { return ((FirBase *)h)->load(c); }
long acref_fir_run(void *h, const long long *in, long n, long long *out) { return ((FirBase *)h)->run(in, n, out); }
double acref_fir_last_seconds(void *h) { return ((FirBase *)h)->run_seconds; }
void acref_fir_destroy(void *h) { delete (FirBase *)h; }

}  // extern "C"
