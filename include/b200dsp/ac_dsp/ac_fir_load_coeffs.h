// b200dsp facade: ac_fir_load_coeffs on the B200 engine.
//
// Drop-in for hlslibs/ac_dsp include/ac_dsp/ac_fir_load_coeffs.h:300-365 -- same class name, template parameters and
// run() signature.  The include guard is the reference's, so a later #include of the original header is a no-op.
#ifndef _INCLUDED_AC_FIR_LOAD_COEFFS_H_
#define _INCLUDED_AC_FIR_LOAD_COEFFS_H_

#include "../fir_block.h"

template <class IN_TYPE, class OUT_TYPE, class COEFF_TYPE, class ACC_TYPE, unsigned N_TAPS, FTYPE ftype>
class ac_fir_load_coeffs {
public:
  ac_fir_load_coeffs() {}

  // Load protocol of the reference (:324-331): at most ONE ld token is consumed per call; the taps are taken only when
  // that token is true AND N_TAPS values are queued on coeffs_ch, otherwise the token is silently dropped.  Then every
  // queued sample is filtered (:335-364).  Samples that arrive before any load make the reference compute with
  // uninitialised (AC_VAL_DC) taps; the engine refuses instead (engine_error, B2D_ESTATE).
  void run(ac_channel<IN_TYPE> &data_in, ac_channel<COEFF_TYPE> &coeffs_ch, ac_channel<OUT_TYPE> &data_out,
           ac_channel<bool> &ld) {
    if (ld.available(1)) {
      const bool ld_t = ld.read();
      if (ld_t && coeffs_ch.available(N_TAPS)) {
        COEFF_TYPE c[N_TAPS];
        for (unsigned i = 0; i < N_TAPS; i++) c[i] = coeffs_ch.read();
        blk.load(c);
      }
    }
    blk.process(data_in, data_out);
  }

  // extensions: raw-array forms of the two phases
  void load_raw(const typename b200dsp::container_sel<COEFF_TYPE::width>::type *taps) { blk.load_raw(taps); }
  void run_raw(const typename b200dsp::container_sel<IN_TYPE::width>::type *in, size_t n,
               typename b200dsp::container_sel<OUT_TYPE::width>::type *out) { blk.process_raw(in, n, out); }

private:
  b200dsp::fir_block<IN_TYPE, OUT_TYPE, COEFF_TYPE, ACC_TYPE, N_TAPS, (int)ftype, B2D_FIR_LOAD> blk;
};

#endif
