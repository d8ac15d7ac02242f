// b200dsp facade: ac_fir_prog_coeffs on the B200 engine.
//
// Drop-in for hlslibs/ac_dsp include/ac_dsp/ac_fir_prog_coeffs.h:261-303 -- same class name, template parameters
// (N_TAPS is `int` here, ftype defaults to SHIFT_REG) and run() signature.  The include guard is the reference's.
#ifndef _INCLUDED_AC_FIR_PROG_COEFFS_H_
#define _INCLUDED_AC_FIR_PROG_COEFFS_H_

#include "../fir_block.h"

template <class IN_TYPE, class OUT_TYPE, class COEFF_TYPE, class ACC_TYPE, int N_TAPS, FTYPE ftype = SHIFT_REG>
class ac_fir_prog_coeffs {
  typedef b200dsp::fir_block<IN_TYPE, OUT_TYPE, COEFF_TYPE, ACC_TYPE, (unsigned)N_TAPS, (int)ftype, B2D_FIR_PROG> block_t;

public:
  ac_fir_prog_coeffs() : have_(false) {}

  // Exactly ONE queued sample per call, with the coefficient array of THIS call (reference :277-303); the taps may
  // differ from call to call while the delay line persists.
  void run(ac_channel<IN_TYPE> &data_in, ac_channel<OUT_TYPE> &data_out, const COEFF_TYPE coeffs[N_TAPS]) {
    if (!data_in.available(1)) return;
    sync_taps(coeffs);
    blk.process(data_in, data_out, 1);
  }

  // extension: every queued sample with one coefficient array (== calling run() until data_in is empty)
  void run_block(ac_channel<IN_TYPE> &data_in, ac_channel<OUT_TYPE> &data_out, const COEFF_TYPE coeffs[N_TAPS]) {
    if (!data_in.available(1)) return;
    sync_taps(coeffs);
    blk.process(data_in, data_out);
  }

private:
  void sync_taps(const COEFF_TYPE *coeffs) {
    bool same = have_;
    for (int i = 0; i < N_TAPS; i++) {
      const typename block_t::coeff_raw_t r = (typename block_t::coeff_raw_t)b200dsp::fixed_traits<COEFF_TYPE>::to_raw(coeffs[i]);
      if (!same || r != taps_[i]) { same = false; taps_[i] = r; }
    }
    if (!same) blk.load_raw(taps_);
    have_ = true;
  }

  block_t blk;
  typename block_t::coeff_raw_t taps_[N_TAPS];
  bool have_;
};

#endif
