// b200dsp facade: ac_fir_const_coeffs on the B200 engine.
//
// Drop-in for hlslibs/ac_dsp include/ac_dsp/ac_fir_const_coeffs.h:309-355 -- same class name, template parameters,
// constructor and run() signature; put this directory in front of the ac_dsp include path and link libb200dsp.
// The include guard is the reference's, so a later #include of the original header is a no-op.
#ifndef _INCLUDED_AC_FIR_CONST_COEFFS_H_
#define _INCLUDED_AC_FIR_CONST_COEFFS_H_

#include "../fir_block.h"

template <class IN_TYPE, class OUT_TYPE, class COEFF_TYPE, class ACC_TYPE, unsigned N_TAPS, FTYPE ftype>
class ac_fir_const_coeffs {
public:
  // The array behind c_ptr usually belongs to a class DERIVED from this one (the reference's wrapper idiom,
  // tests/rtest_ac_fir_const_coeffs.cpp:94-108), so it is not initialised yet while this constructor runs:
  // like the reference, only keep the pointer here and read through it when run() needs the taps.
  ac_fir_const_coeffs(const COEFF_TYPE *const c_ptr) : coeffs(c_ptr) {}

  // Drains data_in, appends one output per input to data_out (reference :321-355).
  void run(ac_channel<IN_TYPE> &data_in, ac_channel<OUT_TYPE> &data_out) {
    if (!data_in.available(1)) return;
    if (!blk.loaded()) blk.load(coeffs);
    blk.process(data_in, data_out);
  }

  // extension: the same call on raw arrays (n samples in the int16/int32/int64 container of IN_TYPE / OUT_TYPE)
  void run_raw(const typename b200dsp::container_sel<IN_TYPE::width>::type *in, size_t n,
               typename b200dsp::container_sel<OUT_TYPE::width>::type *out) {
    if (!blk.loaded()) blk.load(coeffs);
    blk.process_raw(in, n, out);
  }

private:
  const COEFF_TYPE *const coeffs;
  b200dsp::fir_block<IN_TYPE, OUT_TYPE, COEFF_TYPE, ACC_TYPE, N_TAPS, (int)ftype, B2D_FIR_CONST> blk;
};

#endif
