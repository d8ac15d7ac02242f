// b200dsp facade: ac_cic_dec_full on the B200 engine.
//
// Drop-in for hlslibs/ac_dsp include/ac_dsp/ac_cic_dec_full.h:147-222 -- same class name, template parameters and
// run() signature.  The include guard is the reference's, so a later #include of the original header is a no-op.
// The lossless internal type (find_inter_type_cic_dec, :116-137) is derived inside the engine (b2d_cic_int_width).
#ifndef _INCLUDED_AC_CIC_DEC_FULL_H_
#define _INCLUDED_AC_CIC_DEC_FULL_H_

#include "../cic_block.h"

template <class IN_TYPE, class OUT_TYPE, unsigned R_, unsigned M_, unsigned N_>
class ac_cic_dec_full {
public:
  ac_cic_dec_full() {}

  // Integrates every queued input and emits the comb output for inputs 0, R, 2R, ... of the stream (:163-166).
  // State (integrators, phase, comb delays) persists across calls like the reference object's members.
  void run(ac_channel<IN_TYPE> &data_in, ac_channel<OUT_TYPE> &data_out) { blk.process(data_in, data_out); }

  // extension: raw arrays; returns the number of outputs (<= n / R + 1)
  size_t run_raw(const typename b200dsp::container_sel<IN_TYPE::width>::type *in, size_t n,
                 typename b200dsp::container_sel<OUT_TYPE::width>::type *out) { return blk.process_raw(in, n, out); }

private:
  b200dsp::cic_block<IN_TYPE, OUT_TYPE, R_, M_, N_, B2D_CIC_DEC> blk;
};

#endif
