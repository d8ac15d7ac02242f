// b200dsp facade: ac_cic_intr_full on the B200 engine.
//
// Drop-in for hlslibs/ac_dsp include/ac_dsp/ac_cic_intr_full.h:137-215 -- same class name, template parameters and
// run() signature.  The include guard is the reference's, so a later #include of the original header is a no-op.
#ifndef _INCLUDED_AC_CIC_INTR_FULL_H_
#define _INCLUDED_AC_CIC_INTR_FULL_H_

#include "../cic_block.h"

template <class IN_TYPE, class OUT_TYPE, unsigned R_, unsigned M_, unsigned N_>
class ac_cic_intr_full {
public:
  ac_cic_intr_full() {}

  // Comb at the input rate, zero-stuff by R, integrate at the output rate, drop the first N-1 values (:150-153).
  // Stream-edge rule of the reference: after K inputs in total, max(0, (K-1)R + 1 - (N-1)) outputs exist; the
  // remaining R-1 outputs of the last input appear once the next input arrives (on a later run()).
  void run(ac_channel<IN_TYPE> &data_in, ac_channel<OUT_TYPE> &data_out) { blk.process(data_in, data_out); }

  // extension: raw arrays; returns the number of outputs (<= n * R)
  size_t run_raw(const typename b200dsp::container_sel<IN_TYPE::width>::type *in, size_t n,
                 typename b200dsp::container_sel<OUT_TYPE::width>::type *out) { return blk.process_raw(in, n, out); }

private:
  b200dsp::cic_block<IN_TYPE, OUT_TYPE, R_, M_, N_, B2D_CIC_INTR> blk;
};

#endif
