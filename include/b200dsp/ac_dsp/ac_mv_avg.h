// b200dsp facade: ac_mv_avg on the B200 engine.
//
// Drop-in for hlslibs/ac_dsp include/ac_dsp/ac_mv_avg.h:140-204 -- same class name, template parameters, public cff_ptr
// member and run() signature.  The include guard is the reference's.  ac_window_mode (AC_WIN / AC_CLIP / AC_MIRROR) comes
// from <ac_window.h> of the AC Math package, exactly as in the reference header.
// PARITY UNPINNED: the window's boundary semantics are restated from the manual (see include/b200dsp.h).
#ifndef _INCLUDED_AC_MV_AVG_H_
#define _INCLUDED_AC_MV_AVG_H_

#include <ac_window.h>

#include "../marshal.h"

template <int MAX_SAMPLE, int TAPS, ac_window_mode WIN_TYPE, class IN_TYPE, class OUT_TYPE, class ACC_TYPE, class COEFF_TYPE, class S_TYPE>
class ac_mv_avg {
  static_assert(TAPS >= 1 && (TAPS & 1), "b200dsp: the window span must be odd");
  static_assert(IN_TYPE::width <= 32 && COEFF_TYPE::width <= 32 && ACC_TYPE::width <= 64 && OUT_TYPE::width <= 64,
                "b200dsp: type wider than the engine holds");
  typedef typename b200dsp::container_sel<IN_TYPE::width>::type in_raw_t;
  typedef typename b200dsp::container_sel<OUT_TYPE::width>::type out_raw_t;
  typedef typename b200dsp::container_sel<COEFF_TYPE::width>::type coeff_raw_t;

public:
  // public in the reference too, "so that the user can extract the coeffs array" (:144-147)
  const COEFF_TYPE *const cff_ptr;

  ac_mv_avg(const COEFF_TYPE *const c_ptr) : cff_ptr(c_ptr), h_(0) {}
  ~ac_mv_avg() { if (h_) b2d_mvavg_destroy(h_); }

  // Every n_sample token queued is read and the LAST one counts (:165-174); then whole bursts of that many samples are
  // consumed while data is queued (:178-195).  All bursts of a call go to the GPU in one launch.
  void run(ac_channel<IN_TYPE> &data_in, ac_channel<OUT_TYPE> &data_out, ac_channel<S_TYPE> &n_sample) {
    bool have = false;
    unsigned long long ns = 0;
    while (n_sample.available(1)) { ns = (unsigned long long)n_sample.read().to_uint64(); have = true; }
    if (!data_in.available(1)) return;
    if (!have) throw b200dsp::engine_error(B2D_ESTATE, "ac_mv_avg::run: no n_sample token queued (the reference reads an unset value)");
    b200dsp::drain(data_in, in_);
    if (!h_) {
      // the constructor's pointer is read at the first run(), like the constant-coefficient FIR facade (the wrapper idiom
      // initialises the derived class's array after the base class)
      coeff_raw_t raw[TAPS];
      for (int i = 0; i < TAPS; i++) raw[i] = (coeff_raw_t)b200dsp::fixed_traits<COEFF_TYPE>::to_raw(cff_ptr[i]);
      b2d_mvavg_desc d;
      d.in = b200dsp::fixed_traits<IN_TYPE>::fmt(); d.out = b200dsp::fixed_traits<OUT_TYPE>::fmt();
      d.acc = b200dsp::fixed_traits<ACC_TYPE>::fmt(); d.coeff = b200dsp::fixed_traits<COEFF_TYPE>::fmt();
      d.max_sample = MAX_SAMPLE; d.taps = TAPS;
      d.win_type = WIN_TYPE == AC_WIN ? B2D_WIN : (WIN_TYPE == AC_CLIP ? B2D_CLIP : B2D_MIRROR);
      d.device = -1;
      b200dsp::check(b2d_mvavg_create(&h_, &d, raw), "b2d_mvavg_create");
    }
    out_.resize(in_.size());
    size_t n_out = 0;
    b200dsp::check(b2d_mvavg_run(h_, in_.data(), in_.size(), (size_t)ns, out_.data(), &n_out), "b2d_mvavg_run");
    b200dsp::emit(data_out, out_.data(), n_out);
  }

private:
  ac_mv_avg(const ac_mv_avg &);
  ac_mv_avg &operator=(const ac_mv_avg &);
  b2d_mvavg *h_;
  std::vector<in_raw_t> in_;
  std::vector<out_raw_t> out_;
};

#endif
