// b200dsp facade: ac_poly_dec on the B200 engine.
//
// Drop-in for hlslibs/ac_dsp include/ac_dsp/ac_poly_dec.h:87-137 -- same class name, template parameters and run()
// signature (the coefficients arrive as one struct with a coeffs[NTAPS * DF] member on a channel).  The include guard
// is the reference's.
#ifndef _INCLUDED_AC_POLY_DEC_H_
#define _INCLUDED_AC_POLY_DEC_H_

#include "../marshal.h"

template <class IN_TYPE, class COEFF_TYPE, class STR_COEFF_TYPE, class ACC_TYPE, class OUT_TYPE, int NTAPS, int DF>
class ac_poly_dec {
  static_assert(NTAPS >= 1 && DF >= 1, "b200dsp: NTAPS and DF must be positive");
  static_assert(IN_TYPE::width <= 32 && COEFF_TYPE::width <= 32, "b200dsp: IN_TYPE / COEFF_TYPE wider than 32 bits");
  static_assert(ACC_TYPE::width <= 64 && OUT_TYPE::width <= 64, "b200dsp: ACC_TYPE / OUT_TYPE wider than 64 bits");
  typedef typename b200dsp::container_sel<IN_TYPE::width>::type in_raw_t;
  typedef typename b200dsp::container_sel<OUT_TYPE::width>::type out_raw_t;
  typedef typename b200dsp::container_sel<COEFF_TYPE::width>::type coeff_raw_t;

public:
  ac_poly_dec() : h_(0) {}
  ~ac_poly_dec() { if (h_) b2d_polydec_destroy(h_); }

  // Every queued coefficient struct is read, the last one wins (:101-106); then whole groups of DF samples are consumed,
  // one output per group, and an incomplete group stays queued on data_in (:107-109).
  void run(ac_channel<IN_TYPE> &data_in, ac_channel<OUT_TYPE> &data_out, ac_channel<STR_COEFF_TYPE> &coeffs_st) {
    bool fresh = false;
    STR_COEFF_TYPE coeffs_t;
    while (coeffs_st.available(1)) { coeffs_t = coeffs_st.read(); fresh = true; }
    if (fresh) {
      create();
      coeff_raw_t raw[NTAPS * DF];
      for (int i = 0; i < NTAPS * DF; i++) raw[i] = (coeff_raw_t)b200dsp::fixed_traits<COEFF_TYPE>::to_raw(coeffs_t.coeffs[i]);
      b200dsp::check(b2d_polydec_load(h_, raw, NTAPS * DF, -1), "b2d_polydec_load");
    }
    in_.clear();
    while (data_in.available(DF))
      for (int i = 0; i < DF; i++) in_.push_back((in_raw_t)b200dsp::fixed_traits<IN_TYPE>::to_raw(data_in.read()));
    if (in_.empty()) return;
    create();
    out_.resize(in_.size() / DF + 1);
    size_t n_out = 0;
    b200dsp::check(b2d_polydec_run(h_, in_.data(), in_.size(), out_.data(), &n_out), "b2d_polydec_run");
    b200dsp::emit(data_out, out_.data(), n_out);
  }

private:
  ac_poly_dec(const ac_poly_dec &);
  ac_poly_dec &operator=(const ac_poly_dec &);
  void create() {
    if (h_) return;
    b2d_polydec_desc d;
    d.in = b200dsp::fixed_traits<IN_TYPE>::fmt(); d.coeff = b200dsp::fixed_traits<COEFF_TYPE>::fmt();
    d.acc = b200dsp::fixed_traits<ACC_TYPE>::fmt(); d.out = b200dsp::fixed_traits<OUT_TYPE>::fmt();
    d.n_taps = NTAPS; d.df = DF; d.n_channels = 1; d.layout = B2D_PLANAR; d.device = -1;
    b200dsp::check(b2d_polydec_create(&h_, &d), "b2d_polydec_create");
  }
  b2d_polydec *h_;
  std::vector<in_raw_t> in_;
  std::vector<out_raw_t> out_;
};

#endif
