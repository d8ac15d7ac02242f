// b200dsp facade: ac_poly_intr on the B200 engine.
//
// Drop-in for hlslibs/ac_dsp include/ac_dsp/ac_poly_intr.h:261-312 -- same class name, template parameters, polyphase
// FTYPE enum (with the reference's guard macro) and run() signature.  One run() call consumes one read_ctrl token:
// true loads the control struct {bool sign[IF]; ac_int<8,false> corr[IF];} and the coefficient struct
// {COEFF_TYPE coeffs[COEFFSZ];} (:286-288), false consumes one sample and writes IF outputs (the folded forms one
// step late, :160).  The include guard is the reference's.
#ifndef _INCLUDED_AC_POLY_INTR_H_
#define _INCLUDED_AC_POLY_INTR_H_

#include "../marshal.h"

#ifndef __POLY_FILTER_TYPES_ENUM_DEF__
#define __POLY_FILTER_TYPES_ENUM_DEF__
typedef enum { FOLD_EVEN, FOLD_ODD, FOLD_ANTI } FTYPE;
#endif

template <class IN_TYPE, class COEFF_TYPE, class ACC_TYPE, class OUT_TYPE, class STR_CTRL_TYPE, class STR_COEFF_TYPE, int NTAPS, int COEFFSZ, int IF, FTYPE ftype>
class ac_poly_intr {
  static_assert(NTAPS >= 1 && IF >= 1 && IF <= 255, "b200dsp: NTAPS must be positive and IF in 1..255");
  static_assert(IN_TYPE::width <= 32 && COEFF_TYPE::width <= 32, "b200dsp: IN_TYPE / COEFF_TYPE wider than 32 bits");
  static_assert(ACC_TYPE::width <= 64 && OUT_TYPE::width <= 64, "b200dsp: ACC_TYPE / OUT_TYPE wider than 64 bits");
  static const int USED = IF * (ftype == FOLD_EVEN ? NTAPS / 2 : (ftype == FOLD_ODD ? NTAPS / 2 + 1 : NTAPS));
  static_assert(COEFFSZ >= USED, "b200dsp: COEFFSZ smaller than the coefficients this architecture reads");
  typedef typename b200dsp::container_sel<IN_TYPE::width>::type in_raw_t;
  typedef typename b200dsp::container_sel<OUT_TYPE::width>::type out_raw_t;
  typedef typename b200dsp::container_sel<COEFF_TYPE::width>::type coeff_raw_t;

public:
  ac_poly_intr() : h_(0) {}
  ~ac_poly_intr() { if (h_) b2d_polyintr_destroy(h_); }

  void run(ac_channel<IN_TYPE> &data_in, ac_channel<OUT_TYPE> &data_out, ac_channel<STR_CTRL_TYPE> &ctrl_st,
           ac_channel<STR_COEFF_TYPE> &coeffs_st, ac_channel<bool> &read_ctrl_chan) {
    const bool read_ctrl = read_ctrl_chan.read();
    create();
    if (read_ctrl) {
      const STR_CTRL_TYPE ctrl_t = ctrl_st.read();
      const STR_COEFF_TYPE coeffs_t = coeffs_st.read();
      coeff_raw_t raw[USED > 0 ? USED : 1];
      unsigned char sign[IF], corr[IF];
      for (int i = 0; i < USED; i++) raw[i] = (coeff_raw_t)b200dsp::fixed_traits<COEFF_TYPE>::to_raw(coeffs_t.coeffs[i]);
      for (int j = 0; j < IF; j++) { sign[j] = ctrl_t.sign[j] ? 1 : 0; corr[j] = (unsigned char)(long long)ctrl_t.corr[j]; }
      b200dsp::check(b2d_polyintr_load(h_, raw, USED, sign, corr, -1), "b2d_polyintr_load");
      return;
    }
    const in_raw_t x = (in_raw_t)b200dsp::fixed_traits<IN_TYPE>::to_raw(data_in.read());
    out_raw_t y[IF];
    size_t n_out = 0;
    b200dsp::check(b2d_polyintr_run(h_, &x, 1, y, &n_out), "b2d_polyintr_run");
    b200dsp::emit(data_out, y, n_out);
  }

private:
  ac_poly_intr(const ac_poly_intr &);
  ac_poly_intr &operator=(const ac_poly_intr &);
  void create() {
    if (h_) return;
    b2d_polyintr_desc d;
    d.in = b200dsp::fixed_traits<IN_TYPE>::fmt(); d.coeff = b200dsp::fixed_traits<COEFF_TYPE>::fmt();
    d.acc = b200dsp::fixed_traits<ACC_TYPE>::fmt(); d.out = b200dsp::fixed_traits<OUT_TYPE>::fmt();
    d.n_taps = NTAPS; d.intr_factor = IF;
    d.ftype = ftype == FOLD_EVEN ? B2D_PI_FOLD_EVEN : (ftype == FOLD_ODD ? B2D_PI_FOLD_ODD : B2D_PI_FOLD_ANTI);
    d.n_channels = 1; d.layout = B2D_PLANAR; d.device = -1;
    b200dsp::check(b2d_polyintr_create(&h_, &d), "b2d_polyintr_create");
  }
  b2d_polyintr *h_;
};

#endif
