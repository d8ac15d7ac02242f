// b200dsp facade: ac_fir_reg_share on the B200 engine.
//
// Drop-in for hlslibs/ac_dsp include/ac_dsp/ac_fir_reg_share.h:257-307 -- same class name, template parameters
// (with their defaults), constructor, scalar run() and ac_firProgCoeffs_delay_line().  The include guard is the
// reference's.  The delay line is CALLER-OWNED memory in the reference (several objects may point into one register
// array), so this class keeps it on the host exactly as the reference does -- shift, insert the new sample -- and hands
// the engine the resulting window for the tap-MAC (b2d_fir_run_window).  One CUDA launch per sample: this is the
// compatibility path; streams belong on b2d_fir_run with kind = B2D_FIR_REG_SHARE (run_block below).
#ifndef _INCLUDED_AC_FIR_REG_SHARE_H_
#define _INCLUDED_AC_FIR_REG_SHARE_H_

#include "../fir_block.h"

typedef ac_fixed<16, 1, true> DEFAULT_TYPE;   // reference ac_fir_reg_share.h:92

template <int N_TAPS = 2, class IN_TYPE = DEFAULT_TYPE, class OUT_TYPE = DEFAULT_TYPE, class COEFF_TYPE = DEFAULT_TYPE,
          class ACC_TYPE = DEFAULT_TYPE, int MEM_WORD_WIDTH = 1, int BLK_SZ = 1, int BLK_OFFSET = 0, FTYPE ftype = SHIFT_REG>
class ac_fir_reg_share {
  static_assert(ftype == SHIFT_REG || ftype == FOLD_EVEN || ftype == FOLD_ODD || ftype == FOLD_EVEN_ANTI || ftype == FOLD_ODD_ANTI,
                "b200dsp: ac_fir_reg_share dispatches SHIFT_REG, FOLD_EVEN(_ANTI) and FOLD_ODD(_ANTI) only (its run() writes an unset value otherwise)");
  enum {
    USED = (ftype == SHIFT_REG) ? N_TAPS : ((ftype == FOLD_EVEN || ftype == FOLD_EVEN_ANTI) ? N_TAPS / 2 : (N_TAPS - 1) / 2 + 1),
    RAM = (USED / BLK_SZ - 1) * MEM_WORD_WIDTH + BLK_OFFSET + BLK_SZ   // coefficient RAM words the tap loop touches
  };
  static_assert(USED % BLK_SZ == 0, "b200dsp: the tap loop must be a whole number of BLK_SZ blocks (the reference indexes out of range otherwise)");
  typedef typename b200dsp::container_sel<IN_TYPE::width>::type in_raw_t;
  typedef typename b200dsp::container_sel<OUT_TYPE::width>::type out_raw_t;
  typedef typename b200dsp::container_sel<COEFF_TYPE::width>::type coeff_raw_t;

public:
  ac_fir_reg_share(IN_TYPE *ptr_t) : ptr(ptr_t), h_(0), have_(false) {}
  ~ac_fir_reg_share() { if (h_) b2d_fir_destroy(h_); }

  // One sample: shift it into the caller's delay line, filter with the coefficient RAM of THIS call (:277-303).
  void run(IN_TYPE &data_in, COEFF_TYPE coeffs[N_TAPS], OUT_TYPE &data_out) {
    for (int i = N_TAPS - 1; i >= 0; i--) ptr[i] = (i == 0) ? data_in : ptr[i - 1];   // firShiftReg (:105-111)
    sync_taps(coeffs);
    in_raw_t win[N_TAPS];
    for (int i = 0; i < N_TAPS; i++) win[i] = (in_raw_t)b200dsp::fixed_traits<IN_TYPE>::to_raw(ptr[i]);
    out_raw_t y = 0;
    b200dsp::check(b2d_fir_run_window(h_, win, &y), "b2d_fir_run_window");
    data_out = b200dsp::fixed_traits<OUT_TYPE>::from_raw((int64_t)y);
  }

  // The sample leaving the delay line (:304-307): a plain type conversion of caller-owned data.
  void ac_firProgCoeffs_delay_line(OUT_TYPE &core_out) { core_out = ptr[N_TAPS - 1]; }

  // extension: n samples in one launch on the engine's own copy of the delay line (which starts as all zeros and is
  // NOT the caller's array); the caller's array is brought up to date afterwards.
  void run_block(const in_raw_t *in, size_t n, COEFF_TYPE coeffs[N_TAPS], out_raw_t *out) {
    sync_taps(coeffs);
    size_t n_out = 0;
    b200dsp::check(b2d_fir_run(h_, in, n, out, &n_out), "b2d_fir_run");
    for (size_t k = 0; k < n; k++) {
      for (int i = N_TAPS - 1; i >= 1; i--) ptr[i] = ptr[i - 1];
      ptr[0] = b200dsp::fixed_traits<IN_TYPE>::from_raw((int64_t)in[k]);
    }
  }

private:
  ac_fir_reg_share(const ac_fir_reg_share &);
  ac_fir_reg_share &operator=(const ac_fir_reg_share &);

  void sync_taps(const COEFF_TYPE *coeffs) {
    if (!h_) {
      b2d_fir_desc d;
      d.in = b200dsp::fixed_traits<IN_TYPE>::fmt(); d.coeff = b200dsp::fixed_traits<COEFF_TYPE>::fmt();
      d.acc = b200dsp::fixed_traits<ACC_TYPE>::fmt(); d.out = b200dsp::fixed_traits<OUT_TYPE>::fmt();
      d.n_taps = N_TAPS; d.ftype = (int)ftype; d.kind = B2D_FIR_REG_SHARE; d.n_channels = 1; d.layout = B2D_PLANAR; d.device = -1;
      b200dsp::check(b2d_fir_create(&h_, &d), "b2d_fir_create");
    }
    bool same = have_;
    for (int i = 0; i < RAM; i++) {
      const coeff_raw_t r = (coeff_raw_t)b200dsp::fixed_traits<COEFF_TYPE>::to_raw(coeffs[i]);
      if (!same || r != ram_[i]) { same = false; ram_[i] = r; }
    }
    if (!same) b200dsp::check(b2d_fir_load_blocked(h_, ram_, RAM, MEM_WORD_WIDTH, BLK_SZ, BLK_OFFSET, -1), "b2d_fir_load_blocked");
    have_ = true;
  }

  IN_TYPE *ptr;
  b2d_fir *h_;
  coeff_raw_t ram_[RAM];
  bool have_;
};

#endif
