// b200dsp/fir_block.h -- one engine FIR handle behind an ac_dsp-style filter object (shared by the three FIR facades).
#ifndef B200DSP_FIR_BLOCK_H
#define B200DSP_FIR_BLOCK_H

#include "marshal.h"

// Same guard and enumerators as the reference (ac_fir_const_coeffs.h:93-97, ac_fir_load_coeffs.h:103-107,
// ac_fir_prog_coeffs.h:76-80), so facade and reference headers can meet in one translation unit.
#ifndef __FIR_FILTER_TYPES_ENUM_DEF__
#define __FIR_FILTER_TYPES_ENUM_DEF__
typedef enum { SHIFT_REG, ROTATE_SHIFT, C_BUFF, FOLD_EVEN, FOLD_ODD, TRANSPOSED, FOLD_EVEN_ANTI, FOLD_ODD_ANTI } FTYPE;
#endif

namespace b200dsp {

template <class IN_TYPE, class OUT_TYPE, class COEFF_TYPE, class ACC_TYPE, unsigned N_TAPS, int ftype, int KIND>
class fir_block {
  static_assert(N_TAPS >= 1, "b200dsp: N_TAPS must be at least 1");
  static_assert(IN_TYPE::width <= 32 && COEFF_TYPE::width <= 32, "b200dsp: IN_TYPE / COEFF_TYPE wider than 32 bits");
  static_assert(ACC_TYPE::width <= 64 && OUT_TYPE::width <= 64, "b200dsp: ACC_TYPE / OUT_TYPE wider than 64 bits");
  static_assert(ftype != (int)FOLD_EVEN_ANTI && ftype != (int)FOLD_ODD_ANTI,
                "b200dsp: the reference classes do not dispatch the _ANTI architectures (their run() writes an unset value)");

public:
  typedef typename container_sel<IN_TYPE::width>::type in_raw_t;
  typedef typename container_sel<OUT_TYPE::width>::type out_raw_t;
  typedef typename container_sel<COEFF_TYPE::width>::type coeff_raw_t;

  fir_block() : h_(0) {}
  ~fir_block() { if (h_) b2d_fir_destroy(h_); }

  bool loaded() const { return loaded_; }

  // N_TAPS coefficients, in the order of the reference's coeffs[] array
  void load(const COEFF_TYPE *c) {
    coeff_raw_t raw[N_TAPS];
    for (unsigned i = 0; i < N_TAPS; i++) raw[i] = (coeff_raw_t)fixed_traits<COEFF_TYPE>::to_raw(c[i]);
    load_raw(raw);
  }
  void load_raw(const coeff_raw_t *raw) {
    create();
    check(b2d_fir_load(h_, raw, N_TAPS, -1), "b2d_fir_load");
    loaded_ = true;
  }

  // the run() sample loop over everything queued (limit = 0) or at most `limit` samples
  void process(ac_channel<IN_TYPE> &data_in, ac_channel<OUT_TYPE> &data_out, size_t limit = 0) {
    drain(data_in, in_, limit);
    if (in_.empty()) return;
    create();
    out_.resize(in_.size());
    size_t n_out = 0;
    check(b2d_fir_run(h_, in_.data(), in_.size(), out_.data(), &n_out), "b2d_fir_run");
    emit(data_out, out_.data(), n_out);
  }

  // array form of the same call (no ac_channel marshaling): n raw samples in, n raw samples out
  void process_raw(const in_raw_t *in, size_t n, out_raw_t *out) {
    create();
    size_t n_out = 0;
    check(b2d_fir_run(h_, in, n, out, &n_out), "b2d_fir_run");
  }

  b2d_fir *handle() { create(); return h_; }

private:
  fir_block(const fir_block &);
  fir_block &operator=(const fir_block &);

  void create() {
    if (h_) return;
    b2d_fir_desc d;
    d.in = fixed_traits<IN_TYPE>::fmt(); d.coeff = fixed_traits<COEFF_TYPE>::fmt();
    d.acc = fixed_traits<ACC_TYPE>::fmt(); d.out = fixed_traits<OUT_TYPE>::fmt();
    d.n_taps = N_TAPS; d.ftype = ftype; d.kind = KIND; d.n_channels = 1; d.layout = B2D_PLANAR; d.device = -1;
    check(b2d_fir_create(&h_, &d), "b2d_fir_create");
  }

  b2d_fir *h_;
  bool loaded_ = false;
  std::vector<in_raw_t> in_;
  std::vector<out_raw_t> out_;
};

}  // namespace b200dsp

#endif
