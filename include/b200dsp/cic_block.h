// b200dsp/cic_block.h -- one engine CIC handle behind an ac_dsp-style CIC object (shared by the two CIC facades).
#ifndef B200DSP_CIC_BLOCK_H
#define B200DSP_CIC_BLOCK_H

#include "marshal.h"

namespace b200dsp {

template <class IN_TYPE, class OUT_TYPE, unsigned R, unsigned M, unsigned N, int MODE>
class cic_block {
  static_assert(IN_TYPE::width <= 32, "b200dsp: CIC IN_TYPE wider than 32 bits");
  static_assert(OUT_TYPE::width <= 64, "b200dsp: CIC OUT_TYPE wider than 64 bits");
  // 8-bit rate counters in the reference (ac_cic_full_core.h:72-73,91); R = 1 never re-reads in the interpolator
  static_assert(R >= 2 && R <= 256 && M >= 1 && N >= 1 && N <= 255, "b200dsp: CIC R / M / N out of range");

public:
  typedef typename container_sel<IN_TYPE::width>::type in_raw_t;
  typedef typename container_sel<OUT_TYPE::width>::type out_raw_t;

  cic_block() : h_(0) {}
  ~cic_block() { if (h_) b2d_cic_destroy(h_); }

  void process(ac_channel<IN_TYPE> &data_in, ac_channel<OUT_TYPE> &data_out) {
    drain(data_in, in_);
    if (in_.empty()) return;
    create();
    out_.resize(b2d_cic_max_out(h_, in_.size()) + 1);
    size_t n_out = 0;
    check(b2d_cic_run(h_, in_.data(), in_.size(), out_.data(), &n_out), "b2d_cic_run");
    emit(data_out, out_.data(), n_out);
  }

  // array form: returns the number of outputs written (capacity b2d_cic_max_out(handle(), n))
  size_t process_raw(const in_raw_t *in, size_t n, out_raw_t *out) {
    create();
    size_t n_out = 0;
    check(b2d_cic_run(h_, in, n, out, &n_out), "b2d_cic_run");
    return n_out;
  }

  b2d_cic *handle() { create(); return h_; }

private:
  cic_block(const cic_block &);
  cic_block &operator=(const cic_block &);

  void create() {
    if (h_) return;
    b2d_cic_desc d;
    d.in = fixed_traits<IN_TYPE>::fmt(); d.out = fixed_traits<OUT_TYPE>::fmt();
    d.R = R; d.M = M; d.N = N; d.mode = MODE; d.n_channels = 1; d.layout = B2D_PLANAR; d.device = -1;
    check(b2d_cic_create(&h_, &d), "b2d_cic_create");
  }

  b2d_cic *h_;
  std::vector<in_raw_t> in_;
  std::vector<out_raw_t> out_;
};

}  // namespace b200dsp

#endif
