// b200dsp facade: ac_intg_dump on the B200 engine.
//
// Drop-in for hlslibs/ac_dsp include/ac_dsp/ac_intg_dump.h:113-151 -- same class name, template parameters and run()
// signature.  The include guard is the reference's.
#ifndef _INCLUDED_AC_INTG_DUMP_H_
#define _INCLUDED_AC_INTG_DUMP_H_

#include "../marshal.h"

template <class IN_TYPE, class ACC_TYPE, class OUT_TYPE, class N_TYPE, int NS, int CHN>
class ac_intg_dump {
  static_assert(NS >= 1 && CHN >= 1, "b200dsp: NS and CHN must be positive");
  static_assert(IN_TYPE::width <= 32 && ACC_TYPE::width <= 64 && OUT_TYPE::width <= 64, "b200dsp: type wider than the engine holds");
  typedef typename b200dsp::container_sel<IN_TYPE::width>::type in_raw_t;
  typedef typename b200dsp::container_sel<OUT_TYPE::width>::type out_raw_t;

public:
  ac_intg_dump() : h_(0) {}
  ~ac_intg_dump() { if (h_) b2d_intgdump_destroy(h_); }

  // While samples are queued: read one n_sample token, consume that frame (n_sample samples per channel and a dump of
  // CHN sums, or NS samples per channel and no dump when the token is outside 1 .. NS), :133-147.  All frames of a
  // call go to the GPU in one launch.
  void run(ac_channel<IN_TYPE> &data_in, ac_channel<OUT_TYPE> &data_out, ac_channel<N_TYPE> &n_sample) {
    in_.clear();
    tok_.clear();
    while (data_in.available(1)) {
      const unsigned long long n = (unsigned long long)n_sample.read().to_uint64();
      const size_t take = (size_t)((n >= 1 && n <= (unsigned long long)NS) ? n : (unsigned long long)NS) * CHN;
      tok_.push_back(n > 0xFFFFFFFFull ? 0xFFFFFFFFu : (uint32_t)n);
      for (size_t i = 0; i < take; i++) in_.push_back((in_raw_t)b200dsp::fixed_traits<IN_TYPE>::to_raw(data_in.read()));
    }
    if (tok_.empty()) return;
    if (!h_) {
      b2d_intgdump_desc d;
      d.in = b200dsp::fixed_traits<IN_TYPE>::fmt(); d.acc = b200dsp::fixed_traits<ACC_TYPE>::fmt(); d.out = b200dsp::fixed_traits<OUT_TYPE>::fmt();
      d.ns = NS; d.chn = CHN; d.device = -1;
      b200dsp::check(b2d_intgdump_create(&h_, &d), "b2d_intgdump_create");
    }
    out_.resize(tok_.size() * CHN);
    size_t n_out = 0;
    b200dsp::check(b2d_intgdump_run(h_, in_.data(), in_.size(), tok_.data(), tok_.size(), out_.data(), &n_out), "b2d_intgdump_run");
    b200dsp::emit(data_out, out_.data(), n_out);
  }

private:
  ac_intg_dump(const ac_intg_dump &);
  ac_intg_dump &operator=(const ac_intg_dump &);
  b2d_intgdump *h_;
  std::vector<in_raw_t> in_;
  std::vector<uint32_t> tok_;
  std::vector<out_raw_t> out_;
};

#endif
