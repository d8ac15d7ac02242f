// b200dsp/marshal.h -- glue between AC Datatypes objects and the raw-integer C-ABI (include/b200dsp.h).
//
// The header facade in this directory re-creates the five hot-path class templates of hlslibs/ac_dsp
// (same names, template parameters and run() signatures) on top of the CUDA engine.  This file holds what
// they share: ac_fixed<W,I,S,Q,O> -> b2d_fmt, ac_fixed <-> raw two's-complement integer, ac_channel <-> array,
// and the error policy.
//
// Only the PUBLIC AC Datatypes API is used, so the facade works with the real hlslibs/ac_types package or any
// compatible implementation:
//     T::width, T::i_width, T::sign, T::q_mode, T::o_mode      (static members of ac_fixed)
//     x.template slc<W>(0).to_int64()                           (raw bits out)
//     x.set_slc(0, ac_int<W,S>(raw))                            (raw bits in)
//     ac_channel<T>::available(n) / read() / write(v)
//
// Error policy.  The reference's run() returns void and cannot fail; configurations it cannot build are compile
// errors.  Here: widths the engine cannot hold are static_asserts; every run-time engine failure (no CUDA device,
// configuration without a bit-exact CUDA implementation, out of memory) throws b200dsp::engine_error -- there is no
// CPU fallback to fall back to.
#ifndef B200DSP_MARSHAL_H
#define B200DSP_MARSHAL_H

#include <cstdint>
#include <stdexcept>
#include <string>
#include <vector>

#include <ac_channel.h>
#include <ac_fixed.h>
#include <ac_int.h>

#include "../b200dsp.h"

namespace b200dsp {

class engine_error : public std::runtime_error {
public:
  engine_error(int status, const std::string &what) : std::runtime_error(what), status_(status) {}
  int status() const { return status_; }
private:
  int status_;
};

inline void check(int status, const char *where) {
  if (status != B2D_OK)
    throw engine_error(status, std::string(where) + ": " + b2d_strerror(status) + " (" + b2d_last_error() + ")");
}

template <class T>
struct fixed_traits {
  enum { W = T::width, I = T::i_width, S = T::sign ? 1 : 0 };
  static b2d_fmt fmt() {
    b2d_fmt f;
    f.W = W; f.I = I; f.S = S; f.Q = (int32_t)T::q_mode; f.O = (int32_t)T::o_mode;
    return f;
  }
  static int64_t to_raw(const T &x) { return (int64_t)x.template slc<W>(0).to_int64(); }
  static T from_raw(int64_t raw) {
    T t;
    t.set_slc(0, ac_int<W, T::sign>(raw));
    return t;
  }
};

// Raw buffer in the container the C-ABI expects for a W-bit format: int16 / int32 / int64.
template <int BITS> struct container_bits;
template <> struct container_bits<16> { typedef int16_t type; };
template <> struct container_bits<32> { typedef int32_t type; };
template <> struct container_bits<64> { typedef int64_t type; };
template <int W> struct container_sel {
  static_assert(W >= 1 && W <= 64, "b200dsp: ac_fixed widths above 64 bits cannot cross the engine boundary");
  typedef typename container_bits<(W <= 16 ? 16 : (W <= 32 ? 32 : 64))>::type type;
};

// Drain up to `limit` queued values of a channel into a raw array (limit = 0: everything available).
template <class T>
inline void drain(ac_channel<T> &ch, std::vector<typename container_sel<T::width>::type> &raw, size_t limit = 0) {
  typedef typename container_sel<T::width>::type C;
  raw.clear();
  while (ch.available(1) && (limit == 0 || raw.size() < limit)) raw.push_back((C)fixed_traits<T>::to_raw(ch.read()));
}

template <class T>
inline void emit(ac_channel<T> &ch, const typename container_sel<T::width>::type *raw, size_t n) {
  for (size_t i = 0; i < n; i++) ch.write(fixed_traits<T>::from_raw((int64_t)raw[i]));
}

}  // namespace b200dsp

#endif
