// oracle/ac_shim/ac_fixed.h -- TEST INFRASTRUCTURE, not product code.
//
// Clean-room stand-in for the subset of the AC Datatypes `ac_fixed` class that the
// ac_dsp FIR / CIC headers and their test benches use.  The real package
// (hlslibs/ac_types, unpinned) is not vendored by the
// reference and is absent from this image, so the unmodified reference headers are
// compiled against this file instead (-I oracle/ac_shim -I /root/reference/include).
//
// Semantics implemented (published AC Datatypes behaviour):
//   * ac_fixed<W,I,S,Q,O>: W-bit two's-complement raw value, value = raw * 2^-(W-I)
//   * a*b  -> ac_fixed<W1+W2, I1+I2, S1||S2>               (exact)
//   * a+b  -> I = max(I1+(S2&&!S1), I2+(S1&&!S2))+1, F = max(F1,F2), S = S1||S2 (exact)
//   * a-b  -> same widths as a+b but always signed          (exact)
//   * -a   -> ac_fixed<W+1, I+1, true> (exact);  a >> n / a << n keep the type of a (bit-pattern shift)
//   * conversion/assignment: drop fraction bits with quantisation mode Q, then drop
//     integer bits with overflow mode O.  a += b  ==  a = a + b.
// Raw values live in an __int128, so every width the hot path can produce
// (<= 32+64+1 bits) is exact.  `ac_shim::from_raw / to_raw` are shim-only helpers used
// by the oracle driver to move raw integers in and out.
#ifndef B200DSP_ORACLE_AC_SHIM_AC_FIXED_H
#define B200DSP_ORACLE_AC_SHIM_AC_FIXED_H

#include <cmath>
#include <iostream>
#include "ac_int.h"

enum ac_q_mode { AC_TRN, AC_RND, AC_TRN_ZERO, AC_RND_ZERO, AC_RND_INF, AC_RND_MIN_INF, AC_RND_CONV, AC_RND_CONV_ODD };
enum ac_o_mode { AC_WRAP, AC_SAT, AC_SAT_ZERO, AC_SAT_SYM };

namespace ac_shim {
typedef __int128 wide_t;

inline wide_t wrap_bits(wide_t v, int W, bool S) {
  if (W >= 128) return v;
  unsigned __int128 m = (((unsigned __int128)1) << W) - 1;
  unsigned __int128 u = ((unsigned __int128)v) & m;
  if (S && ((u >> (W - 1)) & 1)) u |= ~m;
  return (wide_t)u;
}
inline wide_t max_val(int W, bool S) { return S ? ((((wide_t)1) << (W - 1)) - 1) : ((((wide_t)1) << W) - 1); }
inline wide_t min_val(int W, bool S) { return S ? -(((wide_t)1) << (W - 1)) : 0; }

// q = floor(v / 2^sh) corrected per quantisation mode; sh > 0.
inline wide_t quantize(wide_t v, int sh, ac_q_mode Q) {
  wide_t q = v >> sh;                       // arithmetic shift = floor
  wide_t rem = v - (q << sh);               // 0 <= rem < 2^sh
  bool msb = (rem >> (sh - 1)) & 1;
  bool rest = (rem & ((((wide_t)1) << (sh - 1)) - 1)) != 0;
  bool neg = v < 0;
  switch (Q) {
    case AC_TRN:          break;
    case AC_RND:          q += msb; break;
    case AC_TRN_ZERO:     q += (neg && rem != 0); break;
    case AC_RND_INF:      q += (msb && (rest || !neg)); break;
    case AC_RND_ZERO:     q += (msb && (rest || neg)); break;
    case AC_RND_MIN_INF:  q += (msb && rest); break;
    case AC_RND_CONV:     q += (msb && (rest || (q & 1))); break;
    case AC_RND_CONV_ODD: q += (msb && (rest || !(q & 1))); break;
  }
  return q;
}
inline wide_t overflow(wide_t v, int W, bool S, ac_o_mode O) {
  wide_t hi = max_val(W, S), lo = min_val(W, S);
  switch (O) {
    case AC_WRAP: return wrap_bits(v, W, S);
    case AC_SAT:  return v > hi ? hi : (v < lo ? lo : v);
    case AC_SAT_ZERO: return (v > hi || v < lo) ? 0 : v;
    case AC_SAT_SYM: {
      wide_t slo = S ? -hi : 0;
      return v > hi ? hi : (v < slo ? slo : v);
    }
  }
  return v;
}
// Re-scale raw value with F2 fraction bits into <W, F> with modes Q, O.
inline wide_t convert(wide_t v, int F2, int W, int F, bool S, ac_q_mode Q, ac_o_mode O) {
  if (F2 > F) v = quantize(v, F2 - F, Q);
  else if (F > F2) v = v << (F - F2);
  return overflow(v, W, S, O);
}
template <bool C, int A, int B> struct sel { enum { v = A }; };
template <int A, int B> struct sel<false, A, B> { enum { v = B }; };
template <int A, int B> struct imax { enum { v = (A > B) ? A : B }; };
}  // namespace ac_shim

template <int W, int I, bool S = true, ac_q_mode Q = AC_TRN, ac_o_mode O = AC_WRAP>
class ac_fixed {
public:
  static const int width = W;
  static const int i_width = I;
  static const bool sign = S;
  static const ac_q_mode q_mode = Q;
  static const ac_o_mode o_mode = O;
  enum { F = W - I };

  ac_shim::wide_t v;  // canonical raw value (sign- or zero-extended from W bits)

  ac_fixed() : v(0) {}
  template <int W2, int I2, bool S2, ac_q_mode Q2, ac_o_mode O2>
  ac_fixed(const ac_fixed<W2, I2, S2, Q2, O2> &o) { v = ac_shim::convert(o.v, W2 - I2, W, F, S, Q, O); }
  template <int W2, bool S2>
  ac_fixed(const ac_int<W2, S2> &o) { v = ac_shim::convert((ac_shim::wide_t)o.v, 0, W, F, S, Q, O); }
  ac_fixed(bool b) { set_int(b ? 1 : 0); }
  ac_fixed(char b) { set_int(b); }
  ac_fixed(short b) { set_int(b); }
  ac_fixed(int b) { set_int(b); }
  ac_fixed(unsigned b) { set_int(b); }
  ac_fixed(long b) { set_int(b); }
  ac_fixed(unsigned long b) { set_int((ac_shim::wide_t)b); }
  ac_fixed(long long b) { set_int(b); }
  ac_fixed(unsigned long long b) { set_int((ac_shim::wide_t)b); }
  ac_fixed(double d) { set_real((long double)d); }
  ac_fixed(float d) { set_real((long double)d); }
  ac_fixed(long double d) { set_real(d); }

  template <ac_special_val V>
  ac_fixed &set_val() {
    if (V == AC_VAL_MAX) v = ac_shim::max_val(W, S);
    else if (V == AC_VAL_MIN) v = ac_shim::min_val(W, S);
    else if (V == AC_VAL_QUANTUM) v = 1;
    else v = 0;  // AC_VAL_0; AC_VAL_DC ("don't care") is given a defined value here
    return *this;
  }

  double to_double() const { return (double)std::ldexp((long double)v, -F); }
  long double to_long_double() const { return std::ldexp((long double)v, -F); }
  int to_int() const { return (int)(F >= 0 ? (v >> F) : (v << -F)); }

  // bit slices of the raw value (WS <= 64): read as ac_int<WS,S>, write from any ac_int
  template <int WS>
  ac_int<WS, S> slc(int lsb) const { return ac_int<WS, S>((long long)(v >> lsb)); }
  template <int W2, bool S2>
  ac_fixed &set_slc(int lsb, const ac_int<W2, S2> &s) {
    const unsigned __int128 m = (W2 >= 128 ? ~(unsigned __int128)0 : ((((unsigned __int128)1) << W2) - 1)) << lsb;
    const unsigned __int128 u = (((unsigned __int128)v) & ~m) | ((((unsigned __int128)(ac_shim::wide_t)s.v) << lsb) & m);
    v = ac_shim::wrap_bits((ac_shim::wide_t)u, W, S);
    return *this;
  }

  // --- arithmetic (result types follow the AC Datatypes width rules) ---
  template <int W2, int I2, bool S2, ac_q_mode Q2, ac_o_mode O2>
  ac_fixed<W + W2, I + I2, S || S2> operator*(const ac_fixed<W2, I2, S2, Q2, O2> &o) const {
    ac_fixed<W + W2, I + I2, S || S2> r;
    r.v = v * o.v;
    return r;
  }
  template <int W2, int I2, bool S2>
  struct rt {
    enum {
      F2 = W2 - I2,
      pI = ac_shim::imax<I + (S2 && !S), I2 + (S && !S2)>::v + 1,
      pF = ac_shim::imax<F, F2>::v,
      pW = pI + pF
    };
    typedef ac_fixed<pW, pI, S || S2> plus;
    typedef ac_fixed<pW, pI, true> minus;
  };
  template <int W2, int I2, bool S2, ac_q_mode Q2, ac_o_mode O2>
  typename rt<W2, I2, S2>::plus operator+(const ac_fixed<W2, I2, S2, Q2, O2> &o) const {
    typename rt<W2, I2, S2>::plus r;
    const int rF = rt<W2, I2, S2>::pF;
    r.v = (v << (rF - F)) + (o.v << (rF - (W2 - I2)));
    return r;
  }
  template <int W2, int I2, bool S2, ac_q_mode Q2, ac_o_mode O2>
  typename rt<W2, I2, S2>::minus operator-(const ac_fixed<W2, I2, S2, Q2, O2> &o) const {
    typename rt<W2, I2, S2>::minus r;
    const int rF = rt<W2, I2, S2>::pF;
    r.v = (v << (rF - F)) - (o.v << (rF - (W2 - I2)));
    return r;
  }
  template <int W2, int I2, bool S2, ac_q_mode Q2, ac_o_mode O2>
  ac_fixed &operator+=(const ac_fixed<W2, I2, S2, Q2, O2> &o) { *this = this->operator+(o); return *this; }
  template <int W2, int I2, bool S2, ac_q_mode Q2, ac_o_mode O2>
  ac_fixed &operator-=(const ac_fixed<W2, I2, S2, Q2, O2> &o) { *this = this->operator-(o); return *this; }

  // unary minus: one more integer bit, always signed (AC Datatypes rt_unary::neg); exact
  ac_fixed<W + 1, I + 1, true> operator-() const {
    ac_fixed<W + 1, I + 1, true> r;
    r.v = -v;
    return r;
  }
  // shifts move the bit pattern inside the same type: bits shifted out are lost, the binary point stays
  ac_fixed operator>>(int n) const {
    ac_fixed r;
    r.v = n >= 0 ? ac_shim::wrap_bits(v >> n, W, S) : ac_shim::wrap_bits(v << -n, W, S);
    return r;
  }
  ac_fixed operator<<(int n) const { return this->operator>>(-n); }

  template <int W2, int I2, bool S2, ac_q_mode Q2, ac_o_mode O2>
  bool operator==(const ac_fixed<W2, I2, S2, Q2, O2> &o) const {
    const int rF = ac_shim::imax<F, W2 - I2>::v;
    return (v << (rF - F)) == (o.v << (rF - (W2 - I2)));
  }
  template <int W2, int I2, bool S2, ac_q_mode Q2, ac_o_mode O2>
  bool operator!=(const ac_fixed<W2, I2, S2, Q2, O2> &o) const { return !(*this == o); }

private:
  void set_int(ac_shim::wide_t b) { v = ac_shim::convert(b, 0, W, F, S, Q, O); }
  void set_real(long double d) {
    // Scale to the LSB, split into floor + remainder, apply Q on the remainder, then O.
    long double sc = std::ldexp(d, F);
    long double fl = std::floor(sc);
    long double fr = sc - fl;                 // in [0,1)
    ac_shim::wide_t q = (ac_shim::wide_t)fl;
    bool msb = fr >= 0.5L, rest = (fr != 0.0L && fr != 0.5L), neg = sc < 0, any = fr != 0.0L;
    switch (Q) {
      case AC_TRN: break;
      case AC_RND: q += msb; break;
      case AC_TRN_ZERO: q += (neg && any); break;
      case AC_RND_INF: q += (msb && (rest || !neg)); break;
      case AC_RND_ZERO: q += (msb && (rest || neg)); break;
      case AC_RND_MIN_INF: q += (msb && rest); break;
      case AC_RND_CONV: q += (msb && (rest || (q & 1))); break;
      case AC_RND_CONV_ODD: q += (msb && (rest || !(q & 1))); break;
    }
    v = ac_shim::overflow(q, W, S, O);
  }
};

template <int W, int I, bool S, ac_q_mode Q, ac_o_mode O>
inline std::ostream &operator<<(std::ostream &os, const ac_fixed<W, I, S, Q, O> &x) {
  os << x.to_double();
  return os;
}

namespace ac_shim {
template <class T> inline T from_raw(long long raw) { T t; t.v = wrap_bits((wide_t)raw, T::width, T::sign); return t; }
template <class T> inline long long to_raw(const T &t) { return (long long)t.v; }
}  // namespace ac_shim

namespace ac {
template <ac_special_val V, int W, int I, bool S, ac_q_mode Q, ac_o_mode O>
inline bool init_array(ac_fixed<W, I, S, Q, O> *a, int n) {
  ac_fixed<W, I, S, Q, O> t;
  t.template set_val<V>();
  for (int i = 0; i < n; i++) a[i] = t;
  return true;
}
}  // namespace ac

#endif
