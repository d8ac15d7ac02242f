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
// Storage is tiered like the real package's (which keeps ceil(W/32) native words): the raw value of a
// type that needs <= 32 bits lives in an int32_t, <= 64 bits in an int64_t, anything wider in an __int128
// (every width the hot path can produce, <= 32+64+1 bits, is exact).  Each operator computes in the storage
// type of its RESULT and each conversion in the narrowest tier that holds the re-scaled source plus one
// carry bit, so the common <16,1> x <16,1> -> <40,8> path is plain 32/64-bit machine arithmetic and the CPU
// timings taken with this shim are not handicapped by 128-bit emulation.  `ac_shim::from_raw / to_raw` are
// shim-only helpers used by the oracle driver to move raw integers in and out.
#ifndef B200DSP_ORACLE_AC_SHIM_AC_FIXED_H
#define B200DSP_ORACLE_AC_SHIM_AC_FIXED_H

#include <cmath>
#include <stdint.h>
#include <iostream>
#include "ac_int.h"

enum ac_q_mode { AC_TRN, AC_RND, AC_TRN_ZERO, AC_RND_ZERO, AC_RND_INF, AC_RND_MIN_INF, AC_RND_CONV, AC_RND_CONV_ODD };
enum ac_o_mode { AC_WRAP, AC_SAT, AC_SAT_ZERO, AC_SAT_SYM };

namespace ac_shim {
typedef __int128 wide_t;

template <bool C, class A, class B> struct tsel { typedef A type; };
template <class A, class B> struct tsel<false, A, B> { typedef B type; };
// narrowest signed container with at least BITS bits
template <int BITS> struct store {
  typedef typename tsel<(BITS <= 32), int32_t, typename tsel<(BITS <= 64), int64_t, wide_t>::type>::type type;
};
template <class T> struct uns;
template <> struct uns<int32_t> { typedef uint32_t type; };
template <> struct uns<int64_t> { typedef uint64_t type; };
template <> struct uns<wide_t> { typedef unsigned __int128 type; };

template <class T>
inline T wrap_bits(T v, int W, bool S) {
  typedef typename uns<T>::type U;
  const int TB = (int)sizeof(T) * 8;
  if (W >= TB) return v;
  if (S) return (T)((T)((U)v << (TB - W)) >> (TB - W));   // shift pair = sign-extend from bit W-1 (as the real package normalises)
  return (T)(((U)v) & ((((U)1) << W) - 1));
}
template <class T> inline T max_val(int W, bool S) { return S ? (T)((((T)1) << (W - 1)) - 1) : (T)((((T)1) << W) - 1); }
template <class T> inline T min_val(int W, bool S) { return S ? (T)(-(((T)1) << (W - 1))) : (T)0; }

// q = floor(v / 2^sh) corrected per quantisation mode; 0 < sh < bits(T) - 1.
template <class T>
inline T quantize(T v, int sh, ac_q_mode Q) {
  T q = v >> sh;                            // arithmetic shift = floor
  if (Q == AC_TRN) return q;
  T rem = v - (T)((typename uns<T>::type)q << sh);   // 0 <= rem < 2^sh
  bool msb = (rem >> (sh - 1)) & 1;
  bool rest = (rem & ((((T)1) << (sh - 1)) - 1)) != 0;
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
// v must fit T with at least one bit to spare above W (+1 if unsigned): see conv_bits.
template <class T>
inline T overflow(T v, int W, bool S, ac_o_mode O) {
  if (O == AC_WRAP) return wrap_bits<T>(v, W, S);
  T hi = max_val<T>(W, S), lo = min_val<T>(W, S);
  switch (O) {
    case AC_WRAP: break;
    case AC_SAT:  return v > hi ? hi : (v < lo ? lo : v);
    case AC_SAT_ZERO: return (v > hi || v < lo) ? (T)0 : v;
    case AC_SAT_SYM: {
      T slo = S ? (T)-hi : (T)0;
      return v > hi ? hi : (v < slo ? slo : v);
    }
  }
  return v;
}
// Re-scale raw value with F2 fraction bits into <W, F> with modes Q, O, computing in T.
template <class T>
inline T convert(T v, int F2, int W, int F, bool S, ac_q_mode Q, ac_o_mode O) {
  if (F2 > F) v = quantize<T>(v, F2 - F, Q);
  else if (F > F2) v = (T)((typename uns<T>::type)v << (F - F2));
  return overflow<T>(v, W, S, O);
}
template <bool C, int A, int B> struct sel { enum { v = A }; };
template <int A, int B> struct sel<false, A, B> { enum { v = B }; };
template <int A, int B> struct imax { enum { v = (A > B) ? A : B }; };
// bits the computation of a conversion needs: the source (SB bits as a signed number) moved to the destination's
// binary point plus a carry / rounding bit, the destination's range (DB bits) plus one for the saturation
// compares, and a right-shift count that stays below the container width
template <int SB, int F2, int DB, int F> struct conv_bits {
  enum { up = (F > F2) ? (F - F2) : 0, dn = (F2 > F) ? (F2 - F) : 0,
         v = imax<imax<SB + up + 1, DB + 1>::v, dn + 2>::v };
};
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

  enum { BITS = W + (S ? 0 : 1) };   // bits of the raw value read as a signed number
  typedef typename ac_shim::store<BITS>::type raw_t;
  raw_t v;  // canonical raw value (sign- or zero-extended from W bits)

  ac_fixed() : v(0) {}
  template <int W2, int I2, bool S2, ac_q_mode Q2, ac_o_mode O2>
  ac_fixed(const ac_fixed<W2, I2, S2, Q2, O2> &o) {
    typedef typename ac_shim::store<ac_shim::conv_bits<W2 + (S2 ? 0 : 1), W2 - I2, BITS, F>::v>::type CT;
    v = (raw_t)ac_shim::convert<CT>((CT)o.v, W2 - I2, W, F, S, Q, O);
  }
  template <int W2, bool S2>
  ac_fixed(const ac_int<W2, S2> &o) { set_int((ac_shim::wide_t)o.v); }
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
    if (V == AC_VAL_MAX) v = (raw_t)ac_shim::max_val<ac_shim::wide_t>(W, S);
    else if (V == AC_VAL_MIN) v = (raw_t)ac_shim::min_val<ac_shim::wide_t>(W, S);
    else if (V == AC_VAL_QUANTUM) v = 1;
    else v = 0;  // AC_VAL_0; AC_VAL_DC ("don't care") is given a defined value here
    return *this;
  }

  double to_double() const { return (double)std::ldexp((long double)v, -F); }
  long double to_long_double() const { return std::ldexp((long double)v, -F); }
  int to_int() const { return (int)(F >= 0 ? ((ac_shim::wide_t)v >> F) : ((ac_shim::wide_t)v << -F)); }

  // bit slices of the raw value (WS <= 64): read as ac_int<WS,S>, write from any ac_int
  template <int WS>
  ac_int<WS, S> slc(int lsb) const { return ac_int<WS, S>((long long)((ac_shim::wide_t)v >> lsb)); }
  template <int W2, bool S2>
  ac_fixed &set_slc(int lsb, const ac_int<W2, S2> &s) {
    const unsigned __int128 m = (W2 >= 128 ? ~(unsigned __int128)0 : ((((unsigned __int128)1) << W2) - 1)) << lsb;
    const unsigned __int128 u = (((unsigned __int128)(ac_shim::wide_t)v) & ~m) | ((((unsigned __int128)(ac_shim::wide_t)s.v) << lsb) & m);
    v = (raw_t)ac_shim::wrap_bits<ac_shim::wide_t>((ac_shim::wide_t)u, W, S);
    return *this;
  }

  // --- arithmetic (result types follow the AC Datatypes width rules) ---
  template <int W2, int I2, bool S2, ac_q_mode Q2, ac_o_mode O2>
  ac_fixed<W + W2, I + I2, S || S2> operator*(const ac_fixed<W2, I2, S2, Q2, O2> &o) const {
    typedef ac_fixed<W + W2, I + I2, S || S2> R;
    R r;
    r.v = (typename R::raw_t)v * (typename R::raw_t)o.v;
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
    typedef typename rt<W2, I2, S2>::plus R;
    typedef typename ac_shim::uns<typename R::raw_t>::type U;
    R r;
    const int rF = rt<W2, I2, S2>::pF;
    r.v = (typename R::raw_t)(((U)(typename R::raw_t)v << (rF - F)) + ((U)(typename R::raw_t)o.v << (rF - (W2 - I2))));
    return r;
  }
  template <int W2, int I2, bool S2, ac_q_mode Q2, ac_o_mode O2>
  typename rt<W2, I2, S2>::minus operator-(const ac_fixed<W2, I2, S2, Q2, O2> &o) const {
    typedef typename rt<W2, I2, S2>::minus R;
    typedef typename ac_shim::uns<typename R::raw_t>::type U;
    R r;
    const int rF = rt<W2, I2, S2>::pF;
    r.v = (typename R::raw_t)(((U)(typename R::raw_t)v << (rF - F)) - ((U)(typename R::raw_t)o.v << (rF - (W2 - I2))));
    return r;
  }
  template <int W2, int I2, bool S2, ac_q_mode Q2, ac_o_mode O2>
  ac_fixed &operator+=(const ac_fixed<W2, I2, S2, Q2, O2> &o) { *this = this->operator+(o); return *this; }
  template <int W2, int I2, bool S2, ac_q_mode Q2, ac_o_mode O2>
  ac_fixed &operator-=(const ac_fixed<W2, I2, S2, Q2, O2> &o) { *this = this->operator-(o); return *this; }

  // unary minus: one more integer bit, always signed (AC Datatypes rt_unary::neg); exact
  ac_fixed<W + 1, I + 1, true> operator-() const {
    ac_fixed<W + 1, I + 1, true> r;
    r.v = -(typename ac_fixed<W + 1, I + 1, true>::raw_t)v;
    return r;
  }
  // shifts move the bit pattern inside the same type: bits shifted out are lost, the binary point stays
  ac_fixed operator>>(int n) const {
    ac_fixed r;
    r.v = n >= 0 ? ac_shim::wrap_bits<raw_t>(v >> n, W, S)
                 : ac_shim::wrap_bits<raw_t>((raw_t)((typename ac_shim::uns<raw_t>::type)v << -n), W, S);
    return r;
  }
  ac_fixed operator<<(int n) const { return this->operator>>(-n); }

  template <int W2, int I2, bool S2, ac_q_mode Q2, ac_o_mode O2>
  bool operator==(const ac_fixed<W2, I2, S2, Q2, O2> &o) const {
    const int rF = ac_shim::imax<F, W2 - I2>::v;
    return ((ac_shim::wide_t)v << (rF - F)) == ((ac_shim::wide_t)o.v << (rF - (W2 - I2)));
  }
  template <int W2, int I2, bool S2, ac_q_mode Q2, ac_o_mode O2>
  bool operator!=(const ac_fixed<W2, I2, S2, Q2, O2> &o) const { return !(*this == o); }

private:
  void set_int(ac_shim::wide_t b) { v = (raw_t)ac_shim::convert<ac_shim::wide_t>(b, 0, W, F, S, Q, O); }
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
    v = (raw_t)ac_shim::overflow<ac_shim::wide_t>(q, W, S, O);
  }
};

template <int W, int I, bool S, ac_q_mode Q, ac_o_mode O>
inline std::ostream &operator<<(std::ostream &os, const ac_fixed<W, I, S, Q, O> &x) {
  os << x.to_double();
  return os;
}

namespace ac_shim {
template <class T> inline T from_raw(long long raw) { T t; t.v = (typename T::raw_t)wrap_bits<wide_t>((wide_t)raw, T::width, T::sign); return t; }
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
