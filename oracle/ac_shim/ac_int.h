// oracle/ac_shim/ac_int.h -- TEST INFRASTRUCTURE, not product code.
// Clean-room stand-in for the small part of AC Datatypes `ac_int` used by the ac_dsp
// FIR / CIC headers: counters and ring-buffer pointers (construct from integers, ++,
// use as an array index, compare / add with plain ints).  Arithmetic is done by
// converting to `long long`; assignment wraps to W bits.  See ac_fixed.h in this
// directory for why this shim exists.
#ifndef B200DSP_ORACLE_AC_SHIM_AC_INT_H
#define B200DSP_ORACLE_AC_SHIM_AC_INT_H

enum ac_special_val { AC_VAL_DC, AC_VAL_0, AC_VAL_MIN, AC_VAL_MAX, AC_VAL_QUANTUM };

template <int W, bool S = true>
class ac_int {
public:
  static const int width = W;
  static const bool sign = S;
  long long v;

  ac_int() : v(0) {}
  ac_int(long long x) { set(x); }
  ac_int(int x) { set(x); }
  ac_int(unsigned x) { set((long long)x); }
  ac_int(long x) { set(x); }
  ac_int(unsigned long x) { set((long long)x); }
  ac_int(bool x) { set(x ? 1 : 0); }
  template <int W2, bool S2> ac_int(const ac_int<W2, S2> &o) { set(o.v); }

  operator long long() const { return v; }
  ac_int &operator++() { set(v + 1); return *this; }
  ac_int operator++(int) { ac_int t = *this; set(v + 1); return t; }
  ac_int &operator--() { set(v - 1); return *this; }
  ac_int operator--(int) { ac_int t = *this; set(v - 1); return t; }
  ac_int &operator+=(long long x) { set(v + x); return *this; }
  ac_int &operator-=(long long x) { set(v - x); return *this; }
  int to_int() const { return (int)v; }
  long long to_int64() const { return v; }
  unsigned long long to_uint64() const { return (unsigned long long)v; }

private:
  void set(long long x) {
    if (W <= 0) { v = 0; return; }
    if (W >= 64) { v = x; return; }
    unsigned long long m = (1ULL << (W >= 64 ? 0 : W)) - 1ULL;
    unsigned long long u = ((unsigned long long)x) & m;
    if (S && ((u >> (W > 0 ? W - 1 : 0)) & 1ULL)) u |= ~m;
    v = (long long)u;
  }
};

namespace ac {
template <unsigned long long N> struct log2_floor { enum { val = 1 + log2_floor<(N >> 1)>::val }; };
template <> struct log2_floor<1> { enum { val = 0 }; };
template <> struct log2_floor<0> { enum { val = 0 }; };
// smallest k with 2^k >= N
template <unsigned long long N> struct log2_ceil { enum { val = (N <= 1) ? 0 : (log2_floor<(N <= 1 ? 1 : N - 1)>::val + 1) }; };

template <ac_special_val V, int W, bool S>
inline bool init_array(ac_int<W, S> *a, int n) {
  for (int i = 0; i < n; i++) a[i] = 0;
  return true;
}
}  // namespace ac

#endif
