// tests/cpp/marshal_roundtrip.cpp -- the facade's value marshaling (include/b200dsp/marshal.h: ac_fixed <-> raw
// two's-complement container through the public AC-datatypes API only) against the shim's canonical raw value,
// for every storage tier of the shim (<= 32, <= 64 bits) and both signednesses.  No engine call: runs on CPU.
#include <ac_fixed.h>
#include <ac_int.h>
#include <ac_channel.h>
#include <marshal.h>
#include <cstdio>
#include <cstdlib>

template <class T>
static int check() {
  int bad = 0;
  for (int i = 0; i < 100000; i++) {
    long long r = ((long long)rand() << 33) ^ ((long long)rand() << 11) ^ rand();
    if (i < 4) r = (i & 1) ? -1 - (i >> 1) : (i >> 1);
    const T a = ac_shim::from_raw<T>(r);
    const long long canon = ac_shim::to_raw(a);
    const int64_t o = b200dsp::fixed_traits<T>::to_raw(a);
    const T b = b200dsp::fixed_traits<T>::from_raw(o), c = b200dsp::fixed_traits<T>::from_raw(r);
    bad += (o != canon) + (ac_shim::to_raw(b) != canon) + (ac_shim::to_raw(c) != canon);
  }
  if (bad) std::printf("marshal mismatch: W=%d I=%d S=%d: %d\n", T::width, T::i_width, (int)T::sign, bad);
  return bad;
}

int main() {
  srand(20260101);
  int bad = 0;
  bad += check<ac_fixed<16, 1, true> >();
  bad += check<ac_fixed<20, 5, true> >();
  bad += check<ac_fixed<31, 5, false> >();
  bad += check<ac_fixed<32, 3, false> >();
  bad += check<ac_fixed<33, 5, true> >();
  bad += check<ac_fixed<40, 8, true> >();
  bad += check<ac_fixed<63, 5, false> >();
  bad += check<ac_fixed<64, 3, true> >();
  std::printf("bad=%d\n", bad);
  return bad != 0;
}
