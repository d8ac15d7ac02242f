// common.cuh -- fixed-point format helpers shared by the host runtime and the CUDA kernels.
//
// The arithmetic rules restated here are the published AC Datatypes (hlslibs/ac_types) ones that
// the reference's `acc += a*b`, `data_out = acc` and INT_TYPE -> OUT_TYPE assignments rely on:
// drop fraction bits with the target's quantisation mode, then integer bits with its overflow mode.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>
#include "../../include/b200dsp.h"

namespace b2d {

struct Fmt {
  int W, I, S, Q, O;
  __host__ __device__ int F() const { return W - I; }
};

typedef __int128 i128;
typedef unsigned __int128 u128;

// Keep the low W bits (1 <= W <= 64) and sign-/zero-extend.
__host__ __device__ __forceinline__ int64_t wrap_bits(int64_t v, int W, int S) {
  if (W >= 64) return v;
  const int sh = 64 - W;
  return S ? ((int64_t)((uint64_t)v << sh) >> sh) : (int64_t)(((uint64_t)v << sh) >> sh);
}

// The three primitives below exist for two intermediate widths: W = i128 (every format the engine accepts) and
// W = int64_t (the same arithmetic when the caller has checked that every intermediate fits 62 bits: 1.8 x faster on fir63 on
// the GPU, where 128-bit shifts and multiplies are long instruction sequences).

// floor(v / 2^sh) corrected for quantisation mode Q; 0 < sh < bits(W) - 1.
template <class W>
__host__ __device__ __forceinline__ W quantize_t(W v, int sh, int Q) {
  W q = v >> sh;
  if (Q == B2D_TRN) return q;
  const W rem = v - (q << sh);
  const bool msb = (bool)((rem >> (sh - 1)) & 1);
  if (Q == B2D_RND) return q + (msb ? 1 : 0);
  const bool rest = (rem & ((((W)1) << (sh - 1)) - 1)) != 0;
  const bool neg = v < 0;
  bool up = false;
  switch (Q) {
    case B2D_TRN_ZERO: up = neg && rem != 0; break;
    case B2D_RND_INF: up = msb && (rest || !neg); break;
    case B2D_RND_ZERO: up = msb && (rest || neg); break;
    case B2D_RND_MIN_INF: up = msb && rest; break;
    case B2D_RND_CONV: up = msb && (rest || (q & 1)); break;
    case B2D_RND_CONV_ODD: up = msb && (rest || !(q & 1)); break;
    default: break;
  }
  return q + (up ? 1 : 0);
}

// Assignment of a value with F2 fraction bits to format f (any Q, any O); result fits 64 bits.
template <class W>
__host__ __device__ __forceinline__ int64_t convert_t(W v, int F2, const Fmt &f) {
  const int F = f.F();
  if (F2 > F) v = quantize_t<W>(v, F2 - F, f.Q);
  else if (F > F2) v = v << (F - F2);
  if (f.O == B2D_WRAP) return wrap_bits((int64_t)v, f.W, f.S);
  // bounds of the format; for W = int64_t the caller guarantees f.W <= 62 whenever a saturating mode is in play
  const W hi = f.S ? ((((W)1) << (f.W - 1)) - 1) : ((((W)1) << f.W) - 1);
  const W lo = f.S ? -(((W)1) << (f.W - 1)) : 0;
  if (f.O == B2D_SAT) return (int64_t)(v > hi ? hi : (v < lo ? lo : v));
  if (f.O == B2D_SAT_ZERO) return (int64_t)((v > hi || v < lo) ? 0 : v);
  const W slo = f.S ? -hi : 0;  // B2D_SAT_SYM
  return (int64_t)(v > hi ? hi : (v < slo ? slo : v));
}

// Full `acc += p` of an ACC_TYPE accumulator (any Q / O): exact sum at max(F), then assignment.
template <class W>
__host__ __device__ __forceinline__ int64_t macc_t(int64_t acc, const Fmt &fa, W p, int Fp) {
  const int Fa = fa.F();
  const int rF = Fa > Fp ? Fa : Fp;
  const W s = (W)acc * ((W)1 << (rF - Fa)) + p * ((W)1 << (rF - Fp));   // shifts of negative values, spelled as products
  return convert_t<W>(s, rF, fa);
}

__host__ __device__ __forceinline__ i128 quantize(i128 v, int sh, int Q) { return quantize_t<i128>(v, sh, Q); }
__host__ __device__ __forceinline__ int64_t convert(i128 v, int F2, const Fmt &f) { return convert_t<i128>(v, F2, f); }

// One tap of `acc += p` for an accumulator with Q in {TRN, RND}, O = WRAP, tracked modulo 2^64:
// returns the value to add to the running (unwrapped) raw accumulator.  s = Fp - Facc.
__host__ __device__ __forceinline__ int64_t tap_term(i128 p, int s, int Q) {
  if (s > 0) return (int64_t)quantize(p, s, Q);
  return (int64_t)((u128)p << (-s));
}

__host__ __device__ __forceinline__ int64_t macc(int64_t acc, const Fmt &fa, i128 p, int Fp) { return macc_t<i128>(acc, fa, p, Fp); }

// Can `acc += p` (p: a product of Wp bits incl. sign with Fp fraction bits) and the final assignment to `out` be evaluated
// with 64-bit intermediates?  Every shifted operand must stay within 62 bits so that their sum and the rounding
// increments cannot overflow.
__host__ __device__ __forceinline__ bool fits_i64(const Fmt &acc, const Fmt &out, int Wp, int Fp) {
  const int Fa = acc.F(), rF = Fa > Fp ? Fa : Fp;
  const int Wa = acc.W + (acc.S ? 0 : 1);
  if (Wa + (rF - Fa) > 62 || Wp + (rF - Fp) > 62) return false;
  if (Wa + (out.F() > Fa ? out.F() - Fa : 0) > 62) return false;
  if (out.W + (out.S ? 0 : 1) > 62 && out.O != B2D_WRAP) return false;
  return true;
}

__host__ __device__ __forceinline__ int container_bytes(int W) { return W <= 16 ? 2 : (W <= 32 ? 4 : 8); }

// Raw load/store in the 2/4/8-byte container of a format.
__device__ __forceinline__ int64_t load_raw(const void *p, size_t idx, int bytes, int S) {
  if (bytes == 2) return S ? (int64_t)((const int16_t *)p)[idx] : (int64_t)((const uint16_t *)p)[idx];
  if (bytes == 4) return S ? (int64_t)((const int32_t *)p)[idx] : (int64_t)((const uint32_t *)p)[idx];
  return ((const int64_t *)p)[idx];
}
__device__ __forceinline__ void store_raw(void *p, size_t idx, int bytes, int64_t v) {
  if (bytes == 2) ((int16_t *)p)[idx] = (int16_t)v;
  else if (bytes == 4) ((int32_t *)p)[idx] = (int32_t)v;
  else ((int64_t *)p)[idx] = v;
}

// Element index of (time i, channel c) in a multi-channel buffer holding n samples per channel.
__host__ __device__ __forceinline__ size_t elem_index(size_t i, uint32_t c, size_t n, uint32_t C, int interleaved) {
  return interleaved ? i * C + c : (size_t)c * n + i;
}

}  // namespace b2d
