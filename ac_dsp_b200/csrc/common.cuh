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

// floor(v / 2^sh) corrected for quantisation mode Q; 0 < sh < 127.
__host__ __device__ __forceinline__ i128 quantize(i128 v, int sh, int Q) {
  i128 q = v >> sh;
  if (Q == B2D_TRN) return q;
  const i128 rem = v - (q << sh);
  const bool msb = (bool)((rem >> (sh - 1)) & 1);
  if (Q == B2D_RND) return q + (msb ? 1 : 0);
  const bool rest = (rem & ((((i128)1) << (sh - 1)) - 1)) != 0;
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
__host__ __device__ __forceinline__ int64_t convert(i128 v, int F2, const Fmt &f) {
  const int F = f.F();
  if (F2 > F) v = quantize(v, F2 - F, f.Q);
  else if (F > F2) v = v << (F - F2);
  if (f.O == B2D_WRAP) return wrap_bits((int64_t)v, f.W, f.S);
  const i128 hi = f.S ? ((((i128)1) << (f.W - 1)) - 1) : ((((i128)1) << f.W) - 1);
  const i128 lo = f.S ? -(((i128)1) << (f.W - 1)) : 0;
  if (f.O == B2D_SAT) return (int64_t)(v > hi ? hi : (v < lo ? lo : v));
  if (f.O == B2D_SAT_ZERO) return (int64_t)((v > hi || v < lo) ? 0 : v);
  const i128 slo = f.S ? -hi : 0;  // B2D_SAT_SYM
  return (int64_t)(v > hi ? hi : (v < slo ? slo : v));
}

// One tap of `acc += p` for an accumulator with Q in {TRN, RND}, O = WRAP, tracked modulo 2^64:
// returns the value to add to the running (unwrapped) raw accumulator.  s = Fp - Facc.
__host__ __device__ __forceinline__ int64_t tap_term(i128 p, int s, int Q) {
  if (s > 0) return (int64_t)quantize(p, s, Q);
  return (int64_t)((u128)p << (-s));
}

// Full `acc += p` of an ACC_TYPE accumulator (any Q / O): exact sum at max(F), then assignment.
__host__ __device__ __forceinline__ int64_t macc(int64_t acc, const Fmt &fa, i128 p, int Fp) {
  const int Fa = fa.F();
  const int rF = Fa > Fp ? Fa : Fp;
  const i128 s = (i128)((u128)(i128)acc << (rF - Fa)) + (i128)((u128)p << (rF - Fp));
  return convert(s, rF, fa);
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
