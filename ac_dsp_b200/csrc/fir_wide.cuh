// fir_wide.cuh -- the IMAD.WIDE sliding-window MAC block shared by fir_wide.cu and fir_dec.cu.
#pragma once
#include "kernels.h"

namespace b2d {

constexpr int kWideThreads = 128;
constexpr int kWideT = 8;             // consecutive outputs per thread

// acc[j] += sum_{i < Npad} q( xs[B + j - i] * cs[i] ),  j < kWideT;  q = identity (MODE 0) or floor((p + rnd) / 2^s) (MODE 1).
// xs + B and cs are 16-byte aligned, Npad is a multiple of 8, xs[B - Npad .. B + 7] is readable.
// An 8-sample register window slides 8 taps per step: 2 LDS.128 of samples + 2 LDS.128 of (broadcast) taps per 64 MACs.
template <int MODE>
__device__ __forceinline__ void wide_mac_block(const int32_t *xs, int B, const int32_t *cs, int Npad, int s, long long rnd,
                                               long long (&acc)[kWideT]) {
  int win[kWideT];
  {
    const int4 w0 = *(const int4 *)(xs + B), w1 = *(const int4 *)(xs + B + 4);
    win[0] = w0.x; win[1] = w0.y; win[2] = w0.z; win[3] = w0.w; win[4] = w1.x; win[5] = w1.y; win[6] = w1.z; win[7] = w1.w;
  }
  for (int i0 = 0; i0 < Npad; i0 += 8) {
    int nw[8], h[8];
    {
      const int4 v0 = *(const int4 *)(xs + B - i0 - 8), v1 = *(const int4 *)(xs + B - i0 - 4);
      nw[0] = v0.x; nw[1] = v0.y; nw[2] = v0.z; nw[3] = v0.w; nw[4] = v1.x; nw[5] = v1.y; nw[6] = v1.z; nw[7] = v1.w;
      const int4 c0 = *(const int4 *)(cs + i0), c1 = *(const int4 *)(cs + i0 + 4);
      h[0] = c0.x; h[1] = c0.y; h[2] = c0.z; h[3] = c0.w; h[4] = c1.x; h[5] = c1.y; h[6] = c1.z; h[7] = c1.w;
    }
#pragma unroll
    for (int t = 0; t < 8; t++) {
#pragma unroll
      for (int j = 0; j < kWideT; j++) {
        const int xv = (j - t >= 0) ? win[(j - t) & 7] : nw[(8 + j - t) & 7];   // x[n0 + j - i0 - t]
        const long long p = (long long)xv * (long long)h[t];
        if (MODE == 0) acc[j] += p;
        else acc[j] += (p + rnd) >> s;
      }
    }
#pragma unroll
    for (int j = 0; j < kWideT; j++) win[j] = nw[j];
  }
}

}  // namespace b2d
