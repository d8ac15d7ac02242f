// mv_avg.cu -- ac_mv_avg: weighted 1-D moving average over bursts with clip / mirror boundaries.
//
// Replaces ac_mv_avg_core::mvAvgCore + ac_mv_avg::run (reference include/ac_dsp/ac_mv_avg.h:99-127,154-196):
//   acc = 0;  for j = -TAPS/2 .. TAPS/2:  acc = acc + (ACC_TYPE) w[j] * coeffs[j + TAPS/2];  out = acc
// w[j] is the window class of ac_math (ac_window_1d_flag), absent from the reference tree and this image: its boundary
// behaviour is restated from the manual (AC_CLIP repeats the edge sample of the burst, AC_MIRROR reflects about it, AC_WIN
// emits only full windows) -- PARITY UNPINNED, see DESIGN.md.  The window restarts with every burst and nothing survives a
// run() call (the reference constructs its core object inside run()), so every output is an independent function of one
// burst: one thread per output, taps in the reference's order with its two quantisation points per tap (the ACC_TYPE cast
// of the sample, the ACC_TYPE re-quantisation of the sum), 128-bit intermediates -- every Q / O mode.
#include <cstdlib>

#include "kernels.h"

namespace b2d {

struct MvArgs {
  Fmt in, coeff, acc, out;
  int taps, win, in_bytes, out_bytes;
  const void *x;
  void *y;
  const int64_t *c;         // [taps] raw coefficients
  size_t n_sample, per, n_out;   // burst length, outputs per burst, outputs in all
};

// W: intermediate width, i128 or int64_t (common.cuh: fits_i64)
template <class W>
__global__ void __launch_bounds__(256) mvavg_kernel(MvArgs a) {
  const int H = a.taps / 2, Fin = a.in.F(), Fa = a.acc.F(), Fc = a.coeff.F();
  const long long last = (long long)a.n_sample - 1;
  for (size_t o = (size_t)blockIdx.x * blockDim.x + threadIdx.x; o < a.n_out; o += (size_t)gridDim.x * blockDim.x) {
    const size_t b = o / a.per;
    const long long i = (long long)(o % a.per) + (a.win == B2D_WIN ? H : 0);
    const size_t base = b * a.n_sample;
    int64_t acc = 0;
    for (int j = -H; j <= H; j++) {
      long long k = i + j;
      if (a.win == B2D_CLIP) { k = k < 0 ? 0 : (k > last ? last : k); }
      else if (a.win == B2D_MIRROR) { if (k < 0) k = -k; if (k > last) k = 2 * last - k; }
      const int64_t s = load_raw(a.x, base + (size_t)k, a.in_bytes, a.in.S);
      const int64_t cast = convert_t<W>((W)s, Fin, a.acc);                     // (ACC_TYPE) w[j]
      acc = macc_t<W>(acc, a.acc, (W)cast * (W)a.c[j + H], Fa + Fc);            // acc_reg = acc_reg + ... : re-quantised per tap
    }
    store_raw(a.y, o, a.out_bytes, convert_t<W>((W)acc, Fa, a.out));
  }
}

cudaError_t launch_mvavg(const MvLaunch &p, cudaStream_t st) {
  if (p.n_out == 0) return cudaSuccess;
  MvArgs a;
  a.in = p.fin; a.coeff = p.fcoeff; a.acc = p.facc; a.out = p.fout;
  a.taps = p.taps; a.win = p.win; a.in_bytes = container_bytes(p.fin.W); a.out_bytes = container_bytes(p.fout.W);
  a.x = p.in; a.y = p.out; a.c = p.coeff64; a.n_sample = p.n_sample; a.per = p.per; a.n_out = p.n_out;
  size_t blocks = (p.n_out + 255) / 256;
  if (blocks > 148 * 64) blocks = 148 * 64;
  // 64-bit intermediates when the ACC_TYPE cast of the sample, the ACC x COEFF product and both conversions fit 62 bits
  const int Wi = p.fin.W + (p.fin.S ? 0 : 1), Wa = p.facc.W + (p.facc.S ? 0 : 1), Wc = p.fcoeff.W + (p.fcoeff.S ? 0 : 1);
  const bool small = Wi + (p.facc.F() > p.fin.F() ? p.facc.F() - p.fin.F() : 0) <= 62 && fits_i64(p.facc, p.fout, Wa + Wc, p.facc.F() + p.fcoeff.F());
  const char *f128 = getenv("B2D_GENERIC_I128");
  if (small && !(f128 && *f128 == '1')) mvavg_kernel<int64_t><<<(unsigned)blocks, 256, 0, st>>>(a);
  else mvavg_kernel<i128><<<(unsigned)blocks, 256, 0, st>>>(a);
  return cudaGetLastError();
}

}  // namespace b2d
