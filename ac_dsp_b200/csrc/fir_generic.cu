// fir_generic.cu -- the any-format FIR path: one thread per output sample, taps visited in the
// reference's own order with the full ac_fixed `acc += a*b` re-quantisation at every tap, so every
// quantisation / overflow mode of ACC_TYPE and OUT_TYPE is reproduced (not only the order-independent
// AC_TRN / AC_RND + AC_WRAP ones the fast paths rely on).
//
// Tap order per architecture (reference include/ac_dsp/ac_fir_load_coeffs.h; const / prog are
// line-for-line analogues):
//   SHIFT_REG     :180-188  i = N-1..0   acc += reg[i]*h[i]                 (oldest sample first)
//   ROTATE_SHIFT  :194-208  same order (the rotate only moves data)
//   C_BUFF        :214-224  i = 0..N-1   acc += read(i)*h[i]                (newest sample first)
//   FOLD_EVEN     :231-239  i = N/2-1..0 acc += h[i]*(reg[i]+reg[N-1-i])    (pre-add exact)
//   FOLD_ODD      :246-259  i = 0..(N-1)/2, fold = ACC_TYPE(reg[i]+reg[N-1-i]) (centre: reg[i]), acc += h[i]*fold
//   TRANSPOSED    :265-278  y[n] = q(..q(q(x[n-N+1]h[N-1]) + x[n-N+2]h[N-2]).. + x[n]h[0])  (oldest first)
// reg[k] is the sample k steps back; before the first sample of the stream it is 0 (:134-139).
#include <cstdlib>

#include "kernels.h"

namespace b2d {

struct FirGenArgs {
  Fmt in, coeff, acc, out;
  int N, ftype, ascending;
  uint32_t C;
  int interleaved, in_bytes, out_bytes;
  const void *x;
  void *y;
  size_t n;
  const void *tail;
  const int64_t *h;
  // TRANSPOSED across a coefficient change (see launch_fir_pending below): outputs i < n_limit only; the accumulator of
  // output i < N-1 starts from acc_init[c][i] (the partial sums the old taps left in reg_trans[]); acc_out != null: the
  // inputs of this call are zeros and the raw ACC_TYPE accumulators are written to acc_out[c][i] instead of y.
  size_t n_limit;
  const int64_t *acc_init;
  int64_t *acc_out;
};

__device__ __forceinline__ int64_t fir_gen_sample(const FirGenArgs &a, uint32_t c, size_t i, int k) {
  const int T = a.N - 1;
  if ((size_t)k <= i) return a.acc_out ? 0 : load_raw(a.x, elem_index(i - k, c, a.n, a.C, a.interleaved), a.in_bytes, a.in.S);
  return load_raw(a.tail, (size_t)c * T + (size_t)(T - (k - (int64_t)i)), a.in_bytes, a.in.S);
}

// W: intermediate width, i128 or int64_t (fir_generic_fits64: every product, shifted sum and conversion within 62 bits).
template <class W>
__global__ void __launch_bounds__(256) fir_generic_kernel(FirGenArgs a) {
  const size_t total = a.n_limit * a.C;
  for (size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (size_t)gridDim.x * blockDim.x) {
    const uint32_t c = a.interleaved ? (uint32_t)(t % a.C) : (uint32_t)(t / a.n_limit);
    const size_t i = a.interleaved ? t / a.C : t % a.n_limit;
    const int64_t *h = a.h + (size_t)c * a.N;
    const int N = a.N;
    const int Fin = a.in.F(), Fc = a.coeff.F(), Fa = a.acc.F();
    int64_t acc = (a.acc_init && i < (size_t)(N - 1)) ? a.acc_init[(size_t)c * (N - 1) + i] : 0;
    // _ANTI: pre-subtract instead of pre-add (ac_fir_reg_share.h:151-165,186-205)
    const int anti = a.ftype == B2D_FOLD_EVEN_ANTI || a.ftype == B2D_FOLD_ODD_ANTI;
    switch (a.ftype) {
      case B2D_SHIFT_REG:
      case B2D_ROTATE_SHIFT:
      case B2D_TRANSPOSED:
        if (a.ascending) {
          for (int k = 0; k < N; k++) acc = macc_t<W>(acc, a.acc, (W)fir_gen_sample(a, c, i, k) * (W)h[k], Fin + Fc);
        } else {
          for (int k = N - 1; k >= 0; k--) acc = macc_t<W>(acc, a.acc, (W)fir_gen_sample(a, c, i, k) * (W)h[k], Fin + Fc);
        }
        break;
      case B2D_C_BUFF:
        for (int k = 0; k < N; k++) acc = macc_t<W>(acc, a.acc, (W)fir_gen_sample(a, c, i, k) * (W)h[k], Fin + Fc);
        break;
      case B2D_FOLD_EVEN:
      case B2D_FOLD_EVEN_ANTI:
        for (int q = 0; q < N / 2; q++) {
          const int k = a.ascending ? q : N / 2 - 1 - q;
          const W xb = (W)fir_gen_sample(a, c, i, N - 1 - k);
          const W pre = (W)fir_gen_sample(a, c, i, k) + (anti ? -xb : xb);
          acc = macc_t<W>(acc, a.acc, (W)h[k] * pre, Fin + Fc);
        }
        break;
      case B2D_FOLD_ODD:
      case B2D_FOLD_ODD_ANTI:
        for (int k = 0; k < (N - 1) / 2 + 1; k++) {
          W pre = (W)fir_gen_sample(a, c, i, k);
          if (k != (N - 1) / 2) {
            const W xb = (W)fir_gen_sample(a, c, i, N - 1 - k);
            pre += anti ? -xb : xb;
          }
          const int64_t fold = convert_t<W>(pre, Fin, a.acc);
          acc = macc_t<W>(acc, a.acc, (W)h[k] * (W)fold, Fc + Fa);
        }
        break;
      default: break;
    }
    if (a.acc_out) a.acc_out[(size_t)c * (N - 1) + i] = acc;
    else store_raw(a.y, elem_index(i, c, a.n, a.C, a.interleaved), a.out_bytes, convert_t<W>((W)acc, Fa, a.out));
  }
}

// 64-bit intermediates suffice when every product, every shifted operand of `acc += p` and both conversions stay within
// 62 bits (common.cuh: fits_i64); the folded forms multiply wider operands.
bool fir_generic_fits64(const Fmt &in, const Fmt &coeff, const Fmt &acc, const Fmt &out, int ftype) {
  const int Wi = in.W + (in.S ? 0 : 1), Wc = coeff.W + (coeff.S ? 0 : 1), Wa = acc.W + (acc.S ? 0 : 1);
  switch (ftype) {
    case B2D_FOLD_EVEN: case B2D_FOLD_EVEN_ANTI:
      return fits_i64(acc, out, Wi + 1 + Wc, in.F() + coeff.F());
    case B2D_FOLD_ODD: case B2D_FOLD_ODD_ANTI:
      // fold = ACC_TYPE(a +- b): the pre-add shifted up to F_acc, then h * fold
      if (Wi + 1 + (acc.F() > in.F() ? acc.F() - in.F() : 0) > 62) return false;
      return fits_i64(acc, out, Wc + Wa, coeff.F() + acc.F());
    default:
      return fits_i64(acc, out, Wi + Wc, in.F() + coeff.F());
  }
}

static void fir_generic_launch(const FirGenArgs &a, const FirLaunch &p, size_t blocks, cudaStream_t st) {
  const char *f = getenv("B2D_GENERIC_I128");                 // A/B and test switch: always the 128-bit evaluation
  if (!(f && *f == '1') && fir_generic_fits64(p.fin, p.fcoeff, p.facc, p.fout, a.ftype)) fir_generic_kernel<int64_t><<<(unsigned)blocks, 256, 0, st>>>(a);
  else fir_generic_kernel<i128><<<(unsigned)blocks, 256, 0, st>>>(a);
}

cudaError_t launch_fir_generic(const FirLaunch &p, cudaStream_t st) {
  if (p.n == 0) return cudaSuccess;
  FirGenArgs a;
  a.in = p.fin; a.coeff = p.fcoeff; a.acc = p.facc; a.out = p.fout;
  a.N = p.n_taps; a.ftype = p.ftype; a.ascending = p.ascending; a.C = p.C; a.interleaved = p.interleaved;
  a.in_bytes = container_bytes(p.fin.W); a.out_bytes = container_bytes(p.fout.W);
  a.x = p.in; a.y = p.out; a.n = p.n; a.tail = p.tail; a.h = p.coeff64;
  a.n_limit = p.n; a.acc_init = nullptr; a.acc_out = nullptr;
  const size_t total = p.n * p.C;
  size_t blocks = (total + 255) / 256;
  if (blocks > 148 * 64) blocks = 148 * 64;
  fir_generic_launch(a, p, blocks, st);
  return cudaGetLastError();
}

// TRANSPOSED keeps ACC_TYPE partial sums, not samples (reg_trans[], ac_fir_load_coeffs.h:265-278, ac_fir_prog_coeffs.h:232-247):
//   y[n] = q(..q(q(x[n-N+1] h'[N-1]) + x[n-N+2] h''[N-2]).. + x[n] h[0]),  each tap taken from the set that was active
// when ITS sample arrived.  The engine carries samples; at a coefficient change the runtime turns the history into the
// N-1 pending partial sums (old taps over the history followed by zeros: acc_out mode), clears the history, and the first
// N-1 outputs after the change start their accumulators from those sums (acc_init mode) -- in the reference's own order
// (oldest first), so saturating / sign-dependent ACC_TYPEs come out right as well.
cudaError_t launch_fir_pending(const FirLaunch &p, size_t n_limit, const int64_t *acc_init, int64_t *acc_out, cudaStream_t st) {
  if (n_limit == 0 || p.n_taps < 2) return cudaSuccess;
  FirGenArgs a;
  a.in = p.fin; a.coeff = p.fcoeff; a.acc = p.facc; a.out = p.fout;
  a.N = p.n_taps; a.ftype = B2D_TRANSPOSED; a.ascending = 0; a.C = p.C; a.interleaved = p.interleaved;
  a.in_bytes = container_bytes(p.fin.W); a.out_bytes = container_bytes(p.fout.W);
  a.x = p.in; a.y = p.out; a.n = p.n; a.tail = p.tail; a.h = p.coeff64;
  a.n_limit = n_limit; a.acc_init = acc_init; a.acc_out = acc_out;
  const size_t total = n_limit * p.C;
  size_t blocks = (total + 255) / 256;
  if (blocks > 148 * 64) blocks = 148 * 64;
  fir_generic_launch(a, p, blocks, st);
  return cudaGetLastError();
}

// pending[c][j] <- pending[c][j + n] (zeros shifted in): n more outputs have consumed their partial sums
__global__ void fir_pending_shift_kernel(const int64_t *src, int64_t *dst, size_t n, int T, uint32_t C) {
  const size_t total = (size_t)T * C;
  for (size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (size_t)gridDim.x * blockDim.x) {
    const size_t j = t % T;
    dst[t] = j + n < (size_t)T ? src[t + n] : 0;
  }
}
cudaError_t launch_fir_pending_shift(const int64_t *src, int64_t *dst, size_t n, int T, uint32_t C, cudaStream_t st) {
  if (T <= 0) return cudaSuccess;
  const size_t total = (size_t)T * C;
  size_t blocks = (total + 255) / 256;
  if (blocks > 1024) blocks = 1024;
  fir_pending_shift_kernel<<<(unsigned)blocks, 256, 0, st>>>(src, dst, n, T, C);
  return cudaGetLastError();
}

// ---- history carry -----------------------------------------------------------------------
struct TailArgs {
  const void *in, *tail;
  void *tail_next;
  size_t n;
  int T, bytes;
  uint32_t C;
  int interleaved;
};

__global__ void tail_kernel(TailArgs a) {
  const size_t total = (size_t)a.T * a.C;
  for (size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (size_t)gridDim.x * blockDim.x) {
    const uint32_t c = (uint32_t)(t / a.T);
    const size_t j = t % a.T;
    // position j of the last T elements of (tail ++ in)
    const size_t pos = a.n + j;
    int64_t v;
    if (pos < (size_t)a.T) v = load_raw(a.tail, (size_t)c * a.T + pos, a.bytes, 1);
    else v = load_raw(a.in, elem_index(pos - a.T, c, a.n, a.C, a.interleaved), a.bytes, 1);
    store_raw(a.tail_next, (size_t)c * a.T + j, a.bytes, v);
  }
}

static cudaError_t launch_tail(const void *in, const void *tail, void *tail_next, size_t n, int T, int bytes, uint32_t C,
                               int interleaved, cudaStream_t st) {
  if (T <= 0) return cudaSuccess;
  TailArgs a{in, tail, tail_next, n, T, bytes, C, interleaved};
  const size_t total = (size_t)T * C;
  size_t blocks = (total + 255) / 256;
  if (blocks > 1024) blocks = 1024;
  tail_kernel<<<(unsigned)blocks, 256, 0, st>>>(a);
  return cudaGetLastError();
}

cudaError_t launch_fir_tail(const FirLaunch &p, cudaStream_t st) {
  return launch_tail(p.in, p.tail, p.tail_next, p.n, p.n_taps - 1, container_bytes(p.fin.W), p.C, p.interleaved, st);
}

// reg[N_TAPS-1] of ac_fir_reg_share after this call's n samples: element n-1 of (tail ++ in), converted to OUT_TYPE
struct DelayOutArgs {
  const void *in, *tail;
  int64_t *dl;
  size_t n;
  int T, bytes, in_signed, Fin;
  uint32_t C;
  int interleaved;
  Fmt out;
};
__global__ void fir_delay_out_kernel(DelayOutArgs a) {
  const uint32_t c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= a.C || a.n == 0) return;
  const size_t pos = a.n - 1;
  int64_t v;
  if (pos < (size_t)a.T) v = load_raw(a.tail, (size_t)c * a.T + pos, a.bytes, a.in_signed);
  else v = load_raw(a.in, elem_index(pos - a.T, c, a.n, a.C, a.interleaved), a.bytes, a.in_signed);
  a.dl[c] = convert((i128)v, a.Fin, a.out);
}
cudaError_t launch_fir_delay_out(const FirLaunch &p, int64_t *dl, cudaStream_t st) {
  if (p.n == 0) return cudaSuccess;
  DelayOutArgs a{p.in, p.tail, dl, p.n, p.n_taps - 1, container_bytes(p.fin.W), p.fin.S, p.fin.F(), p.C, p.interleaved, p.fout};
  fir_delay_out_kernel<<<(p.C + 127) / 128, 128, 0, st>>>(a);
  return cudaGetLastError();
}

cudaError_t launch_cic_tail(const CicLaunch &p, cudaStream_t st) {
  return launch_tail(p.in, p.tail, p.tail_next, p.n, p.H, container_bytes(p.fin.W), p.C, p.interleaved, st);
}

}  // namespace b2d
