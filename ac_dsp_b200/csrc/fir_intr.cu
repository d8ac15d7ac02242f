// fir_intr.cu -- polyphase interpolating FIR: ac_poly_intr (reference
// include/ac_dsp/ac_poly_intr.h:103-256, 261-312).
//
// What the reference computes.  One step = one input sample x[k] shifted into the low-rate delay line taps[NTAPS]
// (:127-132) and IF phase accumulators (ACC_TYPE, re-quantised at every `acc +=`):
//   FOLD_EVEN (:122-166)  acc(k,j) = sum_{i=NTAPS/2-1..0} coeffs[i + j*NTAPS/2] * fold_i,
//                         fold_i = ACC_TYPE(taps[i] + tp),  tp = sign[j] ? taps[NTAPS-1-i] : IN_TYPE(-taps[NTAPS-1-i])
//   FOLD_ODD  (:172-226)  i = 0 .. (NTAPS-1)/2 upwards, centre fold = ACC_TYPE(taps[i]), coeffs[i + (NTAPS/2+1)*j]
//   FOLD_ANTI (:232-251)  acc(k,j) = sum_{i=NTAPS-1..0} taps[i] * coeffs[i + NTAPS*j], written at once as OUT_TYPE.
// The folded forms park acc(k,.) in acc_a / acc_b and write, one step LATER (`init`, `flip`, :147-164),
//   out[k][j] = OUT_TYPE( corr[j] == j ? acc(k-1,j) : (acc(k-1,j) + (sign[j] ? ACC(-acc(k-1,corr[j])) : acc(k-1,corr[j]))) >> 1 )
// -- the symmetric-pair technique: two phases whose coefficient sets are mirror images share sum / difference sets.
// So every output is a pure function of one window of NTAPS inputs: one thread per output (row, phase), windows
// read through L1 (the threads of a warp share theirs), the accumulators of the call's last step carried to the next
// call in `carry` (they were computed with the coefficients of THAT step, which a later load must not change).
//   polyintr_kernel<true>   wrapping accumulators with AC_TRN / AC_RND: 64-bit modular arithmetic
//   polyintr_kernel<false>  every format / mode: 128-bit intermediates in the reference's tap order
// The plain polyphase form on 16-bit operands (FOLD_ANTI, exact accumulators) is upfir_lane_kernel (upfir_q15.cu).
#include "kernels.h"
#include <algorithm>

namespace b2d {

struct PiArgs {
  Fmt in, fc, acc, out;
  int nt, ifac, ftype, csz;
  int in_bytes, out_bytes;
  uint32_t C;
  int interleaved;
  const void *x;            // n inputs per channel
  void *y;                  // n_out outputs per channel, planar stride n_out
  size_t n, n_rows;         // rows of IF outputs in this call
  int row_shift;            // source step of row r is r - row_shift (folded forms: 1 once init, 0 before; FOLD_ANTI: 0)
  const void *tail;         // [C][H] previous inputs, oldest first
  int H;
  const int64_t *coeff;     // [C][csz]
  const uint8_t *sign, *corr;   // [C][ifac]
  const int64_t *carry;     // [C][ifac] acc(.) of the step before this call
  int64_t *carry_next;
  // fast path constants
  int sl_fold, sr_fold;     // fold: left shift F_acc - F_in, or right shift F_in - F_acc
  int s_prod;               // product: right shift (> 0) or left shift (<= 0)
  long long rnd_fold, rnd_prod;
};

__device__ __forceinline__ int64_t pi_sample(const PiArgs &a, uint32_t c, long long idx) {
  if (idx >= 0) return load_raw(a.x, elem_index((size_t)idx, c, a.n, a.C, a.interleaved), a.in_bytes, a.in.S);
  if (idx < -(long long)a.H) return 0;
  return load_raw(a.tail, (size_t)c * a.H + (size_t)(a.H + idx), a.in_bytes, a.in.S);
}

// acc(step, j) of channel c; step is local to the call (history below 0)
template <bool FAST>
__device__ int64_t pi_acc(const PiArgs &a, uint32_t c, long long step, int j) {
  const int NT = a.nt;
  const int64_t *h = a.coeff + (size_t)c * a.csz;
  const bool sg = a.sign[(size_t)c * a.ifac + j] != 0;
  const int Fin = a.in.F(), Fc = a.fc.F(), Fa = a.acc.F();
  int64_t acc = 0;
  if (a.ftype == B2D_PI_FOLD_ANTI) {
    const int64_t *hj = h + (size_t)NT * j;
    for (int i = NT - 1; i >= 0; i--) {
      const int64_t x = pi_sample(a, c, step - i);
      if (FAST) {
        const int64_t p = x * hj[i];
        acc += a.s_prod > 0 ? (p + a.rnd_prod) >> a.s_prod : (int64_t)((uint64_t)p << (-a.s_prod));
      } else {
        acc = macc(acc, a.acc, (i128)x * (i128)hj[i], Fin + Fc);
      }
    }
    return FAST ? wrap_bits(acc, a.acc.W, a.acc.S) : acc;
  }
  const bool even = a.ftype == B2D_PI_FOLD_EVEN;
  const int nloop = even ? NT / 2 : (NT - 1) / 2 + 1;
  const int64_t *hj = h + (size_t)(even ? NT / 2 : NT / 2 + 1) * j;
  for (int q = 0; q < nloop; q++) {
    const int i = even ? nloop - 1 - q : q;                       // FOLD_EVEN walks downwards, FOLD_ODD upwards
    const int64_t xa = pi_sample(a, c, step - i);
    const bool centre = !even && i == (NT - 1) / 2;
    const int64_t xb = centre ? 0 : pi_sample(a, c, step - (NT - 1 - i));
    if (FAST) {
      const int64_t tp = centre ? 0 : (sg ? xb : wrap_bits(-xb, a.in.W, a.in.S));
      const int64_t s = xa + tp;
      const int64_t fold = wrap_bits(a.sr_fold > 0 ? (s + a.rnd_fold) >> a.sr_fold : (int64_t)((uint64_t)s << a.sl_fold), a.acc.W, a.acc.S);
      const int64_t p = (int64_t)((uint64_t)fold * (uint64_t)hj[i]);      // modulo 2^64: the bits [F_c, F_c + W_acc) are exact
      acc += a.s_prod > 0 ? (p + a.rnd_prod) >> a.s_prod : (int64_t)((uint64_t)p << (-a.s_prod));
    } else {
      const i128 tp = centre ? 0 : (sg ? (i128)xb : (i128)convert(-(i128)xb, Fin, a.in));
      const int64_t fold = convert((i128)xa + tp, Fin, a.acc);
      acc = macc(acc, a.acc, (i128)hj[i] * (i128)fold, Fc + Fa);
    }
  }
  return FAST ? wrap_bits(acc, a.acc.W, a.acc.S) : acc;
}

template <bool FAST>
__global__ void __launch_bounds__(256) polyintr_kernel(PiArgs a) {
  const uint32_t c = blockIdx.y;
  const size_t total = a.n_rows * a.ifac;
  const int Fa = a.acc.F();
  for (size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (size_t)gridDim.x * blockDim.x) {
    const size_t r = t / a.ifac;
    const int j = (int)(t % a.ifac);
    const long long step = (long long)r - a.row_shift;
    int64_t o;
    if (a.ftype == B2D_PI_FOLD_ANTI) {
      o = convert((i128)pi_acc<FAST>(a, c, step, j), Fa, a.out);
    } else {
      const int cj = a.corr[(size_t)c * a.ifac + j];
      const int64_t t1 = step < 0 ? a.carry[(size_t)c * a.ifac + j] : pi_acc<FAST>(a, c, step, j);
      if (cj != j) {
        const int64_t t2 = step < 0 ? a.carry[(size_t)c * a.ifac + cj] : pi_acc<FAST>(a, c, step, cj);
        const int64_t tn = a.sign[(size_t)c * a.ifac + j] ? convert(-(i128)t2, Fa, a.acc) : t2;
        o = convert(((i128)t1 + (i128)tn) >> 1, Fa, a.out);      // W_acc + 1 bits: the shift drops the LSB only
      } else {
        o = convert((i128)t1, Fa, a.out);
      }
    }
    store_raw(a.y, (size_t)c * (a.n_rows * a.ifac) + t, a.out_bytes, o);
  }
}

// acc(n-1, .) of this call -> carry_next (folded forms)
template <bool FAST>
__global__ void polyintr_carry_kernel(PiArgs a) {
  const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= (size_t)a.C * a.ifac) return;
  const uint32_t c = (uint32_t)(t / a.ifac);
  const int j = (int)(t % a.ifac);
  a.carry_next[t] = pi_acc<FAST>(a, c, (long long)a.n - 1, j);
}

bool polyintr_fast_supported(const Fmt &in, const Fmt &coeff, const Fmt &acc, int ftype) {
  if (acc.O != B2D_WRAP || (acc.Q != B2D_TRN && acc.Q != B2D_RND)) return false;
  const int Fin = in.F(), Fc = coeff.F(), Fa = acc.F();
  if (ftype == B2D_PI_FOLD_ANTI) { const int s = Fin + Fc - Fa; return s < 63 && s > -63; }
  if (in.O != B2D_WRAP) return false;
  if (std::max(Fc, 0) + acc.W > 64 || Fc < -62) return false;
  const int d = Fin - Fa;
  return d < 62 && d > -62;
}

cudaError_t launch_polyintr(const PiLaunch &p, cudaStream_t st) {
  PiArgs a;
  a.in = p.fin; a.fc = p.fcoeff; a.acc = p.facc; a.out = p.fout; a.nt = p.nt; a.ifac = p.ifac; a.ftype = p.ftype; a.csz = p.csz;
  a.in_bytes = container_bytes(p.fin.W); a.out_bytes = container_bytes(p.fout.W);
  a.C = p.C; a.interleaved = p.interleaved && p.C > 1; a.x = p.in; a.y = p.out; a.n = p.n; a.n_rows = p.n_rows; a.row_shift = p.row_shift;
  a.tail = p.tail; a.H = p.H; a.coeff = p.coeff64; a.sign = p.sign; a.corr = p.corr; a.carry = p.carry; a.carry_next = p.carry_next;
  const int Fin = p.fin.F(), Fc = p.fcoeff.F(), Fa = p.facc.F();
  const bool rnd = p.facc.Q == B2D_RND;
  a.sl_fold = a.sr_fold = 0; a.rnd_fold = a.rnd_prod = 0;
  if (p.ftype == B2D_PI_FOLD_ANTI) a.s_prod = Fin + Fc - Fa;
  else {
    a.s_prod = Fc;
    if (Fin > Fa) { a.sr_fold = Fin - Fa; a.rnd_fold = rnd ? 1LL << (a.sr_fold - 1) : 0; }
    else a.sl_fold = Fa - Fin;
  }
  if (a.s_prod > 0 && rnd) a.rnd_prod = 1LL << (a.s_prod - 1);
  const size_t total = p.n_rows * p.ifac;
  if (total) {
    size_t bx = (total + 255) / 256;
    if (bx > 148 * 16) bx = 148 * 16;
    dim3 grid((unsigned)bx, p.C);
    if (p.fast) polyintr_kernel<true><<<grid, 256, 0, st>>>(a);
    else polyintr_kernel<false><<<grid, 256, 0, st>>>(a);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return e;
  }
  if (p.ftype != B2D_PI_FOLD_ANTI && p.n > 0) {
    const unsigned nb = (unsigned)(((size_t)p.C * p.ifac + 127) / 128);
    if (p.fast) polyintr_carry_kernel<true><<<nb, 128, 0, st>>>(a);
    else polyintr_carry_kernel<false><<<nb, 128, 0, st>>>(a);
  }
  return cudaGetLastError();
}

}  // namespace b2d
