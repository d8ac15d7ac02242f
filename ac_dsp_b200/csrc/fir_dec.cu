// fir_dec.cu -- polyphase decimating FIR: ac_poly_dec (reference include/ac_dsp/ac_poly_dec.h:87-137).
//
// What the reference computes.  For every group of DF input samples it shifts them one by one into a register of
// NTAPS*DF samples and, after each shift (phase df = DF-1 .. 0), accumulates acc1[df] = sum_tp taps[tp*DF] *
// coeffs[tp + NTAPS*df], then acc += acc1[df]; one output per group (:110-126).  In closed form, with
// u_r[m] = x[m*DF + DF-1 - r] the r-th polyphase component of the input,
//     out[m] = sum_{r < DF} sum_{tp < NTAPS} coeffs[tp + NTAPS*r] * u_r[m - tp]
// i.e. DF ordinary FIR filters on the de-interleaved streams, summed -- the coefficient array is already in phase
// order.  Every `+=` re-quantises to ACC_TYPE; with Q in {AC_TRN, AC_RND} and O = AC_WRAP that is a modular sum of
// independently quantised products (fir_wide.cu), so phases and taps may run in any order and in parallel.
//
//   polydec_wide_kernel<MODE>  operands <= 32 bits, wrapping <= 64-bit accumulator: a CTA de-interleaves a tile of
//       1024*DF samples (plus history) into DF shared-memory planes; a thread owns 8 consecutive outputs and runs the
//       IMAD.WIDE sliding-window block of fir_wide.cuh once per phase over that phase's plane and taps.
//   polydec_generic_kernel     every Q / O mode: one thread per output, the reference's own order (phases DF-1 .. 0,
//       taps upwards, partial accumulator added to the total with one more ACC_TYPE assignment), 128-bit intermediates.
#include "fir_wide.cuh"

namespace b2d {

constexpr int kDecTile = kWideThreads * kWideT;   // outputs per CTA

struct DecArgs {
  Fmt in, coeff, acc, out;
  int NT, NTpad, DF;
  uint32_t C;
  int interleaved, in_bytes, out_bytes, in_signed, fastout;
  int s;
  long long rnd;
  const void *x;
  void *y;
  const void *tail;       // [C][NT*DF - 1] previous samples
  size_t n, n_out;
  long long n_seen, m_first;
  const int32_t *c32;     // [C][DF][NTpad]
  const int64_t *c64;     // [C][NT*DF] phase order (generic kernel)
};

// sample with global index g (history for g < n_seen, zero before the stream / past the call)
__device__ __forceinline__ int64_t dec_sample(const DecArgs &a, uint32_t c, long long g) {
  const long long li = g - a.n_seen;
  const int T = a.NT * a.DF - 1;
  if (li >= 0) return (size_t)li < a.n ? load_raw(a.x, elem_index((size_t)li, c, a.n, a.C, a.interleaved), a.in_bytes, a.in_signed) : 0;
  if (li < -(long long)T) return 0;
  return load_raw(a.tail, (size_t)c * T + (size_t)(T + li), a.in_bytes, a.in_signed);
}

template <int MODE>
__global__ void __launch_bounds__(kWideThreads) polydec_wide_kernel(DecArgs a) {
  extern __shared__ __align__(16) unsigned char smem[];
  const int XS = a.NTpad + kDecTile;                 // samples per phase plane: NTpad of history + the tile
  int32_t *cs = (int32_t *)smem;                     // [DF][NTpad]
  int32_t *xs = cs + (size_t)a.DF * a.NTpad;         // [DF][XS]: xs[r][NTpad + k] = u_r[m0 + k]
  const uint32_t c = blockIdx.y;
  const long long m0 = a.m_first + (long long)blockIdx.x * kDecTile;

  for (int i = threadIdx.x; i < a.DF * a.NTpad; i += kWideThreads) cs[i] = a.c32[(size_t)c * a.DF * a.NTpad + i];
  const long long g_base = (m0 - a.NTpad) * a.DF;    // first sample staged (consecutive samples -> coalesced loads)
  for (int q = threadIdx.x; q < XS * a.DF; q += kWideThreads) {
    const int k = q / a.DF, pos = q - k * a.DF;      // u_r[m] = x[m*DF + DF-1 - r]  <=>  r = DF-1 - pos
    xs[(size_t)(a.DF - 1 - pos) * XS + k] = (int)dec_sample(a, c, g_base + q);
  }
  __syncthreads();

  const int o = threadIdx.x * kWideT;
  const long long j0 = m0 - a.m_first + o;           // local index of this thread's first output
  if ((size_t)j0 >= a.n_out) return;
  long long acc[kWideT];
#pragma unroll
  for (int j = 0; j < kWideT; j++) acc[j] = 0;
  for (int r = 0; r < a.DF; r++)
    wide_mac_block<MODE>(xs + (size_t)r * XS, a.NTpad + o, cs + (size_t)r * a.NTpad, a.NTpad, a.s, a.rnd, acc);
#pragma unroll
  for (int j = 0; j < kWideT; j++) {
    if ((size_t)(j0 + j) >= a.n_out) break;
    long long v = acc[j];
    if (MODE == 0) v = (long long)((unsigned long long)v << (-a.s));
    v = wrap_bits(v, a.acc.W, a.acc.S);
    const size_t idx = (size_t)c * a.n_out + (size_t)(j0 + j);
    if (a.fastout) ((long long *)a.y)[idx] = v;
    else store_raw(a.y, idx, a.out_bytes, convert((i128)v, a.acc.F(), a.out));
  }
}

__global__ void __launch_bounds__(256) polydec_generic_kernel(DecArgs a) {
  const size_t total = a.n_out * a.C;
  const int Fin = a.in.F(), Fc = a.coeff.F(), Fa = a.acc.F();
  for (size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (size_t)gridDim.x * blockDim.x) {
    const uint32_t c = (uint32_t)(t / a.n_out);
    const size_t j = t % a.n_out;
    const long long m = a.m_first + (long long)j;
    const int64_t *h = a.c64 + (size_t)c * a.NT * a.DF;
    int64_t acc = 0;
    for (int df = a.DF - 1; df >= 0; df--) {                       // ac_poly_dec.h:110-123
      const long long g = m * a.DF + (a.DF - 1 - df);              // newest sample when phase df is evaluated
      int64_t acc1 = 0;
      for (int tp = 0; tp < a.NT; tp++)
        acc1 = macc(acc1, a.acc, (i128)dec_sample(a, c, g - (long long)tp * a.DF) * (i128)h[tp + a.NT * df], Fin + Fc);
      acc = macc(acc, a.acc, (i128)acc1, Fa);                      // acc = acc + acc1[df]
    }
    store_raw(a.y, (size_t)c * a.n_out + j, a.out_bytes, convert((i128)acc, Fa, a.out));
  }
}

// ------------------------------------------------------------------------------------------ host side
static size_t dec_smem(int ntpad, int df) { return (size_t)df * ((size_t)2 * ntpad + kDecTile) * 4; }

int polydec_wide_mode(const Fmt &in, const Fmt &coeff, const Fmt &acc, int ntaps, int df) {
  const int m = fir_wide_mode(in, coeff, acc, 1, B2D_SHIFT_REG);   // format rules of the wide path; the tap count is checked here
  if (m < 0) return -1;
  if (dec_smem(polydec_words(ntaps), df) > 200 * 1024) return -1;
  return m;
}

int polydec_words(int ntaps) { return (ntaps + 7) & ~7; }

// phase-order taps -> [DF][NTpad] int32, zero padded
void polydec_pack(const int64_t *c, int ntaps, int df, int32_t *out) {
  const int ntpad = polydec_words(ntaps);
  for (int r = 0; r < df; r++)
    for (int tp = 0; tp < ntpad; tp++) out[(size_t)r * ntpad + tp] = tp < ntaps ? (int32_t)c[tp + ntaps * r] : 0;
}

cudaError_t launch_polydec(const DecLaunch &p, cudaStream_t st) {
  if (p.n_out == 0) return cudaSuccess;
  DecArgs a;
  a.in = p.fin; a.coeff = p.fcoeff; a.acc = p.facc; a.out = p.fout;
  a.NT = p.nt; a.NTpad = polydec_words(p.nt); a.DF = p.df; a.C = p.C; a.interleaved = p.interleaved;
  a.in_bytes = container_bytes(p.fin.W); a.out_bytes = container_bytes(p.fout.W); a.in_signed = p.fin.S;
  a.fastout = (p.fout.W == p.facc.W && p.fout.I == p.facc.I && p.fout.S == p.facc.S && a.out_bytes == 8) ? 1 : 0;
  a.s = p.fin.F() + p.fcoeff.F() - p.facc.F();
  a.rnd = (a.s > 0 && p.facc.Q == B2D_RND) ? (1LL << (a.s - 1)) : 0;
  a.x = p.in; a.y = p.out; a.tail = p.tail; a.n = p.n; a.n_out = p.n_out;
  a.n_seen = (long long)p.n_seen; a.m_first = (long long)(p.n_seen / (unsigned long long)p.df);
  a.c32 = p.coeff32; a.c64 = p.coeff64;
  if (!p.wide) {
    const size_t total = p.n_out * p.C;
    size_t blocks = (total + 255) / 256;
    if (blocks > 148 * 64) blocks = 148 * 64;
    polydec_generic_kernel<<<(unsigned)blocks, 256, 0, st>>>(a);
    return cudaGetLastError();
  }
  const int mode = polydec_wide_mode(p.fin, p.fcoeff, p.facc, p.nt, p.df);
  if (mode < 0) return cudaErrorNotSupported;
  const size_t smem = dec_smem(a.NTpad, a.DF);
  dim3 grid((unsigned)((p.n_out + kDecTile - 1) / kDecTile), p.C);
  cudaError_t e = cudaSuccess;
  if (mode == 0) {
    if (smem > 48 * 1024) e = cudaFuncSetAttribute(polydec_wide_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e == cudaSuccess) polydec_wide_kernel<0><<<grid, kWideThreads, smem, st>>>(a);
  } else {
    if (smem > 48 * 1024) e = cudaFuncSetAttribute(polydec_wide_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e == cudaSuccess) polydec_wide_kernel<1><<<grid, kWideThreads, smem, st>>>(a);
  }
  return e != cudaSuccess ? e : cudaGetLastError();
}

}  // namespace b2d
