// fir_wide.cu -- FIR path for operands up to 32 bits with a wrapping 64-bit accumulator: the formats of the
// reference's own three FIR benches (<16,8>x<32,16>, <32,16>x<32,16>, <28,6>x<23,7> -> <64,32>, FOLD_ODD;
// tests/rtest_ac_fir_{const,load,prog}_coeffs.cpp) and of BASELINE config 5's second stage (<20,5>x<16,1> -> <40,8>).
//
// Replaces the same tap-MAC loops as fir_q15.cu (reference ac_fir_load_coeffs.h:180-278 and the const / prog
// analogues) whenever ACC_TYPE has Q in {AC_TRN, AC_RND} and O = AC_WRAP, which makes `acc += a*b` a modular sum of
// independently quantised terms:  acc_raw = wrap_W( sum_i q(p_i) ),  q(p) = floor((p + rnd) / 2^s)  for
// s = F_in + F_c - F_acc > 0 (rnd = 2^(s-1) for AC_RND, else 0)  and  q(p) = p * 2^-s  for s <= 0.  Everything is
// evaluated modulo 2^64, which contains modulo 2^W_acc.
//   MODE 0  s <= 0: no bits are dropped, the sum is linear -> one IMAD.WIDE per tap into a 64-bit accumulator and a
//           single shift at the end; the folded architectures are expanded into effective direct-form taps.
//   MODE 1  s > 0, direct forms: per-tap arithmetic shift of the 64-bit product.
//   MODE 2  s > 0, FOLD_EVEN / FOLD_ODD: the shift applies to h[i] * (x[n-i] + x[n-N+1+i]) (:231-259), which differs
//           from the direct form; FOLD_ODD is taken only when its ACC_TYPE `fold` pre-add is exact (fir_wide_supported).
// A CTA stages tile + N_TAPS - 1 samples (sign-extended to int32) and the taps in shared memory; a thread owns 8
// consecutive outputs and slides an 8 + 8 sample register window (2 LDS.128 of samples + 2 LDS.128 of broadcast taps
// per 64 MACs in modes 0 / 1).
#include <vector>

#include "fir_wide.cuh"

namespace b2d {

constexpr int kWideTile = kWideThreads * kWideT;
constexpr int kWideMaxTaps = 4096;

struct WideArgs {
  const void *x;
  void *y;
  const void *tail;
  const int32_t *coeff;   // [C][Npad]
  size_t n;
  int N, Npad;
  uint32_t C;
  int interleaved, in_bytes, out_bytes, in_signed;
  int s;                  // F_in + F_c - F_acc
  long long rnd;          // 2^(s-1) for an AC_RND accumulator with s > 0
  int npairs, centre;     // MODE 2: taps i < npairs are paired with N-1-i; tap `centre` (or -1) is unpaired
  int pair_sign;          // MODE 2: +1 pre-add, -1 pre-subtract (the _ANTI architectures)
  Fmt acc, out;
  int fastout;
};

__device__ __forceinline__ int wide_load(const WideArgs &a, const void *p, size_t idx) {
  if (a.in_bytes == 2) return a.in_signed ? (int)((const int16_t *)p)[idx] : (int)((const uint16_t *)p)[idx];
  return ((const int32_t *)p)[idx];
}

template <int MODE>
__global__ void __launch_bounds__(kWideThreads) fir_wide_kernel(WideArgs a) {
  extern __shared__ __align__(16) unsigned char smem[];
  const int T = a.N - 1;
  const int Tpad = (a.Npad + 3) & ~3;               // history slots in front of the tile (>= Npad - 1 + 1, 16-byte multiple)
  int32_t *cs = (int32_t *)smem;                    // [Npad]
  int32_t *xs = cs + a.Npad;                        // [Tpad + tile]: xs[Tpad + q] = x[tile0 + q]
  const uint32_t c = blockIdx.y;
  const long long tile0 = (long long)blockIdx.x * kWideTile;

  for (int i = threadIdx.x; i < a.Npad; i += kWideThreads) cs[i] = a.coeff[(size_t)c * a.Npad + i];
  for (int q = threadIdx.x; q < Tpad + kWideTile; q += kWideThreads) {
    const long long g = tile0 - Tpad + q;           // sample index inside this call
    int v = 0;
    if (g >= 0) { if ((size_t)g < a.n) v = wide_load(a, a.x, elem_index((size_t)g, c, a.n, a.C, a.interleaved)); }
    else if (g >= -(long long)T) v = wide_load(a, a.tail, (size_t)c * T + (size_t)(T + g));
    xs[q] = v;
  }
  __syncthreads();

  const int o = threadIdx.x * kWideT;               // first output of this thread inside the tile
  const long long n0 = tile0 + o;
  if ((size_t)n0 >= a.n) return;
  const int B = Tpad + o;                           // xs[B + j] = x[n0 + j]
  long long acc[kWideT];
#pragma unroll
  for (int j = 0; j < kWideT; j++) acc[j] = 0;

  if (MODE == 2) {
    for (int i = 0; i < a.npairs; i++) {
      const long long h = cs[i];
#pragma unroll
      for (int j = 0; j < kWideT; j++) {
        const long long pre = (long long)xs[B + j - i] + (long long)(a.pair_sign * (long long)xs[B + j - T + i]);
        acc[j] += (h * pre + a.rnd) >> a.s;
      }
    }
    if (a.centre >= 0) {
      const long long h = cs[a.centre];
#pragma unroll
      for (int j = 0; j < kWideT; j++) acc[j] += (h * (long long)xs[B + j - a.centre] + a.rnd) >> a.s;
    }
  } else {
    wide_mac_block<MODE>(xs, B, cs, a.Npad, a.s, a.rnd, acc);
  }

#pragma unroll
  for (int j = 0; j < kWideT; j++) {
    if ((size_t)(n0 + j) >= a.n) break;
    long long r = acc[j];
    if (MODE == 0) r = (long long)((unsigned long long)r << (-a.s));
    r = wrap_bits(r, a.acc.W, a.acc.S);
    const size_t idx = elem_index((size_t)(n0 + j), c, a.n, a.C, a.interleaved);
    if (a.fastout) ((long long *)a.y)[idx] = r;
    else store_raw(a.y, idx, a.out_bytes, convert((i128)r, a.acc.F(), a.out));
  }
}

// ------------------------------------------------------------------------------------------ host side
static bool fits_i32(const Fmt &f) { return f.W + (f.S ? 0 : 1) <= 32; }

static bool fold_odd_exact(const Fmt &in, const Fmt &acc, bool anti) {
  // `fold` is ACC_TYPE (ac_fir_load_coeffs.h:248-255): exact iff the pre-add neither drops fraction bits nor wraps there;
  // an unsigned ACC_TYPE wraps every negative pre-add (signed samples, or the _ANTI pre-subtract) before the multiply
  if (!acc.S && (in.S || anti)) return false;
  return acc.F() >= in.F() && in.W + 1 + (in.S ? 0 : 1) + (acc.F() - in.F()) <= acc.W + (acc.S ? 0 : 1);
}

int fir_wide_mode(const Fmt &in, const Fmt &coeff, const Fmt &acc, int n_taps, int ftype) {
  if (!fits_i32(in) || !fits_i32(coeff)) return -1;
  if (acc.O != B2D_WRAP || (acc.Q != B2D_TRN && acc.Q != B2D_RND)) return -1;
  if (n_taps > kWideMaxTaps) return -1;
  const int s = in.F() + coeff.F() - acc.F();
  if (s < -63 || s > 61) return -1;
  const bool anti = ftype == B2D_FOLD_EVEN_ANTI || ftype == B2D_FOLD_ODD_ANTI;
  const bool fold = ftype == B2D_FOLD_EVEN || ftype == B2D_FOLD_ODD || anti;
  if ((ftype == B2D_FOLD_ODD || ftype == B2D_FOLD_ODD_ANTI) && !fold_odd_exact(in, acc, ftype == B2D_FOLD_ODD_ANTI)) return -1;
  if (s <= 0) return (anti && coeff.W + (coeff.S ? 0 : 1) > 31) ? -1 : 0;   // the mirrored taps are negated
  if (!fold) return 1;
  if (in.W + (in.S ? 0 : 1) > 31) return -1;     // |h * (xa + xb)| must stay below 2^62
  return 2;
}

bool fir_wide_supported(const Fmt &in, const Fmt &coeff, const Fmt &acc, const Fmt &out, int n_taps, int ftype) {
  (void)out;
  switch (ftype) {
    case B2D_SHIFT_REG: case B2D_ROTATE_SHIFT: case B2D_C_BUFF: case B2D_TRANSPOSED: case B2D_FOLD_EVEN: case B2D_FOLD_ODD:
    case B2D_FOLD_EVEN_ANTI: case B2D_FOLD_ODD_ANTI: break;
    default: return false;
  }
  return fir_wide_mode(in, coeff, acc, n_taps, ftype) >= 0;
}

int fir_wide_words(int n_taps) { return (n_taps + 7) & ~7; }

// taps as the kernel wants them: MODE 0 -> effective direct-form taps (folds expanded), else the raw array; zero padded
void fir_wide_pack(const int64_t *c, int n_taps, int ftype, int mode, int32_t *out, int words) {
  const int N = n_taps;
  std::vector<int64_t> eff(c, c + N);
  const int64_t sg = (ftype == B2D_FOLD_EVEN_ANTI || ftype == B2D_FOLD_ODD_ANTI) ? -1 : 1;   // ac_fir_reg_share.h:151-165,186-205
  if (mode == 0 && (ftype == B2D_FOLD_EVEN || ftype == B2D_FOLD_EVEN_ANTI)) {   // ac_fir_load_coeffs.h:231-239: taps i and N-1-i share h[i], i < N/2
    for (int i = 0; i < N; i++) eff[i] = 0;
    for (int i = 0; i < N / 2; i++) { eff[i] = c[i]; eff[N - 1 - i] = sg * c[i]; }
  } else if (mode == 0 && (ftype == B2D_FOLD_ODD || ftype == B2D_FOLD_ODD_ANTI)) {   // :246-259: i <= (N-1)/2, the last one unpaired
    for (int i = 0; i < N; i++) eff[i] = 0;
    for (int i = 0; i < (N - 1) / 2 + 1; i++) {
      eff[i] = c[i];
      if (i != (N - 1) / 2) eff[N - 1 - i] = sg * c[i];
    }
  }
  for (int i = 0; i < words; i++) out[i] = i < N ? (int32_t)eff[i] : 0;
}

cudaError_t launch_fir_wide(const FirLaunch &p, cudaStream_t st) {
  if (p.n == 0) return cudaSuccess;
  WideArgs a;
  const int mode = fir_wide_mode(p.fin, p.fcoeff, p.facc, p.n_taps, p.ftype);
  if (mode < 0) return cudaErrorNotSupported;
  a.x = p.in; a.y = p.out; a.tail = p.tail; a.coeff = p.coeff32; a.n = p.n;
  a.N = p.n_taps; a.Npad = fir_wide_words(p.n_taps); a.C = p.C; a.interleaved = p.interleaved;
  a.in_bytes = container_bytes(p.fin.W); a.out_bytes = container_bytes(p.fout.W); a.in_signed = p.fin.S;
  a.s = p.fin.F() + p.fcoeff.F() - p.facc.F();
  a.rnd = (a.s > 0 && p.facc.Q == B2D_RND) ? (1LL << (a.s - 1)) : 0;
  // FOLD_EVEN: i < N/2 paired, nothing else is read (:231-239).  FOLD_ODD: i < (N-1)/2 paired, i = (N-1)/2 unpaired (:246-259)
  const bool odd = p.ftype == B2D_FOLD_ODD || p.ftype == B2D_FOLD_ODD_ANTI;
  a.npairs = odd ? (p.n_taps - 1) / 2 : p.n_taps / 2;
  a.centre = odd ? (p.n_taps - 1) / 2 : -1;
  a.pair_sign = (p.ftype == B2D_FOLD_EVEN_ANTI || p.ftype == B2D_FOLD_ODD_ANTI) ? -1 : 1;
  a.acc = p.facc; a.out = p.fout;
  a.fastout = (p.fout.W == p.facc.W && p.fout.I == p.facc.I && p.fout.S == p.facc.S && a.out_bytes == 8) ? 1 : 0;
  const int Tpad = (a.Npad + 3) & ~3;
  const size_t smem = ((size_t)a.Npad + Tpad + kWideTile) * 4;
  dim3 grid((unsigned)((p.n + kWideTile - 1) / kWideTile), p.C);
  cudaError_t e = cudaSuccess;
#define B2D_WIDE_LAUNCH(M)                                                                                              \
  do {                                                                                                                  \
    if (smem > 48 * 1024) e = cudaFuncSetAttribute(fir_wide_kernel<M>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); \
    if (e == cudaSuccess) fir_wide_kernel<M><<<grid, kWideThreads, smem, st>>>(a);                                      \
  } while (0)
  if (mode == 0) B2D_WIDE_LAUNCH(0);
  else if (mode == 1) B2D_WIDE_LAUNCH(1);
  else B2D_WIDE_LAUNCH(2);
#undef B2D_WIDE_LAUNCH
  return e != cudaSuccess ? e : cudaGetLastError();
}

}  // namespace b2d
