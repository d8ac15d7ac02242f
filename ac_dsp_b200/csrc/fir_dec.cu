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
#include <cstdlib>

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

// W: intermediate width, i128 or int64_t (common.cuh: fits_i64)
template <class W>
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
        acc1 = macc_t<W>(acc1, a.acc, (W)dec_sample(a, c, g - (long long)tp * a.DF) * (W)h[tp + a.NT * df], Fin + Fc);
      acc = macc_t<W>(acc, a.acc, (W)acc1, Fa);                    // acc = acc + acc1[df]
    }
    store_raw(a.y, (size_t)c * a.n_out + j, a.out_bytes, convert_t<W>((W)acc, Fa, a.out));
  }
}

// ------------------------------------------------------------------------------------------ 16-bit DP2A decimator
// polydec_q15_kernel: the decimator on the byte-plane DP2A arithmetic of fir_q15.cu (16-bit samples and taps, exact
// left-shift accumulator): the DDC partner of the R = 8 CIC decimator at the INT-pipe roofline.  Each phase plane is an
// ordinary FIR for the thread's 8 consecutive outputs: 3 LDS.128 of samples + 11 PRMT + 2 LDS.128 of broadcast taps per
// 256 DP2A, the int32 plane accumulators are flushed to int64 every <= 256 accumulated taps.
struct DecQArgs {
  const void *x;
  void *y;
  const void *tail;
  const uint32_t *cpk;    // [C][DF][pkw] byte-plane packed, reversed taps per phase (fir_q15_pack)
  size_t n, n_out;
  long long n_seen, m_first;
  int NT, Npad, pkw, DF;
  uint32_t C;
  int interleaved, lsh, out_bytes, fastout;
  Fmt acc, out;
};

__device__ __forceinline__ int dq_dp2a_lo(uint32_t a, uint32_t b, int c) {
  int d; asm("dp2a.lo.s32.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c)); return d;
}
__device__ __forceinline__ int dq_dp2a_hi(uint32_t a, uint32_t b, int c) {
  int d; asm("dp2a.hi.s32.s32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c)); return d;
}

template <int NP>
__global__ void __launch_bounds__(kWideThreads) polydec_q15_kernel(DecQArgs a) {
  extern __shared__ __align__(16) unsigned char smem[];
  const int XS = (kDecTile + a.Npad + 8 + 7) & ~7;               // samples per phase plane
  uint32_t *cw = (uint32_t *)smem;                                // [NP][DF][pkw]
  int16_t *xs = (int16_t *)(smem + (size_t)NP * a.DF * a.pkw * 4);   // [NP][DF][XS]: plane r holds u_r[m0 - (NT-1) + i]
  const uint32_t c0 = NP == 2 ? 0 : blockIdx.y;
  const long long m0 = a.m_first + (long long)blockIdx.x * kDecTile;
  const int T = a.NT * a.DF - 1;

  for (int i = threadIdx.x; i < NP * a.DF * a.pkw; i += kWideThreads) cw[i] = a.cpk[(size_t)c0 * a.DF * a.pkw + i];
  // Stage XS*DF consecutive samples per channel, de-interleaved by phase: sample q of the tile belongs to plane
  // r = DF-1 - q % DF at index q / DF (u_r[m] = x[m*DF + DF-1 - r]).  Consecutive threads take consecutive samples
  // (coalesced); (k, pos) = (q / DF, q % DF) advance incrementally, loads are issued in batches of 8 before their stores.
  const long long g_base = (m0 - (a.NT - 1)) * a.DF;
  const long long l0 = g_base - a.n_seen;                        // index of sample q = 0 in this call's input
  const int total = XS * a.DF;
  const bool interior = l0 >= 0 && (size_t)(l0 + total) <= a.n;
  const int inc_k = kWideThreads / a.DF, inc_pos = kWideThreads % a.DF;
  int k = threadIdx.x / a.DF, pos = threadIdx.x % a.DF;
  auto put = [&](uint32_t w) {                                   // store the sample(s) of the current (k, pos), then advance
    const int r = a.DF - 1 - pos;
    xs[(size_t)r * XS + k] = (int16_t)(w & 0xFFFF);
    if (NP == 2) xs[(size_t)(a.DF + r) * XS + k] = (int16_t)(w >> 16);
    k += inc_k; pos += inc_pos;
    if (pos >= a.DF) { pos -= a.DF; k++; }
  };
  auto fetch = [&](int q) -> uint32_t {                          // checked load: history / zero outside this call
    const long long li = l0 + q;
    if (NP == 2) {
      if (li >= 0) return (size_t)li < a.n ? ((const uint32_t *)a.x)[li] : 0u;
      if (li < -(long long)T) return 0u;
      return (uint32_t)((const uint16_t *)a.tail)[(size_t)(T + li)] | ((uint32_t)((const uint16_t *)a.tail)[(size_t)T + (size_t)(T + li)] << 16);
    }
    if (li >= 0) return (size_t)li < a.n ? (uint32_t)((const uint16_t *)a.x)[elem_index((size_t)li, c0, a.n, a.C, a.interleaved)] : 0u;
    if (li < -(long long)T) return 0u;
    return (uint32_t)((const uint16_t *)a.tail)[(size_t)c0 * T + (size_t)(T + li)];
  };
  const bool planar1 = NP == 1 && (!a.interleaved || a.C == 1);
  const int lg = (a.DF & (a.DF - 1)) == 0 ? 31 - __clz(a.DF) : -1;   // DF a power of two: (k, pos) by shift / mask, no chain
  if (interior && NP == 2 && lg >= 2 && ((((uintptr_t)((const uint32_t *)a.x + l0)) & 15) == 0)) {
    // IQ words, DF a multiple of 4: a 128-bit load is 4 consecutive phases of one row k; two rows (k, k+1) of the same phase
    // pack into one 32-bit shared-memory word per channel: 2 LDG.128 + 8 PRMT + 8 STS.32 per 8 IQ samples
    const uint4 *x128 = (const uint4 *)((const uint32_t *)a.x + l0);
    const int lg4 = lg - 2, G4 = a.DF >> 2, items = (XS >> 1) * G4, mask = a.DF - 1;
    for (int it = threadIdx.x; it < items; it += 4 * kWideThreads) {
      uint4 w0[4], w1[4];
#pragma unroll
      for (int u = 0; u < 4; u++) {
        const int ii = it + u * kWideThreads;
        if (ii < items) {
          const int kkp = ii >> lg4, pg = ii & (G4 - 1);
          w0[u] = x128[(size_t)(2 * kkp) * G4 + pg];
          w1[u] = x128[(size_t)(2 * kkp + 1) * G4 + pg];
        }
      }
#pragma unroll
      for (int u = 0; u < 4; u++) {
        const int ii = it + u * kWideThreads;
        if (ii < items) {
          const int kkp = ii >> lg4, pg = ii & (G4 - 1);
          const uint32_t a0[4] = {w0[u].x, w0[u].y, w0[u].z, w0[u].w}, a1[4] = {w1[u].x, w1[u].y, w1[u].z, w1[u].w};
#pragma unroll
          for (int i = 0; i < 4; i++) {
            const int r = mask - (4 * pg + i);
            ((uint32_t *)(xs + (size_t)r * XS))[kkp] = __byte_perm(a0[i], a1[i], 0x5410);
            ((uint32_t *)(xs + (size_t)(a.DF + r) * XS))[kkp] = __byte_perm(a0[i], a1[i], 0x7632);
          }
        }
      }
    }
  } else if (interior && (NP == 2 || planar1) && lg >= 0) {
    const uint32_t *x32 = (const uint32_t *)a.x + l0;
    const uint16_t *x16 = (const uint16_t *)a.x + (size_t)c0 * a.n + l0;
    const int mask = a.DF - 1;
    for (int q = threadIdx.x; q < total; q += 8 * kWideThreads) {
      uint32_t w[8];
#pragma unroll
      for (int u = 0; u < 8; u++) {
        const int qq = q + u * kWideThreads;
        w[u] = qq < total ? (NP == 2 ? x32[qq] : (uint32_t)x16[qq]) : 0u;
      }
#pragma unroll
      for (int u = 0; u < 8; u++) {
        const int qq = q + u * kWideThreads;
        if (qq < total) {
          const int kk = qq >> lg, r = mask - (qq & mask);
          xs[(size_t)r * XS + kk] = (int16_t)(w[u] & 0xFFFF);
          if (NP == 2) xs[(size_t)(a.DF + r) * XS + kk] = (int16_t)(w[u] >> 16);
        }
      }
    }
  } else if (interior && (NP == 2 || planar1)) {
    const uint32_t *x32 = (const uint32_t *)a.x + l0;
    const uint16_t *x16 = (const uint16_t *)a.x + (size_t)c0 * a.n + l0;
    for (int q = threadIdx.x; q < total; q += 8 * kWideThreads) {
      uint32_t w[8];
#pragma unroll
      for (int u = 0; u < 8; u++) {
        const int qq = q + u * kWideThreads;
        w[u] = qq < total ? (NP == 2 ? x32[qq] : (uint32_t)x16[qq]) : 0u;
      }
#pragma unroll
      for (int u = 0; u < 8; u++)
        if (q + u * kWideThreads < total) put(w[u]);
    }
  } else {
    for (int q = threadIdx.x; q < total; q += kWideThreads) put(fetch(q));
  }
  __syncthreads();

  const int o = threadIdx.x * kWideT;
  const long long j0 = m0 - a.m_first + o;
  if ((size_t)j0 >= a.n_out) return;
#pragma unroll
  for (int p = 0; p < NP; p++) {
    long long tot[kWideT];
    int lo[kWideT], hi[kWideT];
#pragma unroll
    for (int j = 0; j < kWideT; j++) { tot[j] = 0; lo[j] = 0; hi[j] = 0; }
    int pending = 0;                                             // taps accumulated in lo / hi since the last flush
    for (int r = 0; r < a.DF; r++) {
      if (pending + a.Npad > 256) {
#pragma unroll
        for (int j = 0; j < kWideT; j++) { tot[j] += ((long long)hi[j] << 8) + (long long)lo[j]; lo[j] = 0; hi[j] = 0; }
        pending = 0;
      }
      pending += a.Npad;
      const uint4 *x4 = (const uint4 *)(xs + (size_t)(p * a.DF + r) * XS + o);
      const uint4 *c4 = (const uint4 *)(cw + (size_t)(p * a.DF + r) * a.pkw);
#pragma unroll 1
      for (int k0 = 0; k0 < a.Npad; k0 += 16) {
        uint32_t E[12], O[11], cwv[8];
        const uint4 v0 = x4[k0 / 8], v1 = x4[k0 / 8 + 1], v2 = x4[k0 / 8 + 2];
        const uint4 q0 = c4[k0 / 8], q1 = c4[k0 / 8 + 1];
        E[0] = v0.x; E[1] = v0.y; E[2] = v0.z; E[3] = v0.w; E[4] = v1.x; E[5] = v1.y; E[6] = v1.z; E[7] = v1.w;
        E[8] = v2.x; E[9] = v2.y; E[10] = v2.z; E[11] = v2.w;
        cwv[0] = q0.x; cwv[1] = q0.y; cwv[2] = q0.z; cwv[3] = q0.w; cwv[4] = q1.x; cwv[5] = q1.y; cwv[6] = q1.z; cwv[7] = q1.w;
#pragma unroll
        for (int i = 0; i < 11; i++) O[i] = __byte_perm(E[i], E[i + 1], 0x5432);
#pragma unroll
        for (int q = 0; q < 8; q++) {
#pragma unroll
          for (int j = 0; j < kWideT; j++) {
            const uint32_t sw = (j & 1) ? O[q + j / 2] : E[q + j / 2];
            lo[j] = dq_dp2a_lo(sw, cwv[q], lo[j]);
            hi[j] = dq_dp2a_hi(sw, cwv[q], hi[j]);
          }
        }
      }
    }
    // ---- epilogue: planar outputs, 8 consecutive values of channel c0 + p
    const uint32_t c = c0 + p;
    long long res[kWideT];
#pragma unroll
    for (int j = 0; j < kWideT; j++) {
      const long long t = tot[j] + ((long long)hi[j] << 8) + (long long)lo[j];
      res[j] = wrap_bits((long long)((unsigned long long)t << a.lsh), a.acc.W, a.acc.S);
    }
    if (a.fastout) {
      long long *yc = (long long *)a.y + (size_t)c * a.n_out + j0;
      if ((size_t)(j0 + kWideT) <= a.n_out && (((uintptr_t)yc) & 15) == 0) {
#pragma unroll
        for (int j = 0; j < kWideT; j += 2) { longlong2 v; v.x = res[j]; v.y = res[j + 1]; *(longlong2 *)(yc + j) = v; }
      } else {
#pragma unroll
        for (int j = 0; j < kWideT; j++) if ((size_t)(j0 + j) < a.n_out) yc[j] = res[j];
      }
    } else {
#pragma unroll
      for (int j = 0; j < kWideT; j++)
        if ((size_t)(j0 + j) < a.n_out) store_raw(a.y, (size_t)c * a.n_out + (size_t)(j0 + j), a.out_bytes, convert((i128)res[j], a.acc.F(), a.out));
    }
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

static size_t decq_smem(int ntaps, int df, int np) {
  const int npad = (ntaps + 15) / 16 * 16;
  const size_t XS = (size_t)(kDecTile + npad + 8 + 7) & ~(size_t)7;
  return (size_t)np * df * (npad / 2) * 4 + (size_t)np * df * XS * 2;
}

// 16-bit DP2A path: signed 16-bit containers on both sides, no bits dropped per tap, wrapping accumulator
bool polydec_q15_supported(const Fmt &in, const Fmt &coeff, const Fmt &acc, int ntaps, int df) {
  if (in.W + (in.S ? 0 : 1) > 16 || coeff.W + (coeff.S ? 0 : 1) > 16) return false;
  if (acc.O != B2D_WRAP || (acc.Q != B2D_TRN && acc.Q != B2D_RND)) return false;
  const int lsh = acc.F() - in.F() - coeff.F();
  if (lsh < 0 || lsh > 40 || lsh >= acc.W) return false;
  if ((ntaps + 15) / 16 * 16 > 256) return false;
  return decq_smem(ntaps, df, 2) <= 200 * 1024;
}
int polydec_q15_words(int ntaps, int df) { return df * ((ntaps + 15) / 16 * 16 / 2); }
void polydec_q15_pack(const Fmt &coeff, const int64_t *c, int ntaps, int df, uint32_t *out) {
  const int pkw = (ntaps + 15) / 16 * 16 / 2;
  for (int r = 0; r < df; r++) fir_q15_pack(coeff, c + (size_t)ntaps * r, ntaps, B2D_SHIFT_REG, out + (size_t)r * pkw, pkw);
}

static cudaError_t launch_polydec_q15(const DecLaunch &p, cudaStream_t st) {
  DecQArgs a;
  a.x = p.in; a.y = p.out; a.tail = p.tail; a.cpk = p.coeff_pk; a.n = p.n; a.n_out = p.n_out;
  a.n_seen = (long long)p.n_seen; a.m_first = (long long)(p.n_seen / (unsigned long long)p.df);
  a.NT = p.nt; a.Npad = (p.nt + 15) / 16 * 16; a.pkw = a.Npad / 2; a.DF = p.df; a.C = p.C; a.interleaved = p.interleaved;
  a.lsh = p.facc.F() - p.fin.F() - p.fcoeff.F(); a.out_bytes = container_bytes(p.fout.W);
  a.fastout = (p.fout.W == p.facc.W && p.fout.I == p.facc.I && p.fout.S == p.facc.S && a.out_bytes == 8) ? 1 : 0;
  a.acc = p.facc; a.out = p.fout;
  const int np = (p.interleaved && p.C == 2) ? 2 : 1;
  const size_t smem = decq_smem(p.nt, p.df, np);
  dim3 grid((unsigned)((p.n_out + kDecTile - 1) / kDecTile), np == 2 ? 1 : p.C);
  cudaError_t e = cudaSuccess;
  if (np == 2) {
    if (smem > 48 * 1024) e = cudaFuncSetAttribute(polydec_q15_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e == cudaSuccess) polydec_q15_kernel<2><<<grid, kWideThreads, smem, st>>>(a);
  } else {
    if (smem > 48 * 1024) e = cudaFuncSetAttribute(polydec_q15_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e == cudaSuccess) polydec_q15_kernel<1><<<grid, kWideThreads, smem, st>>>(a);
  }
  return e != cudaSuccess ? e : cudaGetLastError();
}

// phase-order taps -> [DF][NTpad] int32, zero padded
void polydec_pack(const int64_t *c, int ntaps, int df, int32_t *out) {
  const int ntpad = polydec_words(ntaps);
  for (int r = 0; r < df; r++)
    for (int tp = 0; tp < ntpad; tp++) out[(size_t)r * ntpad + tp] = tp < ntaps ? (int32_t)c[tp + ntaps * r] : 0;
}

cudaError_t launch_polydec(const DecLaunch &p, cudaStream_t st) {
  if (p.n_out == 0) return cudaSuccess;
  if (p.wide == 2) return launch_polydec_q15(p, st);
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
    const char *f128 = getenv("B2D_GENERIC_I128");
    const int Wp = p.fin.W + (p.fin.S ? 0 : 1) + p.fcoeff.W + (p.fcoeff.S ? 0 : 1);
    if (!(f128 && *f128 == '1') && fits_i64(p.facc, p.fout, Wp, p.fin.F() + p.fcoeff.F()))
      polydec_generic_kernel<int64_t><<<(unsigned)blocks, 256, 0, st>>>(a);
    else
      polydec_generic_kernel<i128><<<(unsigned)blocks, 256, 0, st>>>(a);
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
