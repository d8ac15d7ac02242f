// fir_q15.cu -- the 16-bit FIR hot path (BASELINE configs 1, 2, 4: ac_fixed<16,1> x ac_fixed<16,1> -> <40,8>).
//
// Replaces the tap-MAC loop  acc += reg[i]*h[i]  of fir*ShiftReg / RotateShift / CircularBuff / Transposed and,
// through an effective-coefficient expansion, SymmetricEvenTaps / SymmetricOddTaps
// (reference include/ac_dsp/ac_fir_load_coeffs.h:180-278, ac_fir_const_coeffs.h:190-296, ac_fir_prog_coeffs.h:147-247).
//
// Exactness argument.  The path is taken only when s = F_in + F_c - F_acc <= 0, so the per-tap
// re-quantisation of `acc += a*b` is an exact left shift, and ACC_TYPE wraps (AC_WRAP): then
//   acc_raw = wrap_Wacc( (sum_i x[n-i]*h[i]) << -s )
// whatever the tap order or architecture -- an integer dot product, evaluated here exactly.
//
// Arithmetic.  16x16 products overflow an int32 accumulator after two taps and IMAD.WIDE issues at less than half
// the IMAD rate (profiles/r01_ubench_pipes.jsonl), so the coefficient is split into byte planes
// h = 256*hh + hl (hl unsigned, hh signed) and each plane is accumulated with DP2A (IDP.2A: two 16b x 8b
// products per lane per instruction).  |x*hl| < 2^23, |x*hh| <= 2^22: 256 taps fit an int32 per plane
// exactly; longer filters flush to int64 every 256 taps (128 for unsigned samples).
//
// Data movement.  One CTA = one tile of consecutive outputs of one channel (both channels of an interleaved
// IQ pair).  Samples + (N_TAPS-1) of history are staged once into shared memory as planar int16, coefficients
// as reversed packed words {hl[k], hl[k+1], hh[k], hh[k+1]} so that sample and tap index advance together.
// A thread owns 8 consecutive outputs; per 16 taps it slides a 24-sample register window (3 LDS.128, the
// odd-aligned pairs built with PRMT) and reads 8 coefficient words as a shared-memory broadcast (2 LDS.128):
// 5 loads + 11 PRMT feed 128 DP2A.
#include <vector>

#include "kernels.h"

namespace b2d {

constexpr int kThreads = 128;
constexpr int kT = 8;          // outputs per thread per pass
constexpr int kChunk = 16;     // taps per unrolled chunk
constexpr int kMaxTapsQ15 = 2048;

struct Q15Args {
  const void *x;
  void *y;
  const void *tail;
  const uint32_t *cpk;   // [C][pkw]
  size_t n;
  int N, Npad, pkw, passes;
  uint32_t C;
  int interleaved;
  int lsh;               // -s: exact left shift of the dot product
  Fmt acc, out;
  int out_bytes;
  int vec_ok;            // in/out base pointers 16-byte aligned
};

template <int XS>
__device__ __forceinline__ int dp2a_lo_u8(uint32_t a, uint32_t b, int c) {
  int d;
  if (XS) asm("dp2a.lo.s32.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
  else asm("dp2a.lo.u32.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
  return d;
}
template <int XS, int CS>
__device__ __forceinline__ int dp2a_hi_b8(uint32_t a, uint32_t b, int c) {
  int d;
  if (XS && CS) asm("dp2a.hi.s32.s32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
  else if (XS && !CS) asm("dp2a.hi.s32.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
  else if (!XS && CS) asm("dp2a.hi.u32.s32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
  else asm("dp2a.hi.u32.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
  return d;
}

// Stage samples [g0, g0 + count) of plane (channel) `c` into xs[0..count): history for negative indices, zeros past n.
template <int NP>
__device__ __forceinline__ void stage_tile(const Q15Args &a, uint32_t c0, long long g0, int count, int stride, int16_t *xs) {
  const int T = a.N - 1;
  const uint16_t *x16 = (const uint16_t *)a.x;
  const uint16_t *t16 = (const uint16_t *)a.tail;
  if (NP == 2) {
    // interleaved IQ: one 32-bit word = (I, Q); 128-bit loads carry 4 pairs
    const uint32_t *x32 = (const uint32_t *)a.x;
    uint16_t *p0 = (uint16_t *)xs, *p1 = (uint16_t *)xs + stride;
    // head: history / unaligned part, element-wise
    int s = threadIdx.x;
    long long head_end = g0 < 0 ? 0 : g0;
    head_end = (head_end + 3) & ~3LL;                 // first 4-pair aligned global index >= max(g0, 0)
    int head = (int)(head_end - g0);
    if (head > count) head = count;
    for (; s < head; s += kThreads) {
      const long long g = g0 + s;
      uint16_t vi = 0, vq = 0;
      if (g < 0) { vi = t16[(size_t)0 * T + (size_t)(T + g)]; vq = t16[(size_t)1 * T + (size_t)(T + g)]; }
      else if ((size_t)g < a.n) { const uint32_t w = x32[g]; vi = (uint16_t)w; vq = (uint16_t)(w >> 16); }
      p0[s] = vi; p1[s] = vq;
    }
    // body: groups of 4 pairs
    const int groups = (count - head) / 4;
    for (int q = threadIdx.x; q < groups; q += kThreads) {
      const long long g = head_end + 4LL * q;
      const int so = head + 4 * q;
      uint4 w = make_uint4(0, 0, 0, 0);
      if (a.vec_ok && (size_t)(g + 4) <= a.n) w = *(const uint4 *)(x32 + g);
      else {
        if ((size_t)g < a.n) w.x = x32[g];
        if ((size_t)(g + 1) < a.n) w.y = x32[g + 1];
        if ((size_t)(g + 2) < a.n) w.z = x32[g + 2];
        if ((size_t)(g + 3) < a.n) w.w = x32[g + 3];
      }
      p0[so] = (uint16_t)w.x; p1[so] = (uint16_t)(w.x >> 16);
      p0[so + 1] = (uint16_t)w.y; p1[so + 1] = (uint16_t)(w.y >> 16);
      p0[so + 2] = (uint16_t)w.z; p1[so + 2] = (uint16_t)(w.z >> 16);
      p0[so + 3] = (uint16_t)w.w; p1[so + 3] = (uint16_t)(w.w >> 16);
    }
    for (s = head + 4 * groups + threadIdx.x; s < count; s += kThreads) {
      const long long g = g0 + s;
      uint16_t vi = 0, vq = 0;
      if ((size_t)g < a.n) { const uint32_t w = x32[g]; vi = (uint16_t)w; vq = (uint16_t)(w >> 16); }
      p0[s] = vi; p1[s] = vq;
    }
  } else {
    uint16_t *p0 = (uint16_t *)xs;
    const bool planar = !a.interleaved || a.C == 1;
    const uint16_t *xc = planar ? x16 + (size_t)c0 * a.n : x16;
    long long head_end = g0 < 0 ? 0 : g0;
    head_end = (head_end + 7) & ~7LL;
    int head = (int)(head_end - g0);
    if (head > count || !planar || !a.vec_ok || ((a.n & 7) && c0)) head = count;  // element-wise everywhere
    for (int s = threadIdx.x; s < head; s += kThreads) {
      const long long g = g0 + s;
      uint16_t v = 0;
      if (g < 0) v = t16[(size_t)c0 * T + (size_t)(T + g)];
      else if ((size_t)g < a.n) v = planar ? xc[g] : x16[(size_t)g * a.C + c0];
      p0[s] = v;
    }
    const int groups = (count - head) / 8;
    for (int q = threadIdx.x; q < groups; q += kThreads) {
      const long long g = head_end + 8LL * q;
      const int so = head + 8 * q;
      if ((size_t)(g + 8) <= a.n) {
        const uint4 w = *(const uint4 *)(xc + g);
        p0[so] = (uint16_t)w.x; p0[so + 1] = (uint16_t)(w.x >> 16);
        p0[so + 2] = (uint16_t)w.y; p0[so + 3] = (uint16_t)(w.y >> 16);
        p0[so + 4] = (uint16_t)w.z; p0[so + 5] = (uint16_t)(w.z >> 16);
        p0[so + 6] = (uint16_t)w.w; p0[so + 7] = (uint16_t)(w.w >> 16);
      } else {
        for (int e = 0; e < 8; e++) p0[so + e] = (size_t)(g + e) < a.n ? xc[g + e] : (uint16_t)0;
      }
    }
    for (int s = head + 8 * groups + threadIdx.x; s < count; s += kThreads) {
      const long long g = g0 + s;
      p0[s] = (size_t)g < a.n ? xc[g] : (uint16_t)0;
    }
  }
}

// XS / CS: samples / coefficients signed.  NP: planes per CTA (2 = interleaved IQ pair).
// FASTOUT: OUT_TYPE == ACC_TYPE in an int64 container (the BASELINE configs): no conversion, 128-bit stores.
// ONE: the padded tap count fits one int32 accumulation block (<= 256 taps, 128 for unsigned samples): no running 64-bit
// totals, 16 registers fewer.
template <int XS, int CS, int NP, bool FASTOUT, bool ONE>
__global__ void __launch_bounds__(kThreads) fir_q15_kernel(Q15Args a) {
  extern __shared__ __align__(16) unsigned char smem[];
  const int tile = kThreads * kT * a.passes;
  const int stride = ((tile + a.Npad + 8 + 7) & ~7);             // samples per plane in smem (16-byte multiple)
  uint32_t *cw = (uint32_t *)smem;                                // [NP][pkw]
  int16_t *xs = (int16_t *)(smem + (size_t)NP * a.pkw * 4);       // [NP][stride]
  const uint32_t c0 = NP == 2 ? 0 : blockIdx.y;
  const long long out0 = (long long)blockIdx.x * tile;

  for (int i = threadIdx.x; i < NP * a.pkw; i += kThreads) cw[i] = a.cpk[(size_t)c0 * a.pkw + i];
  stage_tile<NP>(a, c0, out0 - (a.N - 1), stride, stride, xs);
  __syncthreads();

  constexpr int KC = XS ? 256 : 128;  // taps per int32 accumulation block
  for (int pass = 0; pass < a.passes; pass++) {
    const int o = (pass * kThreads + threadIdx.x) * kT;
    const long long n0 = out0 + o;
    if ((size_t)n0 >= a.n) break;
    long long res[NP][kT];
#pragma unroll
    for (int p = 0; p < NP; p++) {
      const uint4 *x4 = (const uint4 *)(xs + (size_t)p * stride + o);
      const uint4 *c4 = (const uint4 *)(cw + (size_t)p * a.pkw);
      long long tot[kT];
      if (!ONE) {
#pragma unroll
        for (int j = 0; j < kT; j++) tot[j] = 0;
      }
      for (int kb = 0; kb < (ONE ? 1 : a.Npad); kb += KC) {
        int lo[kT], hi[kT];
#pragma unroll
        for (int j = 0; j < kT; j++) { lo[j] = 0; hi[j] = 0; }
        const int kend = ONE ? a.Npad : (kb + KC < a.Npad ? kb + KC : a.Npad);
        // four chunks per trip: the loads use immediate offsets and the trip's address arithmetic (three IMAD.IADD on the
        // DP2A pipe per chunk before) is paid once per 512 DP2A; r02 A/B on a B200: 33.27 -> 34.5 G IQ samples/s at 256 taps
        // (unroll 2: 33.98, 8: 28.9, 16: 27.1 -- the loop body outgrows the instruction cache)
#pragma unroll 4
        for (int k0 = kb; k0 < kend; k0 += kChunk) {
          uint32_t E[12], O[11], cwv[8];
          const uint4 v0 = x4[k0 / 8], v1 = x4[k0 / 8 + 1], v2 = x4[k0 / 8 + 2];
          const uint4 q0 = c4[k0 / 8], q1 = c4[k0 / 8 + 1];
          E[0] = v0.x; E[1] = v0.y; E[2] = v0.z; E[3] = v0.w;
          E[4] = v1.x; E[5] = v1.y; E[6] = v1.z; E[7] = v1.w;
          E[8] = v2.x; E[9] = v2.y; E[10] = v2.z; E[11] = v2.w;
          cwv[0] = q0.x; cwv[1] = q0.y; cwv[2] = q0.z; cwv[3] = q0.w;
          cwv[4] = q1.x; cwv[5] = q1.y; cwv[6] = q1.z; cwv[7] = q1.w;
#pragma unroll
          for (int i = 0; i < 11; i++) O[i] = __byte_perm(E[i], E[i + 1], 0x5432);
#pragma unroll
          for (int q = 0; q < 8; q++) {
#pragma unroll
            for (int j = 0; j < kT; j++) {
              const uint32_t s = (j & 1) ? O[q + j / 2] : E[q + j / 2];
              lo[j] = dp2a_lo_u8<XS>(s, cwv[q], lo[j]);
              hi[j] = dp2a_hi_b8<XS, CS>(s, cwv[q], hi[j]);
            }
          }
        }
#pragma unroll
        for (int j = 0; j < kT; j++) {
          const long long l = XS ? (long long)lo[j] : (long long)(unsigned)lo[j];
          const long long h = (XS || CS) ? (long long)hi[j] : (long long)(unsigned)hi[j];
          if (ONE) tot[j] = (h << 8) + l; else tot[j] += (h << 8) + l;
        }
      }
#pragma unroll
      for (int j = 0; j < kT; j++) res[p][j] = wrap_bits((long long)((unsigned long long)tot[j] << a.lsh), a.acc.W, a.acc.S);
    }
    // ---- epilogue
    if (FASTOUT) {
      long long *y = (long long *)a.y;
      if (NP == 2) {
        if (a.vec_ok && (size_t)(n0 + kT) <= a.n) {
#pragma unroll
          for (int j = 0; j < kT; j++) {
            longlong2 v; v.x = res[0][j]; v.y = res[NP - 1][j];
            *(longlong2 *)(y + 2 * (size_t)(n0 + j)) = v;
          }
        } else {
#pragma unroll
          for (int j = 0; j < kT; j++)
            if ((size_t)(n0 + j) < a.n) { y[2 * (size_t)(n0 + j)] = res[0][j]; y[2 * (size_t)(n0 + j) + 1] = res[NP - 1][j]; }
        }
      } else {
        const bool planar = !a.interleaved || a.C == 1;
        if (planar && a.vec_ok && (size_t)(n0 + kT) <= a.n && !((a.n & 1) && c0)) {
          long long *yc = y + (size_t)c0 * a.n + n0;
#pragma unroll
          for (int j = 0; j < kT; j += 2) {
            longlong2 v; v.x = res[0][j]; v.y = res[0][j + 1];
            *(longlong2 *)(yc + j) = v;
          }
        } else {
#pragma unroll
          for (int j = 0; j < kT; j++)
            if ((size_t)(n0 + j) < a.n) y[elem_index((size_t)(n0 + j), c0, a.n, a.C, a.interleaved)] = res[0][j];
        }
      }
    } else {
#pragma unroll
      for (int p = 0; p < NP; p++)
#pragma unroll
        for (int j = 0; j < kT; j++)
          if ((size_t)(n0 + j) < a.n)
            store_raw(a.y, elem_index((size_t)(n0 + j), NP == 2 ? (uint32_t)p : c0, a.n, a.C, a.interleaved), a.out_bytes,
                      convert((i128)res[p][j], a.acc.F(), a.out));
    }
  }
}

// ------------------------------------------------------------------------------------------ host side
bool fir_q15_supported(const Fmt &in, const Fmt &coeff, const Fmt &acc, const Fmt &out, int n_taps, int ftype) {
  (void)out;
  if (in.W > 16 || coeff.W > 16) return false;
  if (acc.O != B2D_WRAP || (acc.Q != B2D_TRN && acc.Q != B2D_RND)) return false;
  const int lsh = acc.F() - in.F() - coeff.F();
  if (lsh < 0 || lsh > 40) return false;
  if (n_taps > kMaxTapsQ15) return false;
  switch (ftype) {
    case B2D_SHIFT_REG: case B2D_ROTATE_SHIFT: case B2D_C_BUFF: case B2D_TRANSPOSED: case B2D_FOLD_EVEN: return true;
    case B2D_FOLD_EVEN_ANTI:   // mirrored taps are negated: they must still fit the signed 16-bit byte planes
      return coeff.W + (coeff.S ? 0 : 1) <= 15;
    case B2D_FOLD_ODD_ANTI:
      if (coeff.W + (coeff.S ? 0 : 1) > 15) return false;
      // fall through
    case B2D_FOLD_ODD:
      // `fold` is ACC_TYPE (ac_fir_load_coeffs.h:248-255): exact only if the pre-add neither truncates nor wraps there;
      // an unsigned ACC_TYPE wraps every negative pre-add (signed samples, or the _ANTI pre-subtract) before the multiply
      if (!acc.S && (in.S || ftype == B2D_FOLD_ODD_ANTI)) return false;
      return acc.F() >= in.F() && in.W + 1 + (in.S ? 0 : 1) + (acc.F() - in.F()) <= acc.W + (acc.S ? 0 : 1);
    default: return false;
  }
}

int fir_q15_pk_words(int n_taps, int ftype) {
  (void)ftype;
  const int npad = (n_taps + kChunk - 1) / kChunk * kChunk;
  return npad / 2;
}

// Effective direct-form coefficients of the folded architectures, reversed, byte-plane packed.
void fir_q15_pack(const Fmt &coeff, const int64_t *c, int n_taps, int ftype, uint32_t *pk, int pk_words) {
  (void)coeff;
  const int N = n_taps;
  std::vector<int64_t> eff(N, 0);
  const int64_t sg = (ftype == B2D_FOLD_EVEN_ANTI || ftype == B2D_FOLD_ODD_ANTI) ? -1 : 1;   // ac_fir_reg_share.h:151-165,186-205
  if (ftype == B2D_FOLD_EVEN || ftype == B2D_FOLD_EVEN_ANTI) {   // ac_fir_load_coeffs.h:231-239: taps i and N-1-i share h[i], i < N/2
    for (int i = 0; i < N / 2; i++) { eff[i] = c[i]; eff[N - 1 - i] = sg * c[i]; }
  } else if (ftype == B2D_FOLD_ODD || ftype == B2D_FOLD_ODD_ANTI) {   // :246-259: i <= (N-1)/2, the last one unpaired
    for (int i = 0; i < (N - 1) / 2 + 1; i++) {
      eff[i] = c[i];
      if (i != (N - 1) / 2) eff[N - 1 - i] = sg * c[i];
    }
  } else {
    for (int i = 0; i < N; i++) eff[i] = c[i];
  }
  for (int w = 0; w < pk_words; w++) {
    uint32_t word = 0;
    for (int e = 0; e < 2; e++) {
      const int k = 2 * w + e;                       // reversed index: g[k] = h[N-1-k], zero padded past N
      const int64_t g = k < N ? eff[N - 1 - k] : 0;
      const uint32_t lo = (uint32_t)(g & 0xFF);
      const uint32_t hi = (uint32_t)((g >> 8) & 0xFF);  // arithmetic shift: signed high byte (unsigned formats: 0..255)
      word |= lo << (8 * e);
      word |= hi << (16 + 8 * e);
    }
    pk[w] = word;
  }
}

template <int XS, int CS, int NP, bool FASTOUT, bool ONE>
static cudaError_t launch_variant1(const Q15Args &a, dim3 grid, size_t smem, cudaStream_t st) {
  if (smem > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(fir_q15_kernel<XS, CS, NP, FASTOUT, ONE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
  }
  static bool carved = false;    // per instantiation: ask for the largest shared-memory carve-out once (the tiles, not L1, hold the working set)
  if (!carved) {
    cudaFuncSetAttribute(fir_q15_kernel<XS, CS, NP, FASTOUT, ONE>, cudaFuncAttributePreferredSharedMemoryCarveout, (int)cudaSharedmemCarveoutMaxShared);
    carved = true;
  }
  fir_q15_kernel<XS, CS, NP, FASTOUT, ONE><<<grid, kThreads, smem, st>>>(a);
  return cudaGetLastError();
}
template <int XS, int CS, int NP, bool FASTOUT>
static cudaError_t launch_variant(const Q15Args &a, dim3 grid, size_t smem, cudaStream_t st) {
  const bool one = a.Npad <= (XS ? 256 : 128);
  return one ? launch_variant1<XS, CS, NP, FASTOUT, true>(a, grid, smem, st) : launch_variant1<XS, CS, NP, FASTOUT, false>(a, grid, smem, st);
}

template <int XS, int CS>
static cudaError_t launch_sc(const Q15Args &a, int np, bool fastout, dim3 grid, size_t smem, cudaStream_t st) {
  if (np == 2) return fastout ? launch_variant<XS, CS, 2, true>(a, grid, smem, st) : launch_variant<XS, CS, 2, false>(a, grid, smem, st);
  return fastout ? launch_variant<XS, CS, 1, true>(a, grid, smem, st) : launch_variant<XS, CS, 1, false>(a, grid, smem, st);
}

cudaError_t launch_fir_q15(const FirLaunch &p, cudaStream_t st) {
  if (p.n == 0) return cudaSuccess;
  Q15Args a;
  a.x = p.in; a.y = p.out; a.tail = p.tail; a.cpk = p.coeff_pk; a.n = p.n;
  a.N = p.n_taps; a.pkw = p.pk_words; a.Npad = p.pk_words * 2;
  a.C = p.C; a.interleaved = p.interleaved;
  a.lsh = p.facc.F() - p.fin.F() - p.fcoeff.F();
  a.acc = p.facc; a.out = p.fout; a.out_bytes = container_bytes(p.fout.W);
  a.vec_ok = (((uintptr_t)p.in | (uintptr_t)p.out) & 15) == 0;
  const int np = (p.interleaved && p.C == 2) ? 2 : 1;
  const bool fastout = p.fout.W == p.facc.W && p.fout.I == p.facc.I && p.fout.S == p.facc.S && a.out_bytes == 8;
  const size_t per_pass = (size_t)kThreads * kT;
  // tile = 1..4 passes of 1024 outputs: four amortise the staged halo best (the tuned shape of the long calls), but a CTA walks
  // its passes one after the other, so a call too short to fill the GPU with four-pass tiles (25 CTAs for 10^5 samples: a
  // 30 us wave at 256 taps whatever the length, profiles/r02_ovs_crossover.txt) takes fewer passes per tile instead
  const size_t ctas_per_wave = 148 * 9;
  const size_t chans = (p.interleaved && p.C == 2) ? 1 : p.C;
  size_t passes = (p.n * chans + per_pass * ctas_per_wave - 1) / (per_pass * ctas_per_wave);
  if (passes > 4) passes = 4;
  if (passes < 1) passes = 1;
  a.passes = (int)passes;
  const size_t tile = per_pass * passes;
  const size_t stride = (tile + a.Npad + 8 + 7) & ~(size_t)7;
  const size_t smem = (size_t)np * a.pkw * 4 + (size_t)np * stride * 2;
  dim3 grid((unsigned)((p.n + tile - 1) / tile), np == 2 ? 1 : p.C);
  // the mirrored taps of the _ANTI folds are negated by fir_q15_pack: the effective taps are signed whatever COEFF_TYPE is
  const bool anti = p.ftype == B2D_FOLD_EVEN_ANTI || p.ftype == B2D_FOLD_ODD_ANTI;
  const int xs = p.fin.S ? 1 : 0, cs = (p.fcoeff.S || anti) ? 1 : 0;
  if (xs && cs) return launch_sc<1, 1>(a, np, fastout, grid, smem, st);
  if (xs && !cs) return launch_sc<1, 0>(a, np, fastout, grid, smem, st);
  if (!xs && cs) return launch_sc<0, 1>(a, np, fastout, grid, smem, st);
  return launch_sc<0, 0>(a, np, fastout, grid, smem, st);
}

}  // namespace b2d
