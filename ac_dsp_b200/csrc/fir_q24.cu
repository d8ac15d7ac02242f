// fir_q24.cu -- FIR on samples of up to 24 bits with coefficients of up to 16 bits (BASELINE configs[4] second stage
// unfused: the CIC interpolator's ac_fixed<20,5> stream x ac_fixed<16,1> taps -> <40,8>; any IN_TYPE of 17..24 bits).
//
// Replaces the same tap-MAC loops as fir_q15.cu (reference include/ac_dsp/ac_fir_load_coeffs.h:180-278,
// ac_fir_const_coeffs.h:190-296, ac_fir_prog_coeffs.h:147-247) when the samples no longer fit the 16-bit DP2A lane.
// Round 1 ran these formats on IMAD.WIDE (fir_wide.cu: 26 lanes/clk/SM, 2.4 pipe slots per MAC); here the roles of the
// DP2A operands are swapped with respect to fir_q15: the COEFFICIENT pair sits in the two 16-bit lanes and the SAMPLE is
// split into three byte planes x = x0 + 256 x1 + 65536 x2 (x0, x1 unsigned, x2 signed), each accumulated exactly in
// int32 (|h * byte| < 2^23: 128 taps per block, then flushed to int64): 1.5 pipe slots per MAC at 64 lanes/clk/SM.
//
// Exactness: taken only when s = F_in + F_c - F_acc <= 0 and ACC_TYPE wraps (see fir_q15.cu) -- an exact integer dot
// product whatever the tap order; the folded architectures become effective direct-form taps.
//
// Data movement: a CTA stages tile + N_TAPS - 1 samples of one channel as three byte arrays; a 32-bit word of a plane
// holds the bytes of four consecutive samples, so one word feeds dp2a.lo (bytes 0, 1) and dp2a.hi (bytes 2, 3) of two
// tap pairs; odd sample offsets use a copy shifted by one byte built in registers (PRMT).  A thread owns 8 consecutive
// outputs; per 16 taps and plane: 2 LDS.128 of samples + 6 PRMT + the broadcast coefficient words feed 64 DP2A.
#include <vector>

#include "kernels.h"

namespace b2d {

constexpr int kQ24Threads = 128;
constexpr int kQ24T = 8;            // outputs per thread per pass
constexpr int kQ24Chunk = 16;       // taps per unrolled chunk
constexpr int kQ24MaxTaps = 2048;
constexpr int kQ24Block = 128;      // taps per int32 accumulation block

struct Q24Args {
  const void *x;
  void *y;
  const void *tail;
  const uint32_t *cpk;   // [C][pkw] coefficient pair words, reversed taps
  size_t n;
  int N, Npad, pkw, passes;
  uint32_t C;
  int interleaved;
  int lsh;
  Fmt in, acc, out;
  int out_bytes, fastout;
  int vec_ok;            // planar input, 16-byte aligned base
};

__device__ __forceinline__ int q24_dp2a_lo_u(uint32_t a, uint32_t b, int c) { int d; asm("dp2a.lo.s32.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c)); return d; }
__device__ __forceinline__ int q24_dp2a_hi_u(uint32_t a, uint32_t b, int c) { int d; asm("dp2a.hi.s32.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c)); return d; }
__device__ __forceinline__ int q24_dp2a_lo_s(uint32_t a, uint32_t b, int c) { int d; asm("dp2a.lo.s32.s32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c)); return d; }
__device__ __forceinline__ int q24_dp2a_hi_s(uint32_t a, uint32_t b, int c) { int d; asm("dp2a.hi.s32.s32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c)); return d; }

// sample at global index g of channel c (history for g < 0, zero past the end)
__device__ __forceinline__ int q24_sample(const Q24Args &a, uint32_t c, long long g) {
  const int T = a.N - 1;
  if (g < 0) return g >= -(long long)T ? ((const int *)a.tail)[(size_t)c * T + (size_t)(T + g)] : 0;   // older than the history: meets a zero tap
  if ((size_t)g >= a.n) return 0;
  return ((const int *)a.x)[elem_index((size_t)g, c, a.n, a.C, a.interleaved)];
}

// LEAD = (-(N_TAPS - 1)) mod 4: where the first staged sample of a tile sits inside its 16-byte group (tiles are multiples
// of 1024 outputs, so it is the same for every tile of a launch) -- a template parameter, so that picking four samples out
// of two aligned 128-bit loads costs no moves.
// ONE: the padded tap count fits one int32 accumulation block (<= 128 taps): no running 64-bit totals.
template <int LEAD, bool ONE>
__global__ void __launch_bounds__(kQ24Threads) fir_q24_kernel(Q24Args a) {
  extern __shared__ __align__(16) unsigned char smem[];
  const int tile = kQ24Threads * kQ24T * a.passes;
  const int stride = (tile + a.Npad + 16 + 15) & ~15;                // bytes per plane (16-byte multiple)
  uint32_t *cw = (uint32_t *)smem;                                   // [pkw]
  unsigned char *pl = smem + (size_t)((a.pkw + 3) & ~3) * 4;         // [3][stride]
  const uint32_t c = blockIdx.y;
  const long long out0 = (long long)blockIdx.x * tile;
  const long long g0 = out0 - (a.N - 1);
  constexpr int lead = LEAD;

  for (int i = threadIdx.x; i < a.pkw; i += kQ24Threads) cw[i] = a.cpk[(size_t)c * a.pkw + i];
  // ---- stage: groups of four samples -> one word per plane
  const bool planar = !a.interleaved || a.C == 1;
  const int *xc = (const int *)a.x + (planar ? (size_t)c * a.n : 0);
  for (int q = threadIdx.x; q < stride / 4; q += kQ24Threads) {
    const long long g = g0 + 4LL * q;
    int v0, v1, v2, v3;
    const long long ga = g - lead;                                    // 16-byte aligned: the four samples straddle two 128-bit loads
    if (planar && ga >= 0 && (size_t)(ga + 8) <= a.n && a.vec_ok && (((size_t)c * a.n) & 3) == 0) {
      const int4 w0 = *(const int4 *)(xc + ga);
      if (lead == 0) { v0 = w0.x; v1 = w0.y; v2 = w0.z; v3 = w0.w; }
      else {
        const int4 w1 = *(const int4 *)(xc + ga + 4);                 // the neighbouring thread's first load: an L1 hit
        if (lead == 1) { v0 = w0.y; v1 = w0.z; v2 = w0.w; v3 = w1.x; }
        else if (lead == 2) { v0 = w0.z; v1 = w0.w; v2 = w1.x; v3 = w1.y; }
        else { v0 = w0.w; v1 = w1.x; v2 = w1.y; v3 = w1.z; }
      }
    } else {
      v0 = q24_sample(a, c, g); v1 = q24_sample(a, c, g + 1); v2 = q24_sample(a, c, g + 2); v3 = q24_sample(a, c, g + 3);
    }
    const uint32_t lo01 = __byte_perm(v0, v1, 0x5410), lo23 = __byte_perm(v2, v3, 0x5410);   // bytes (v0.b0 v0.b1 v1.b0 v1.b1)
    const uint32_t hi01 = __byte_perm(v0, v1, 0x7632), hi23 = __byte_perm(v2, v3, 0x7632);   // bytes (v0.b2 v0.b3 v1.b2 v1.b3)
    ((uint32_t *)pl)[q] = __byte_perm(lo01, lo23, 0x6420);                                   // plane 0: byte 0 of the four samples
    ((uint32_t *)(pl + stride))[q] = __byte_perm(lo01, lo23, 0x7531);                        // plane 1: byte 1
    ((uint32_t *)(pl + 2 * stride))[q] = __byte_perm(hi01, hi23, 0x6420);                    // plane 2: byte 2 (signed top byte of a <= 24-bit value)
  }
  __syncthreads();

  for (int pass = 0; pass < a.passes; pass++) {
    const int o = (pass * kQ24Threads + threadIdx.x) * kQ24T;        // staged byte offset of sample x[n0 - (N-1)]: window of output j, tap pair k starts at o + j + k
    const long long n0 = out0 + o;
    if ((size_t)n0 >= a.n) break;
    long long tot[kQ24T];
    if (!ONE) {
#pragma unroll
      for (int j = 0; j < kQ24T; j++) tot[j] = 0;
    }
    for (int kb = 0; kb < (ONE ? 1 : a.Npad); kb += kQ24Block) {
      int acc[3][kQ24T];
#pragma unroll
      for (int p = 0; p < 3; p++)
#pragma unroll
        for (int j = 0; j < kQ24T; j++) acc[p][j] = 0;
      const int kend = ONE ? a.Npad : (kb + kQ24Block < a.Npad ? kb + kQ24Block : a.Npad);
#pragma unroll 1
      for (int k0 = kb; k0 < kend; k0 += kQ24Chunk) {
        uint32_t cwv[8];
        {
          const uint4 q0 = *(const uint4 *)(cw + k0 / 2), q1 = *(const uint4 *)(cw + k0 / 2 + 4);
          cwv[0] = q0.x; cwv[1] = q0.y; cwv[2] = q0.z; cwv[3] = q0.w; cwv[4] = q1.x; cwv[5] = q1.y; cwv[6] = q1.z; cwv[7] = q1.w;
        }
#pragma unroll
        for (int p = 0; p < 3; p++) {
          uint32_t E[8], O[6];
          {
            const uint2 *w2 = (const uint2 *)(pl + (size_t)p * stride + o + k0);   // o is a multiple of 8, k0 and stride of 16
            const uint2 h0 = w2[0], h1 = w2[1], h2 = w2[2], h3 = w2[3];
            E[0] = h0.x; E[1] = h0.y; E[2] = h1.x; E[3] = h1.y; E[4] = h2.x; E[5] = h2.y; E[6] = h3.x; E[7] = h3.y;
          }
#pragma unroll
          for (int w = 0; w < 6; w++) O[w] = __byte_perm(E[w], E[w + 1], 0x4321);   // the same bytes one sample later
#pragma unroll
          for (int q = 0; q < 8; q++) {                 // tap pair (k0 + 2q, k0 + 2q + 1)
#pragma unroll
            for (int j = 0; j < kQ24T; j++) {
              const int t = j + 2 * q;                  // byte offset of the pair's first sample in the window
              const int te = (j & 1) ? t - 1 : t;       // even offset into E (even j) or O (odd j)
              const uint32_t s = (j & 1) ? O[te / 4] : E[te / 4];
              if (p < 2) acc[p][j] = (te & 2) ? q24_dp2a_hi_u(cwv[q], s, acc[p][j]) : q24_dp2a_lo_u(cwv[q], s, acc[p][j]);
              else acc[p][j] = (te & 2) ? q24_dp2a_hi_s(cwv[q], s, acc[p][j]) : q24_dp2a_lo_s(cwv[q], s, acc[p][j]);
            }
          }
        }
      }
#pragma unroll
      for (int j = 0; j < kQ24T; j++) {
        const long long blk = (long long)acc[0][j] + ((long long)acc[1][j] << 8) + ((long long)acc[2][j] << 16);
        if (ONE) tot[j] = blk; else tot[j] += blk;
      }
    }
    long long res[kQ24T];
#pragma unroll
    for (int j = 0; j < kQ24T; j++) res[j] = wrap_bits((long long)((unsigned long long)tot[j] << a.lsh), a.acc.W, a.acc.S);
    if (a.fastout && planar && a.vec_ok && (size_t)(n0 + kQ24T) <= a.n && (((size_t)c * a.n) & 1) == 0) {
      long long *yc = (long long *)a.y + (size_t)c * a.n + n0;          // n0 is a multiple of 8: 64-byte aligned
#pragma unroll
      for (int j = 0; j < kQ24T; j += 2) *(longlong2 *)(yc + j) = make_longlong2(res[j], res[j + 1]);
    } else {
#pragma unroll
      for (int j = 0; j < kQ24T; j++) {
        if ((size_t)(n0 + j) >= a.n) break;
        const size_t idx = elem_index((size_t)(n0 + j), c, a.n, a.C, a.interleaved);
        if (a.fastout) ((long long *)a.y)[idx] = res[j];
        else store_raw(a.y, idx, a.out_bytes, convert((i128)res[j], a.acc.F(), a.out));
      }
    }
  }
}

// ------------------------------------------------------------------------------------------ host side
bool fir_q24_supported(const Fmt &in, const Fmt &coeff, const Fmt &acc, const Fmt &out, int n_taps, int ftype) {
  (void)out;
  if (in.W <= 16 || in.W + (in.S ? 0 : 1) > 24) return false;          // 16-bit samples belong to fir_q15
  if (coeff.W + (coeff.S ? 0 : 1) > 16) return false;                  // the taps ride in the signed 16-bit DP2A lanes
  if (acc.O != B2D_WRAP || (acc.Q != B2D_TRN && acc.Q != B2D_RND)) return false;
  const int lsh = acc.F() - in.F() - coeff.F();
  if (lsh < 0 || lsh > 40) return false;
  if (n_taps > kQ24MaxTaps) return false;
  switch (ftype) {
    case B2D_SHIFT_REG: case B2D_ROTATE_SHIFT: case B2D_C_BUFF: case B2D_TRANSPOSED: case B2D_FOLD_EVEN: return true;
    case B2D_FOLD_EVEN_ANTI:          // the negated mirrored taps must still fit the signed lane
      return coeff.W + (coeff.S ? 0 : 1) <= 15;
    case B2D_FOLD_ODD_ANTI:
      if (coeff.W + (coeff.S ? 0 : 1) > 15) return false;
      // fall through
    case B2D_FOLD_ODD:
      // `fold` is ACC_TYPE (ac_fir_load_coeffs.h:248-255): exact only if the pre-add neither truncates nor wraps there
      if (!acc.S && (in.S || ftype == B2D_FOLD_ODD_ANTI)) return false;
      return acc.F() >= in.F() && in.W + 1 + (in.S ? 0 : 1) + (acc.F() - in.F()) <= acc.W + (acc.S ? 0 : 1);
    default: return false;
  }
}

int fir_q24_pk_words(int n_taps) {
  const int npad = (n_taps + kQ24Chunk - 1) / kQ24Chunk * kQ24Chunk;
  return npad / 2;
}

// Effective direct-form taps (folds expanded), reversed (g[k] = h[N-1-k], zero padded), two per word.
void fir_q24_pack(const int64_t *c, int n_taps, int ftype, uint32_t *pk, int pk_words) {
  const int N = n_taps;
  std::vector<int64_t> eff(N, 0);
  const int64_t sg = (ftype == B2D_FOLD_EVEN_ANTI || ftype == B2D_FOLD_ODD_ANTI) ? -1 : 1;   // ac_fir_reg_share.h:151-165,186-205
  if (ftype == B2D_FOLD_EVEN || ftype == B2D_FOLD_EVEN_ANTI) {          // ac_fir_load_coeffs.h:231-239
    for (int i = 0; i < N / 2; i++) { eff[i] = c[i]; eff[N - 1 - i] = sg * c[i]; }
  } else if (ftype == B2D_FOLD_ODD || ftype == B2D_FOLD_ODD_ANTI) {     // :246-259
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
      const int k = 2 * w + e;
      const int64_t g = k < N ? eff[N - 1 - k] : 0;
      word |= (uint32_t)(g & 0xFFFF) << (16 * e);
    }
    pk[w] = word;
  }
}

cudaError_t launch_fir_q24(const FirLaunch &p, cudaStream_t st) {
  if (p.n == 0) return cudaSuccess;
  Q24Args a;
  a.x = p.in; a.y = p.out; a.tail = p.tail; a.cpk = p.coeff_pk; a.n = p.n;
  a.N = p.n_taps; a.pkw = p.pk_words; a.Npad = p.pk_words * 2;
  a.C = p.C; a.interleaved = p.interleaved;
  a.lsh = p.facc.F() - p.fin.F() - p.fcoeff.F();
  a.in = p.fin; a.acc = p.facc; a.out = p.fout; a.out_bytes = container_bytes(p.fout.W);
  a.fastout = (p.fout.W == p.facc.W && p.fout.I == p.facc.I && p.fout.S == p.facc.S && a.out_bytes == 8) ? 1 : 0;
  a.vec_ok = (((uintptr_t)p.in) & 15) == 0;
  const size_t per_pass = (size_t)kQ24Threads * kQ24T;
  // as in launch_fir_q15: a call too short to fill the GPU with four-pass tiles takes fewer passes per tile
  size_t passes = (p.n * p.C + per_pass * (148 * 9) - 1) / (per_pass * (148 * 9));
  if (passes > 4) passes = 4;
  if (passes < 1) passes = 1;
  a.passes = (int)passes;
  const size_t tile = per_pass * passes;
  const size_t stride = (tile + a.Npad + 16 + 15) & ~(size_t)15;
  const size_t smem = (size_t)((a.pkw + 3) & ~3) * 4 + 3 * stride;
  dim3 grid((unsigned)((p.n + tile - 1) / tile), p.C);
  const int lead = (4 - ((p.n_taps - 1) & 3)) & 3;
  cudaError_t e = cudaSuccess;
  const bool one = a.Npad <= kQ24Block;
#define B2D_Q24_LAUNCH(L)                                                                                                    \
  do {                                                                                                                       \
    if (one) {                                                                                                               \
      if (smem > 48 * 1024) e = cudaFuncSetAttribute(fir_q24_kernel<L, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); \
      if (e == cudaSuccess) fir_q24_kernel<L, true><<<grid, kQ24Threads, smem, st>>>(a);                                     \
    } else {                                                                                                                 \
      if (smem > 48 * 1024) e = cudaFuncSetAttribute(fir_q24_kernel<L, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); \
      if (e == cudaSuccess) fir_q24_kernel<L, false><<<grid, kQ24Threads, smem, st>>>(a);                                    \
    }                                                                                                                        \
  } while (0)
  if (lead == 0) B2D_Q24_LAUNCH(0);
  else if (lead == 1) B2D_Q24_LAUNCH(1);
  else if (lead == 2) B2D_Q24_LAUNCH(2);
  else B2D_Q24_LAUNCH(3);
#undef B2D_Q24_LAUNCH
  return e != cudaSuccess ? e : cudaGetLastError();
}

}  // namespace b2d
