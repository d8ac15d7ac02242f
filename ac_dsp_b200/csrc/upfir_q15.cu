// upfir_q15.cu -- interpolating polyphase FIR on 16-bit samples:  out[k*R + ph] = sum_m c[ph + R*m] * x[k - m].
//
// This is the fused form of BASELINE config 5, ac_cic_intr_full (R, M, N) followed by an ac_fir_* filter:
// the interpolator is the FIR boxcar(R*M)^(*N) on the zero-stuffed input (cic_intr_fast.cu; reference
// ac_cic_intr_full.h:150-215), the second stage is a direct-form FIR on its output (ac_fir_load_coeffs.h:180-188), and
// when neither stage drops bits -- the CIC's lossless INT_TYPE is passed on unchanged and the FIR accumulator has
// s = F_mid + F_c - F_acc <= 0 with AC_WRAP -- the cascade is ONE integer FIR with the composite taps
// c = boxcar^N * g, evaluated here modulo 2^64 and wrapped to ACC_TYPE exactly like the two reference objects in
// sequence.  Only every R-th sample of the zero-stuffed input is non-zero, so an output costs ceil(len(c)/R) MACs on
// the ORIGINAL 16-bit samples (18 instead of 3 adds + 63 MACs on 20-bit samples for R=4, N=3, 63 taps), and the
// 20-bit intermediate stream (16 B per input sample each way) never touches HBM: 2 B in, R*8 B out.
//
// Arithmetic: as in fir_q15.cu, DP2A on byte planes of the taps -- the composite taps need up to 24 bits, i.e. three
// planes (unsigned low, unsigned middle, signed high byte), each exact in an int32 for <= 128 taps per phase.
// A lane owns one input period of each of JT groups of 128 periods (the lane-per-period kernel below; the round-1
// thread-owns-8-periods kernel with a staging tile lost every A/B against it and was removed in round 2).  CTAs are
// persistent over an 8-fold over-decomposed grid; the next tile's samples are fetched into registers while the current
// one computes.
#include <cstdlib>
#include <type_traits>
#include <vector>

#include "kernels.h"

namespace b2d {

constexpr int kUpThreads = 128;

struct UpArgs {
  const int16_t *x;       // inputs of this call
  void *y;                // outputs, planar, stride n_out
  const int16_t *tail;    // [C][H] previous inputs
  const uint32_t *cw;     // [C][R][TP][WPP] packed taps
  size_t n, n_out;
  long long n_seen, out_first;
  int H, Tc, TP;          // taps per phase, pairs per phase (Tc padded to even = 2*TP)
  uint32_t C;
  int interleaved;
  int lsh;
  Fmt acc, out;
  int out_bytes, fastout;
  long long ntiles;
};

__device__ __forceinline__ int up_dp2a_lo_u(uint32_t a, uint32_t b, int c) {
  int d; asm("dp2a.lo.s32.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c)); return d;
}
__device__ __forceinline__ int up_dp2a_hi_u(uint32_t a, uint32_t b, int c) {
  int d; asm("dp2a.hi.s32.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c)); return d;
}
__device__ __forceinline__ int up_dp2a_lo_s(uint32_t a, uint32_t b, int c) {
  int d; asm("dp2a.lo.s32.s32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c)); return d;
}
__device__ __forceinline__ int up_dp2a_hi_s(uint32_t a, uint32_t b, int c) {
  int d; asm("dp2a.hi.s32.s32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c)); return d;
}

// NPAIRS tap pairs starting at pair p0 for KT outputs: E words from shared memory, odd-aligned words by PRMT.
// ------------------------------------------------------------------------------------------------------------------
// Lane-per-period form.  A lane owns ONE input period k of each of JT groups of 128 consecutive periods,
// i.e. the R consecutive outputs k*R .. k*R+R-1, and keeps R * PLANES accumulators per group.  Its tap-pair operands
// are the 32-bit words (x[k-(Tcp-1)+2p], x[k-(Tcp-1)+2p+1]); for odd k those straddle word boundaries, so the tile is
// staged twice, the second copy shifted by one sample and placed 16 banks away: even lanes read word (q/2 + p) of the
// first copy, odd lanes the same index of the second -- 32 distinct banks per LDS.32.  The packed taps of all R
// phases of one pair are contiguous (one or two broadcast LDS.128 per pair serve R * PLANES * JT DP2As).  The R results
// of a period are stored straight from registers as 128-bit words: consecutive lanes write consecutive R*8-byte
// groups, so a pair of store instructions covers 1 KB contiguously -- no staging tile, no copy-out pass, and one
// barrier per tile (the sample double buffer).
template <int R, int JT, int PLANES, int NARROW, bool PEEL = false>
__global__ void __launch_bounds__(kUpThreads) upfir_lane_kernel(UpArgs a) {
  extern __shared__ __align__(16) unsigned char smem[];
  constexpr int WPP = PLANES == 3 ? 2 : 1;
  constexpr int CW = R * WPP;                      // coefficient words per tap pair, all phases
  constexpr int TILE = kUpThreads * JT;            // input periods per CTA and tile
  const int TP = a.TP, Tcp = 2 * TP;
  const int ncw = TP * CW;
  const int ncw_pad = (ncw + 3) & ~3;
  const int nxs = TILE + Tcp;                                    // samples staged per tile
  const int XS = ((((nxs + 1) / 2 + 1) + 31) & ~31) + 16;        // words per copy: the second copy sits 16 banks away
  uint32_t *cw = (uint32_t *)smem;                               // [TP][R][WPP]
  uint32_t *xsw = cw + ncw_pad;                                  // [2 buffers][2 copies][XS]
  long long *ys = (long long *)(xsw + 4 * XS);                   // [TILE][R] staging, used by edge tiles only
  const uint32_t c = blockIdx.y;
  const long long kbase = a.out_first / R;
  const int16_t *xc = a.interleaved ? a.x + c : a.x + (size_t)c * a.n;
  const size_t xstride = a.interleaved ? a.C : 1;
  const long long lo = a.out_first, hi = a.out_first + (long long)a.n_out;
  const int sh_dn = 64 - a.acc.W, sh_up = sh_dn + a.lsh;
  const bool unsigned_acc = !(a.acc.S || a.acc.W >= 64);
  const long long umask = unsigned_acc ? (long long)((1ULL << a.acc.W) - 1) : -1LL;

  for (int i = threadIdx.x; i < ncw; i += kUpThreads) {          // [R][TP][WPP] -> [TP][R][WPP]
    const int w = i % WPP, ph = (i / WPP) % R, p = i / CW;
    cw[i] = a.cw[(size_t)c * ncw + ((size_t)ph * TP + p) * WPP + w];
  }

  constexpr int NL = (TILE + 130 + kUpThreads - 1) / kUpThreads;   // Tcp <= 130
  int16_t pre[NL];
  auto fetch = [&](long long tile) {
    const long long k_tile = kbase + tile * TILE;
    const long long l0 = k_tile - (Tcp - 1) - a.n_seen;          // index of xb[0] in this call's input
#pragma unroll
    for (int q = 0; q < NL; q++) pre[q] = 0;
    if (l0 >= 0 && (size_t)(l0 + nxs) <= a.n) {                  // interior tile: no bounds, no history
      const int16_t *xp = xc + (size_t)l0 * xstride;
      if (xstride == 1) {
#pragma unroll
        for (int q = 0; q < NL; q++) { const int i = threadIdx.x + q * kUpThreads; if (i < nxs) pre[q] = xp[i]; }
      } else {
#pragma unroll
        for (int q = 0; q < NL; q++) { const int i = threadIdx.x + q * kUpThreads; if (i < nxs) pre[q] = xp[(size_t)i * xstride]; }
      }
      return;
    }
#pragma unroll
    for (int q = 0; q < NL; q++) {
      const int i = threadIdx.x + q * kUpThreads;
      const long long li = l0 + i;
      if (i < nxs) {
        if (li >= 0) { if ((size_t)li < a.n) pre[q] = xc[(size_t)li * xstride]; }
        else if (li >= -(long long)a.H) pre[q] = a.tail[(size_t)c * a.H + (size_t)(a.H + li)];
      }
    }
  };
  if ((long long)blockIdx.x < a.ntiles) fetch(blockIdx.x);

  int it = 0;
  for (long long tile = blockIdx.x; tile < a.ntiles; tile += gridDim.x, it++) {
    uint32_t *xb = xsw + (size_t)(it & 1) * 2 * XS;               // copy 0: xb16[i] = x[k_tile-(Tcp-1)+i]; copy 1: shifted by one
    {
      int16_t *xb16 = (int16_t *)xb, *xo16 = (int16_t *)(xb + XS);
#pragma unroll
      for (int q = 0; q < NL; q++) {
        const int i = threadIdx.x + q * kUpThreads;
        if (i < nxs) { xb16[i] = pre[q]; if (i > 0) xo16[i - 1] = pre[q]; }
      }
    }
    __syncthreads();           // the buffer written two tiles ago was last read before the previous barrier
    if (tile + gridDim.x < a.ntiles) fetch(tile + gridDim.x);

    const long long k_tile = kbase + tile * TILE;
    {
      int acc[JT][R][PLANES];
      if (PEEL) {                                 // A/B variant: the first tap pair (TP >= 1) initialises the accumulators, no zeroing pass
        const uint32_t *xq = xb + ((threadIdx.x & 1) ? XS : 0) + (threadIdx.x >> 1);
        // one tap pair of every phase for the JT periods of this lane; `first` (a compile-time flag) starts the
        // accumulators from zero operands instead of reading them
        auto pair = [&](const int p, auto first) {
          constexpr bool FIRST = decltype(first)::value;
          uint32_t w[CW];
          if (CW % 4 == 0) {
#pragma unroll
            for (int i = 0; i < CW / 4; i++) {
              const uint4 v = *(const uint4 *)(cw + p * CW + 4 * i);
              w[4 * i] = v.x; w[4 * i + 1] = v.y; w[4 * i + 2] = v.z; w[4 * i + 3] = v.w;
            }
          } else {                                                  // R = 2, two planes: 2 words per pair
            const uint2 v = *(const uint2 *)(cw + p * CW);
            w[0] = v.x; w[1] = v.y;
          }
#pragma unroll
          for (int j = 0; j < JT; j++) {
            const uint32_t sx = xq[j * (kUpThreads / 2) + p];
#pragma unroll
            for (int ph = 0; ph < R; ph++) {
              const uint32_t wa = w[ph * WPP];
              acc[j][ph][0] = up_dp2a_lo_u(sx, wa, FIRST ? 0 : acc[j][ph][0]);
              if (PLANES == 3) {
                acc[j][ph][1] = up_dp2a_hi_u(sx, wa, FIRST ? 0 : acc[j][ph][1]);
                acc[j][ph][2] = up_dp2a_lo_s(sx, w[ph * WPP + 1], FIRST ? 0 : acc[j][ph][2]);
              } else {
                acc[j][ph][1] = up_dp2a_hi_s(sx, wa, FIRST ? 0 : acc[j][ph][1]);
              }
            }
          }
        };
        pair(0, std::true_type());
#pragma unroll 8
        for (int p = 1; p < TP; p++) pair(p, std::false_type());
      } else {
#pragma unroll
        for (int j = 0; j < JT; j++)
#pragma unroll
          for (int ph = 0; ph < R; ph++)
#pragma unroll
            for (int pl = 0; pl < PLANES; pl++) acc[j][ph][pl] = 0;
        const uint32_t *xq = xb + ((threadIdx.x & 1) ? XS : 0) + (threadIdx.x >> 1);   // period q = j*128 + tid: word q/2 + p
#pragma unroll 8
        for (int p = 0; p < TP; p++) {
          uint32_t w[CW];
          if (CW % 4 == 0) {
#pragma unroll
            for (int i = 0; i < CW / 4; i++) {
              const uint4 v = *(const uint4 *)(cw + p * CW + 4 * i);
              w[4 * i] = v.x; w[4 * i + 1] = v.y; w[4 * i + 2] = v.z; w[4 * i + 3] = v.w;
            }
          } else {                                                  // R = 2, two planes: 2 words per pair
            const uint2 v = *(const uint2 *)(cw + p * CW);
            w[0] = v.x; w[1] = v.y;
          }
#pragma unroll
          for (int j = 0; j < JT; j++) {
            const uint32_t sx = xq[j * (kUpThreads / 2) + p];
#pragma unroll
            for (int ph = 0; ph < R; ph++) {
              const uint32_t wa = w[ph * WPP];
              acc[j][ph][0] = up_dp2a_lo_u(sx, wa, acc[j][ph][0]);
              if (PLANES == 3) {
                acc[j][ph][1] = up_dp2a_hi_u(sx, wa, acc[j][ph][1]);
                acc[j][ph][2] = up_dp2a_lo_s(sx, w[ph * WPP + 1], acc[j][ph][2]);
              } else {
                acc[j][ph][1] = up_dp2a_hi_s(sx, wa, acc[j][ph][1]);
              }
            }
          }
        }
      }
      // ---- results of period k: outputs k*R .. k*R+R-1
      const long long o_tile = k_tile * R;
      long long *yc = (long long *)a.y + (size_t)c * a.n_out;
      const bool interior = a.fastout && o_tile >= lo && o_tile + (long long)TILE * R <= hi;
      long long *yt = yc + (o_tile - lo) + threadIdx.x * R;                           // dereferenced on interior tiles only
      // a call starts wherever the previous one stopped (ac_cic_intr_full.h:196-214 leaves a period half emitted), so
      // a period's R outputs are either 16-byte aligned or off by one element: 64 + 128 ... + 64 bit stores then
      const bool odd = (((uintptr_t)yt) & 8) != 0;
      long long *yrow = ys + threadIdx.x * R;
#pragma unroll
      for (int j = 0; j < JT; j++) {
        long long res[R];
#pragma unroll
        for (int ph = 0; ph < R; ph++) {
          if (NARROW) {
            // 32 < W_acc <= 40 + lsh: only the low W_acc - lsh bits of the sum count, so the upper planes combine modulo
            // 2^32 and the 64-bit sum, shift and sign extension are done on 32-bit halves (no IMAD.WIDE on the DP2A pipe)
            // (a 32-bit-only form -- low word (acc0 + (m << 8)) << lsh, high bits from m + (acc0 >> 8) -- saves four instructions
            // per result but ptxas turns its shift-adds into IMADs on the DP2A pipe: 103.5 vs 111.3 G inputs/s; the same form
            // forced onto the ALU pipe with funnel shifts and three-operand adds: 113.6 vs 113.2 G, not worth a second code
            // path -- r02 A/B runs)
            const uint32_t m = PLANES == 3 ? (uint32_t)acc[j][ph][1] + ((uint32_t)acc[j][ph][2] << 8) : (uint32_t)acc[j][ph][1];
            uint32_t tl, th;
            asm("add.cc.u32 %0, %2, %3;\n\taddc.u32 %1, %4, %5;" : "=r"(tl), "=r"(th)
                : "r"((uint32_t)acc[j][ph][0]), "r"(m << 8), "r"((uint32_t)(acc[j][ph][0] >> 31)), "r"(m >> 24));
            const uint32_t rl = tl << a.lsh, rf = __funnelshift_l(tl, th, a.lsh);
            const uint32_t rh = NARROW == 2 ? (rf << sh_dn) >> sh_dn : (uint32_t)((int)(rf << sh_dn) >> sh_dn);   // keep W_acc - 32 bits
            res[ph] = (long long)(((unsigned long long)rh << 32) | rl);
          } else {
            unsigned long long tot;
            if (PLANES == 3) tot = (unsigned long long)((long long)acc[j][ph][0] + ((long long)acc[j][ph][1] << 8) + ((long long)acc[j][ph][2] << 16));
            else tot = (unsigned long long)((long long)acc[j][ph][0] + ((long long)acc[j][ph][1] << 8));
            res[ph] = (long long)(tot << sh_up) >> sh_dn;                            // wrap_W(tot << lsh), sign-extended
            if (unsigned_acc) res[ph] &= umask;
          }
        }
        if (interior) {
          long long *yp = yt + j * (kUpThreads * R);
          if (!odd) {
#pragma unroll
            for (int ph = 0; ph < R; ph += 2) *(longlong2 *)(yp + ph) = make_longlong2(res[ph], res[ph + 1]);
          } else {
            yp[0] = res[0];
#pragma unroll
            for (int ph = 1; ph + 1 < R; ph += 2) *(longlong2 *)(yp + ph) = make_longlong2(res[ph], res[ph + 1]);
            yp[R - 1] = res[R - 1];
          }
        } else {
#pragma unroll
          for (int ph = 0; ph < R; ph += 2) *(longlong2 *)(yrow + j * (kUpThreads * R) + ph) = make_longlong2(res[ph], res[ph + 1]);
        }
      }
      if (!interior) {                            // first / last tile of the call, or OUT_TYPE != ACC_TYPE: checked copy-out
        __syncthreads();
        for (int e = threadIdx.x; e < TILE * R; e += kUpThreads) {
          const long long og = o_tile + e;
          if (og >= lo && og < hi) {
            if (a.fastout) yc[og - lo] = ys[e];
            else store_raw(a.y, (size_t)c * a.n_out + (size_t)(og - lo), a.out_bytes, convert((i128)ys[e], a.acc.F(), a.out));
          }
        }
        __syncthreads();
      }
    }
  }
}

// Grid size of the two persistent kernels, in units of "what is resident at once" (148 SMs x CTAs per SM).  1 = one
// CTA per slot striding over the tiles.  ncu shows 16.6 of 24 resident warps per SM on average for that shape
// (profiles/r01_source_level_notes.md): the warp scheduler favours the older CTAs of an SM, they finish their equal
// share of tiles early and the SM runs the tail under-occupied.  B2D_UPFIR_WAVES = W (1..64) launches W x as many CTAs,
// each with 1/W of the tiles, so the hardware CTA scheduler refills a slot as soon as it frees (A/B switch; the tile
// loop is written for any grid size, results do not depend on it).
static int up_grid_waves() {
  const char *w = getenv("B2D_UPFIR_WAVES");
  const int v = w ? atoi(w) : 8;          // r02 A/B on a B200 (profiles/r02_upfir_ab_and_engine_fuzz.txt): 1 -> 8 is +8.6 % on cicfir, +7.6 % on polyintr
  return v < 1 ? 1 : (v > 64 ? 64 : v);
}

template <int R, int JT, int PLANES, int NARROW, bool PEEL = false>
static cudaError_t launch_up_lane_n(UpArgs a, cudaStream_t st) {
  constexpr int TILE = kUpThreads * JT;
  const long long kbase = a.out_first / R;
  const long long klast = (a.out_first + (long long)a.n_out - 1) / R;
  const long long nper = klast - kbase + 1;
  const int ncw = R * a.TP * (PLANES == 3 ? 2 : 1);
  const int nxs = TILE + 2 * a.TP;
  const int XS = ((((nxs + 1) / 2 + 1) + 31) & ~31) + 16;
  const size_t smem = (size_t)((ncw + 3) & ~3) * 4 + (size_t)4 * XS * 4 + (size_t)TILE * R * 8;
  a.ntiles = (nper + TILE - 1) / TILE;
  cudaError_t e = cudaFuncSetAttribute(upfir_lane_kernel<R, JT, PLANES, NARROW, PEEL>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  int per_sm = 4;
  cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, upfir_lane_kernel<R, JT, PLANES, NARROW, PEEL>, kUpThreads, smem);
  if (per_sm < 1) per_sm = 1;
  long long gx = a.ntiles;
  const long long cap = (148LL * per_sm * up_grid_waves() + a.C - 1) / a.C;
  if (gx > cap) gx = cap;
  dim3 grid((unsigned)gx, a.C);
  upfir_lane_kernel<R, JT, PLANES, NARROW, PEEL><<<grid, kUpThreads, smem, st>>>(a);
  return cudaGetLastError();
}

template <int R, int JT, int PLANES>
static cudaError_t launch_up_lane(const UpArgs &a, cudaStream_t st) {
  // W_acc - lsh <= 40: bits of the upper planes above 2^32 fall off the accumulator (see the epilogue)
  if (a.acc.W - a.lsh <= 40 && a.acc.W > 32 && a.lsh < 32) {
    // A/B switch (profiles/r01_source_level_notes.md): accumulators started by the first tap pair; instantiated for
    // the two bench geometries (R = 4, JT = 4, signed <40,8>-style accumulator) only
    // three planes: +3 % with the peeled start (72 registers instead of 80); two planes: -6 % (70 instead of 54 registers).
    // Tap loop unrolled 8-fold: +1.8 % / +1.5 % over 2-fold on cicfir / polyintr (r02 A/B, gpurun_out/r02_m_*.json)
    const char *peel = getenv("B2D_UPFIR_PEEL");
    const bool use_peel = peel ? *peel == '1' : PLANES == 3;
    if (R == 4 && JT == 4 && a.acc.S && use_peel) return launch_up_lane_n<4, 4, PLANES, 1, true>(a, st);
    return a.acc.S ? launch_up_lane_n<R, JT, PLANES, 1>(a, st) : launch_up_lane_n<R, JT, PLANES, 2>(a, st);
  }
  return launch_up_lane_n<R, JT, PLANES, 0>(a, st);
}

// ------------------------------------------------------------------------------------------ host side
bool upfir_q15_geometry(int R, int taps_total, int max_abs_bits) {
  if (R != 2 && R != 4 && R != 8) return false;
  const int Tc = (taps_total + R - 1) / R;
  return Tc >= 1 && Tc <= 128 && max_abs_bits <= 24;
}

int upfir_q15_planes(int max_abs_bits) { return max_abs_bits <= 16 ? 2 : 3; }

int upfir_q15_words(int R, int taps_total, int planes) {
  const int Tc = (taps_total + R - 1) / R;
  const int TP = (Tc + 1) / 2;
  return R * TP * (planes == 3 ? 2 : 1);
}

// c[0 .. taps_total): composite taps (signed, |c| < 2^23 for 3 planes, < 2^15 for 2).  Per phase the taps are reversed
// (sample and tap index advance together); an odd tap count is padded with one zero tap on the oldest sample.
void upfir_q15_pack(const int64_t *c, int taps_total, int R, int planes, uint32_t *out) {
  const int Tc = (taps_total + R - 1) / R;
  const int TP = (Tc + 1) / 2;
  const int wpp = planes == 3 ? 2 : 1;
  for (int ph = 0; ph < R; ph++)
    for (int p = 0; p < TP; p++) {
      int64_t g[2];
      for (int e = 0; e < 2; e++) {
        const int k = 2 * p + e;                       // reversed index: g[k] = c_ph[2*TP-1-k]
        const int m = 2 * TP - 1 - k;
        const int idx = ph + R * m;
        g[e] = (m >= 0 && m < Tc && idx < taps_total) ? c[idx] : 0;
      }
      uint32_t wa = 0, wb = 0;
      for (int e = 0; e < 2; e++) {
        const uint32_t b0 = (uint32_t)(g[e] & 0xFF), b1 = (uint32_t)((g[e] >> 8) & 0xFF), b2 = (uint32_t)((g[e] >> 16) & 0xFF);
        wa |= b0 << (8 * e);
        wa |= b1 << (16 + 8 * e);                      // planes == 2: signed high byte; planes == 3: unsigned middle byte
        wb |= b2 << (8 * e);                           // signed high byte (arithmetic shift above)
      }
      out[(ph * TP + p) * wpp] = wa;
      if (planes == 3) out[(ph * TP + p) * wpp + 1] = wb;
    }
}

cudaError_t launch_upfir_q15(const UpLaunch &p, cudaStream_t st) {
  if (p.n_out == 0) return cudaSuccess;
  UpArgs a;
  a.x = (const int16_t *)p.in; a.y = p.out; a.tail = (const int16_t *)p.tail; a.cw = p.cw; a.n = p.n; a.n_out = p.n_out;
  a.n_seen = (long long)p.n_seen; a.out_first = (long long)p.out_first; a.H = p.H;
  a.Tc = (p.taps_total + p.R - 1) / p.R; a.TP = (a.Tc + 1) / 2;
  a.C = p.C; a.interleaved = p.interleaved && p.C > 1; a.lsh = p.lsh; a.acc = p.facc; a.out = p.fout;
  a.out_bytes = container_bytes(p.fout.W);
  a.fastout = (p.fout.W == p.facc.W && p.fout.I == p.facc.I && p.fout.S == p.facc.S && a.out_bytes == 8) ? 1 : 0;
  if (p.planes == 3) {
    if (p.R == 2) return launch_up_lane<2, 4, 3>(a, st);
    if (p.R == 4) return launch_up_lane<4, 4, 3>(a, st);
    if (p.R == 8) return launch_up_lane<8, 2, 3>(a, st);
  } else {
    if (p.R == 2) return launch_up_lane<2, 8, 2>(a, st);
    if (p.R == 4) return launch_up_lane<4, 4, 2>(a, st);
    if (p.R == 8) return launch_up_lane<8, 2, 2>(a, st);
  }
  return cudaErrorNotSupported;
}

}  // namespace b2d
