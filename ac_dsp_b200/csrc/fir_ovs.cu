// fir_ovs.cu -- overlap-save evaluation of long FIR filters on 16-bit samples (BASELINE configs 1 and 3: 256 and
// 1024 taps, ac_fixed<16,1> x ac_fixed<16,1> -> <40,8>).  See fir_ovs.cuh for the transform and the phases.
//
// Taken instead of fir_q15 when the call is long enough to repay a round of blocks (rt_fir.cu: fir_ovs_worth), the filter has
// at least kMinTaps taps and the a-priori error bound of the FP64 evaluation is below 1/2 for the COEFFICIENTS ACTUALLY
// LOADED (rt_fir.cu: fir_ovs_armed, evaluated at b2d_fir_load with fir_ovs_error_bound below), so that the
// nearest integer is the exact dot product  sum_i x[n-i]*h[i]  the reference's loop accumulates
// (ac_fir_load_coeffs.h:180-278; same exactness argument as fir_q15: s <= 0, ACC_TYPE wraps).
//
// Error bound.  With u = 2^-53, block length N = 4096, |x| <= xmax per component,
// every radix-16 pass is 4 levels of additions (u each, normwise), one internal and one external twiddle
// multiplication ((sqrt(5) + 1) u each: Brent-Percival-Zimmermann bound plus the rounding of the tabulated twiddle),
// 10.5 u in all; nine of the fifteen external twiddles are applied as two factors (fir_ovs.cuh: Tw6), one more
// (sqrt(5) + 1) u: 13.8 u for the four passes with external twiddles, 7.3 u for the two without: 69.8 u for both
// transforms.  H is computed in extended precision on the host and rounded once (|dH_k| <= 1.1 u ||h||_1), the pointwise
// product adds sqrt(5) u.  In the 2-norm the exact transform passes are unitary up to scale, so the relative errors add:
//     ||y_computed - y||_inf <= ||.||_2 <= 73.2 u * ||x||_2 * ||h||_1 <= 73.2 u * sqrt(2 N) xmax ||h||_1 .
// kErrK = 76 is used.  For 256 full-scale Q15 taps (||h||_1 <= 2^23) the bound is 0.21; measured residuals
// |v - rint(v)| stay below 1e-4 (tests/test_fir_ovs.py, tests/cpp/fir_ovs_check.cu print them).
#include <algorithm>
#include <cfloat>
#include <cmath>
#include <cstdlib>
#include <vector>

#include "fir_ovs.cuh"
#include "kernels.h"

namespace b2d {

using namespace ovs;

constexpr double kErrK = 76.0;
constexpr int kMinTaps = 96;          // below this the DP2A kernel is faster
constexpr int kMaxTapsOvs = 2049;     // D <= N / 2

__host__ __device__ __forceinline__ int64_t ovs_to_acc(double d, const Args &a) {
#if defined(__CUDA_ARCH__)
  const long long s = __double2ll_rn(d);
#else
  const long long s = llrint(d);
#endif
  return wrap_bits((int64_t)((uint64_t)s << a.lsh), a.acc.W, a.acc.S);
}

constexpr int kTwElems = 6 * 256 + 6 * 16;
constexpr int kCtaThreads = 2 * kThreads;
constexpr size_t kSmemBytes = (size_t)(kTwElems + kN + 2 * kSmElems) * sizeof(double2);   // 230,912 bytes: one CTA per SM

__device__ __forceinline__ void half_sync(int half) { asm volatile("bar.sync %0, %1;" ::"r"(half + 1), "n"(kThreads) : "memory"); }

// The same value without the conversion instruction: v + 1.5 * 2^52 holds round(v) in the low bits of its mantissa (two's
// complement, |v| < 2^51), and the accumulator keeps fewer than 52 of them (a.magic_shl >= 13 pushes exponent and bit 51 out).
__device__ __forceinline__ int64_t ovs_to_acc_magic(double d, const Args &a) {
  const long long b = __double_as_longlong(d + 6755399441055744.0) << a.magic_shl;
  return a.acc.S ? (b >> a.wrap_shr) : (long long)((unsigned long long)b >> a.wrap_shr);
}

// Phase E and the interior epilogue in one: the last butterfly stage is done one group of four at a time, each group converted
// and stored while the next one is computed.
template <int NP>
__device__ __forceinline__ void phase_e_store(const Args &a, const double2 *tw1, uint32_t c0, long long blk, int tid, const double2 *sm, int k0) {
  double2 v[16];
  const double2 *s0 = sm + tid + (tid >> 4);
  const Tw6 t = load_tw6(tw1, 256, tid);
#pragma unroll
  for (int q = 0; q < 16; q += 4) {
#pragma unroll
    for (int j = q; j < q + 4; j++) v[perm(j)] = twiddle_j<true>(s0[272 * j], t, j);
    OVS_FENCE();
  }
#pragma unroll
  for (int c = 0; c < 4; c++) bfly4<true>(v[4 * c], v[4 * c + 1], v[4 * c + 2], v[4 * c + 3]);
  twiddle16<true>(v);
  long long *yp = NP == 2 ? (long long *)a.y + 2 * (blk * a.L - a.D + tid) : (long long *)a.y + (size_t)c0 * a.n + (2 * blk * a.L - a.D + tid);
#pragma unroll
  for (int b = 0; b < 4; b++) {
    bfly4<true>(v[b], v[4 + b], v[8 + b], v[12 + b]);
#pragma unroll
    for (int q = 0; q < 4; q++) {
      const int k = 4 * q + b;
      if (k == 0 || k < k0) continue;
      if (NP == 2) {
        longlong2 o; o.x = ovs_to_acc_magic(v[k].x, a); o.y = ovs_to_acc_magic(v[k].y, a);
        *(longlong2 *)(yp + 512 * k) = o;
      } else {
        yp[256 * k] = ovs_to_acc_magic(v[k].x, a);
        yp[256 * k + a.L] = ovs_to_acc_magic(v[k].y, a);
      }
    }
    OVS_FENCE();
  }
}

// Persistent CTA of two independent halves (256 threads each, own block buffer, own named barrier) that share one copy of
// the twiddle tables and of ONE channel's spectrum in shared memory -- the transform reads nothing but its samples from
// global memory.  blockIdx.y is the channel (0 for an IQ pair); half h of CTA b takes the channel's work items 2 b + h,
// 2 b + h + 2 gridDim.x, ...  A work item is one block of an IQ pair (NP == 2) or a pair of consecutive blocks of one
// real channel (NP == 1).
// Interior blocks whose accumulator takes the bit-pattern conversion run phase E fused with the epilogue (phase_e_store).
// Measured and dropped (profiles/r02_ovs_variants.txt): the next block's samples carried in registers across the epilogue
// (+1 % before the warp-level barriers, -3 % after), no L2 prefetch (-6 % on real channels).
template <int NP, bool FASTOUT>
__global__ void __launch_bounds__(kCtaThreads, 1) fir_ovs_kernel(Args a) {
  extern __shared__ __align__(16) double2 smem[];
  double2 *tw1 = smem, *tw2 = smem + 6 * 256, *hsm = smem + kTwElems;
  const uint32_t c0 = blockIdx.y;
  for (int i = threadIdx.x; i < kTwElems; i += kCtaThreads) smem[i] = a.tw[i];
  for (int i = threadIdx.x; i < kN; i += kCtaThreads) hsm[i] = a.hs[(size_t)c0 * kN + i];
  __syncthreads();
  const int half = threadIdx.x >> 8, tid = threadIdx.x & (kThreads - 1);
  double2 *sm = smem + kTwElems + kN + half * kSmElems;
  const int k0 = a.D >> 8;
  double rmax = 0.0;
  for (unsigned item = 2 * blockIdx.x + half; item < a.per_channel; item += 2 * gridDim.x) {
    const long long blk = item;
    const bool interior = block_interior<NP>(a, blk);
    {  // the samples of this half's next item on their way into L2 while this one is transformed (one 128-byte line per thread)
      const unsigned nx = item + 2 * gridDim.x;
      if (nx < a.per_channel) {
        if (NP == 2) {
          const long long g = (long long)nx * a.L - a.D + 32 * tid;
          if (tid < 128 && g >= 0 && (size_t)g < a.n) asm volatile("prefetch.global.L2 [%0];" ::"l"((const uint32_t *)a.x + g));
        } else {
          const long long g = 2 * (long long)nx * a.L - a.D + 64 * tid;
          if (tid < 128 && g >= 0 && (size_t)g < a.n) asm volatile("prefetch.global.L2 [%0];" ::"l"((const uint16_t *)a.x + (size_t)c0 * a.n + g));
        }
      }
    }
    if (interior) phase_a<NP, true>(a, tw1, c0, blk, tid, sm);
    else phase_a<NP, false>(a, tw1, c0, blk, tid, sm);
    half_sync(half);
    // passes 2, 3 and their mirrors exchange data inside groups of 16 threads only (positions 256 b + ..., b = tid >> 4: a thread
    // of phase C reads what the threads of its own group wrote in phase B, and so on): a warp-level barrier is enough
    phase_b(tw2, tid, sm);
    __syncwarp();
    phase_c(hsm, tid, sm);
    __syncwarp();
    phase_d(tw2, tid, sm);
    half_sync(half);
    if (FASTOUT && interior && a.magic_shl && !a.resid) {
      phase_e_store<NP>(a, tw1, c0, blk, tid, sm, k0);
      continue;
    }
    double2 v[16];
    phase_e(tw1, tid, sm, v);
    // no barrier here: phase A of the next item writes exactly the shared-memory elements this thread has just read
    // (positions tid + 256 j in both), and nobody else touches them before the barrier that follows phase A

    if (a.resid) {
#pragma unroll
      for (int k = 0; k < 16; k++)
        if (k >= k0) { rmax = fmax(rmax, fabs(v[k].x - rint(v[k].x))); rmax = fmax(rmax, fabs(v[k].y - rint(v[k].y))); }
    }
    if (FASTOUT && interior) {      // every kept output lands inside the call: one base pointer, immediate offsets
      if (NP == 2) {
        long long *yp = (long long *)a.y + 2 * (blk * a.L - a.D + tid);
        if (a.magic_shl) {
#pragma unroll
          for (int k = 1; k < 16; k++) {
            if (k % 5 == 1) OVS_FENCE();
            if (k < k0) continue;
            longlong2 o; o.x = ovs_to_acc_magic(v[k].x, a); o.y = ovs_to_acc_magic(v[k].y, a);
            *(longlong2 *)(yp + 512 * k) = o;
          }
        } else {
#pragma unroll
          for (int k = 1; k < 16; k++) {
            if (k % 5 == 1) OVS_FENCE();
            if (k < k0) continue;
            longlong2 o; o.x = ovs_to_acc(v[k].x, a); o.y = ovs_to_acc(v[k].y, a);
            *(longlong2 *)(yp + 512 * k) = o;
          }
        }
      } else {
        long long *yp = (long long *)a.y + (size_t)c0 * a.n + (2 * blk * a.L - a.D + tid);
        if (a.magic_shl) {
#pragma unroll
          for (int k = 1; k < 16; k++) {
            if (k % 5 == 1) OVS_FENCE();
            if (k < k0) continue;
            yp[256 * k] = ovs_to_acc_magic(v[k].x, a);
            yp[256 * k + a.L] = ovs_to_acc_magic(v[k].y, a);
          }
        } else {
#pragma unroll
          for (int k = 1; k < 16; k++) {
            if (k % 5 == 1) OVS_FENCE();
            if (k < k0) continue;
            yp[256 * k] = ovs_to_acc(v[k].x, a);
            yp[256 * k + a.L] = ovs_to_acc(v[k].y, a);
          }
        }
      }
    } else if (NP == 2) {
      const long long gb = blk * a.L - a.D + tid;
#pragma unroll
      for (int k = 1; k < 16; k++) {
        OVS_FENCE();
        const long long g = gb + 256 * k;
        if (k < k0 || (size_t)g >= a.n) continue;
        const int64_t ri = ovs_to_acc(v[k].x, a), rq = ovs_to_acc(v[k].y, a);
        if (FASTOUT) {
          longlong2 o; o.x = ri; o.y = rq;
          *(longlong2 *)((long long *)a.y + 2 * (size_t)g) = o;
        } else {
          store_raw(a.y, 2 * (size_t)g, a.out_bytes, convert((i128)ri, a.acc.F(), a.out));
          store_raw(a.y, 2 * (size_t)g + 1, a.out_bytes, convert((i128)rq, a.acc.F(), a.out));
        }
      }
    } else {
#pragma unroll
      for (int k = 1; k < 16; k++) {
        OVS_FENCE();
#pragma unroll
        for (int e = 0; e < 2; e++) {
          const long long g = (2 * blk + e) * a.L - a.D + tid + 256 * k;
          if (k < k0 || (size_t)g >= a.n) continue;
          const int64_t r = ovs_to_acc(e ? v[k].y : v[k].x, a);
          if (FASTOUT) ((long long *)a.y)[(size_t)c0 * a.n + (size_t)g] = r;
          else store_raw(a.y, (size_t)c0 * a.n + (size_t)g, a.out_bytes, convert((i128)r, a.acc.F(), a.out));
        }
      }
    }
  }
  if (a.resid) atomicMax((unsigned long long *)a.resid, (unsigned long long)__double_as_longlong(rmax));
}

// ------------------------------------------------------------------------------------------ host side
// Effective direct-form taps of every architecture the exact-sum paths take (the expansion fir_q15_pack uses).
void fir_effective_taps(const int64_t *c, int N, int ftype, int64_t *eff) {
  for (int i = 0; i < N; i++) eff[i] = 0;
  const int64_t sg = (ftype == B2D_FOLD_EVEN_ANTI || ftype == B2D_FOLD_ODD_ANTI) ? -1 : 1;
  if (ftype == B2D_FOLD_EVEN || ftype == B2D_FOLD_EVEN_ANTI) {
    for (int i = 0; i < N / 2; i++) { eff[i] = c[i]; eff[N - 1 - i] = sg * c[i]; }
  } else if (ftype == B2D_FOLD_ODD || ftype == B2D_FOLD_ODD_ANTI) {
    for (int i = 0; i < (N - 1) / 2 + 1; i++) {
      eff[i] = c[i];
      if (i != (N - 1) / 2) eff[N - 1 - i] = sg * c[i];
    }
  } else {
    for (int i = 0; i < N; i++) eff[i] = c[i];
  }
}

int fir_ovs_discard(int n_taps) { return n_taps <= 1 ? 256 : ((n_taps - 1 + 255) / 256) * 256; }

bool fir_ovs_geometry(int n_taps, uint32_t C, int interleaved) {
  if (n_taps < kMinTaps || n_taps > kMaxTapsOvs) return false;
  if (interleaved && C == 2) return true;              // one IQ pair = one complex sequence
  return (!interleaved || C == 1) && C <= 65535;       // planar real channels: one grid row per channel
}

// Upper bound of |computed - exact| for samples of format `in` and taps of 1-norm l1 (see the header of this file).
double fir_ovs_error_bound(const Fmt &in, double l1) {
  const double xmax = in.S ? std::ldexp(1.0, in.W - 1) : std::ldexp(1.0, in.W) - 1.0;
  return kErrK * std::ldexp(1.0, -53) * std::sqrt(2.0 * kN) * xmax * l1;
}

// Twiddle tables, rounded from extended precision (fir_ovs.cuh: Tw6): tw1[i][t], t < 256, i = 0..2: W_4096^(t (i + 1)),
// i = 3..5: W_4096^(4 t (i - 2)); tw2[i][u], u < 16: the same of W_256.
void fir_ovs_tables(double2 *tw1, double2 *tw2) {
  const long double tau = 6.283185307179586476925286766559005768L;
  for (int i = 0; i < 6; i++) {
    const int mul = i < 3 ? i + 1 : 4 * (i - 2);
    for (int t = 0; t < 256; t++) {
      const long double ang = tau * (long double)((t * mul) % kN) / (long double)kN;
      tw1[i * 256 + t] = make_double2((double)cosl(ang), (double)-sinl(ang));
    }
    for (int u = 0; u < 16; u++) {
      const long double ang = tau * (long double)((u * mul) % 256) / 256.0L;
      tw2[i * 16 + u] = make_double2((double)cosl(ang), (double)-sinl(ang));
    }
  }
}

// Spectrum of one channel's effective taps over 4096 points, divided by 4096, in the order phase C reads it:
// hs[j * 256 + c] = H[spectrum_index(16 c + j)] / 4096.  Direct sums in extended precision, blocked (16 taps per partial
// sum) so that the accumulated rounding stays below 2^-56 ||h||_1; real taps: H[N - f] = conj(H[f]).
void fir_ovs_spectrum(const int64_t *eff, int n_taps, double2 *hs) {
  static_assert(LDBL_MANT_DIG >= 64, "the spectrum is summed in extended precision (x87 80-bit or IEEE binary128)");
  const long double tau = 6.283185307179586476925286766559005768L;
  std::vector<long double> wr(kN), wi(kN);
  for (int m = 0; m < kN; m++) {
    const long double ang = tau * (long double)m / (long double)kN;
    wr[m] = cosl(ang); wi[m] = -sinl(ang);
  }
  std::vector<long double> hr(kN), hi(kN);
  for (int f = 0; f <= kN / 2; f++) {
    long double sr = 0, si = 0;
    for (int n0 = 0; n0 < n_taps; n0 += 16) {
      long double pr = 0, pi = 0;
      const int n1 = n0 + 16 < n_taps ? n0 + 16 : n_taps;
      for (int n = n0; n < n1; n++) {
        const int m = (int)(((long long)f * n) & (kN - 1));
        pr += (long double)eff[n] * wr[m];
        pi += (long double)eff[n] * wi[m];
      }
      sr += pr; si += pi;
    }
    hr[f] = sr; hi[f] = si;
    if (f && f < kN / 2) { hr[kN - f] = sr; hi[kN - f] = -si; }
  }
  for (int c = 0; c < 256; c++)
    for (int j = 0; j < 16; j++) {
      const int f = spectrum_index(16 * c + j);
      hs[j * 256 + c] = make_double2((double)(hr[f] / (long double)kN), (double)(hi[f] / (long double)kN));
    }
}

template <int NP, bool FASTOUT>
static cudaError_t launch_k(const Args &a, dim3 grid, cudaStream_t st) {
  const cudaError_t e = cudaFuncSetAttribute(fir_ovs_kernel<NP, FASTOUT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemBytes);
  if (e != cudaSuccess) return e;
  fir_ovs_kernel<NP, FASTOUT><<<grid, kCtaThreads, kSmemBytes, st>>>(a);
  return cudaGetLastError();
}

cudaError_t launch_fir_ovs(const FirLaunch &p, const double2 *tw, const double2 *hs, double *resid, cudaStream_t st) {
  if (p.n == 0) return cudaSuccess;
  Args a;
  a.x = p.in; a.y = p.out; a.tail = p.tail; a.tw = tw; a.hs = hs; a.n = p.n;
  a.T = p.n_taps - 1; a.D = fir_ovs_discard(p.n_taps); a.L = kN - a.D;
  a.C = p.C; a.xs = p.fin.S ? 1 : 0;
  a.lsh = p.facc.F() - p.fin.F() - p.fcoeff.F();
  a.acc = p.facc; a.out = p.fout; a.out_bytes = container_bytes(p.fout.W);
  a.fastout = (p.fout.W == p.facc.W && p.fout.I == p.facc.I && p.fout.S == p.facc.S && a.out_bytes == 8 && (((uintptr_t)p.out) & 15) == 0) ? 1 : 0;
  a.resid = resid;
  // the bit-pattern conversion needs every kept bit of the sum inside the mantissa window: W_acc - lsh <= 51
  a.magic_shl = 0; a.wrap_shr = 0;
  if (p.facc.W <= 64 && p.facc.W - a.lsh <= 51 && p.facc.W - a.lsh >= 1) { a.magic_shl = a.lsh + 64 - p.facc.W; a.wrap_shr = 64 - p.facc.W; }
  int dev = 0, sms = 148;
  if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const size_t blocks = (p.n + a.L - 1) / a.L;
  const bool iq = p.interleaved && p.C == 2;
  const size_t per_channel = iq ? blocks : (blocks + 1) / 2;
  const uint32_t chans = iq ? 1 : p.C;
  if (per_channel > 0xFFFFFFF0u || chans > 65535) return cudaErrorInvalidValue;
  a.per_channel = (unsigned)per_channel;
  // a CTA is bound to one channel (its spectrum sits in shared memory): the SMs are divided among the channels
  const unsigned per_ch_ctas = (unsigned)std::min<size_t>((per_channel + 1) / 2, std::max<unsigned>(1u, (unsigned)sms / chans));
  const dim3 grid(per_ch_ctas, chans);
  if (iq) return a.fastout ? launch_k<2, true>(a, grid, st) : launch_k<2, false>(a, grid, st);
  return a.fastout ? launch_k<1, true>(a, grid, st) : launch_k<1, false>(a, grid, st);
}

}  // namespace b2d
