// fir_ovs.cuh -- thread-level phases of the overlap-save FIR kernel (fir_ovs.cu), written so that the same code runs
// as CUDA threads on the GPU and as a loop over thread ids on the CPU (tests/cpp/fir_ovs_check.cu checks the indexing
// and the exactness margin against a direct integer convolution without a GPU).
//
// The tap-MAC loop  acc += reg[i]*h[i]  (reference include/ac_dsp/ac_fir_load_coeffs.h:180-278) of a long filter over a
// long stream is a linear convolution.  fir_q15 evaluates it tap by tap at one DP2A per 16x16 MAC, which pins the
// 256-tap filter at the INT pipe's issue rate (10 % of HBM).  Here a block of 4096 complex samples (an interleaved IQ
// pair, or two consecutive blocks of one real channel as real and imaginary part) is convolved through a 4096-point
// FP64 FFT held in registers and shared memory:  y = IFFT(FFT(x) . H),  H = FFT(taps)/4096 prepared at load time --
// about 80 FP64 instructions per sample whatever the tap count.  The result is an integer below 2^53 known to within
// an a-priori error bound that the launch predicate keeps under 1/2 (fir_ovs.cu: ovs_error_bound), so rounding to the
// nearest integer returns the exact dot product, and the ACC_TYPE / OUT_TYPE epilogue is the one of fir_q15.
//
// Transform: 4096 = 16 x 16 x 16, three radix-16 passes, decimation in frequency forward (natural order in, base-16
// digit-reversed order out) and the mirrored decimation in time backward, so the spectrum is never reordered: H is
// stored in the digit-reversed order the forward passes produce.  One thread owns 16 points of every pass.
#pragma once
#include "common.cuh"

#if defined(__CUDA_ARCH__)
#define OVS_LDG(p) __ldg(p)
#else
#define OVS_LDG(p) (*(p))
#endif
// The compiler may not move memory operations across this point: the shared-memory loads of a phase are issued four at a
// time next to their use.  Hoisted above the butterflies (what the scheduler does when left alone) they cost up to 60 more
// live registers and the kernel spills at its 128-register budget (512 threads per SM).
#define OVS_FENCE() asm volatile("" ::: "memory")

namespace b2d {
namespace ovs {

constexpr int kN = 4096;
constexpr int kThreads = 256;               // 16 points per thread
constexpr int kSmElems = kN + kN / 16;      // one pad element per 16: the stride-16 pass reads conflict-free
constexpr double kC1 = 0.92387953251128673848;   // cos(pi/8)
constexpr double kS1 = 0.38268343236508978178;   // sin(pi/8)
constexpr double kR = 0.70710678118654752440;    // sqrt(1/2)

__host__ __device__ __forceinline__ int pad(int i) { return i + (i >> 4); }
// register that holds element j of a digit-transposed 16-vector (dft16_* below)
__host__ __device__ constexpr int perm(int j) { return 4 * (j & 3) + (j >> 2); }

struct Args {
  const void *x;          // input samples (int16 containers)
  void *y;                // output containers
  const void *tail;       // [C][T] history, planar
  const double2 *tw;      // [6][256]: W_4096^(t b), b = 1..3, W_4096^(4 t a), a = 1..3; then [6][16]: the same of W_256 (fir_ovs_tables)
  const double2 *hs;      // [C][16][256]: spectrum of the taps / 4096 at position 16*c + j, stored [j][c]
  size_t n;               // samples per channel in this call
  int T, D, L;            // history length (n_taps - 1), discarded head of a block (multiple of 256, >= T), L = kN - D
  uint32_t C;
  unsigned per_channel;   // work items per channel (blocks of an IQ pair, or block pairs of a real channel)
  int xs;                 // samples signed
  int lsh;                // exact left shift of the dot product into ACC_TYPE
  int magic_shl, wrap_shr; // > 0: ACC_TYPE raw = (bits(v + 1.5 * 2^52) << magic_shl) >> wrap_shr (arithmetic if ACC_TYPE is signed)
  Fmt acc, out;
  int out_bytes, fastout;
  double *resid;          // optional: max |v - rint(v)| over the launch (tests), as the bits of a non-negative double
};

// Samples are read once per launch: no L1 allocation.
__host__ __device__ __forceinline__ double2 ld_stream(const double2 *p) {
#if defined(__CUDA_ARCH__)
  double2 v;
  asm volatile("ld.global.nc.L1::no_allocate.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "l"(p));
  return v;
#else
  return *p;
#endif
}
__host__ __device__ __forceinline__ uint32_t ld_stream(const uint32_t *p) {
#if defined(__CUDA_ARCH__)
  uint32_t v;
  asm volatile("ld.global.nc.L1::no_allocate.b32 %0, [%1];" : "=r"(v) : "l"(p));
  return v;
#else
  return *p;
#endif
}
__host__ __device__ __forceinline__ uint16_t ld_stream(const uint16_t *p) {
#if defined(__CUDA_ARCH__)
  uint16_t v;
  asm volatile("ld.global.nc.L1::no_allocate.b16 %0, [%1];" : "=h"(v) : "l"(p));
  return v;
#else
  return *p;
#endif
}

__host__ __device__ __forceinline__ double2 cadd(double2 a, double2 b) { return make_double2(a.x + b.x, a.y + b.y); }
__host__ __device__ __forceinline__ double2 csub(double2 a, double2 b) { return make_double2(a.x - b.x, a.y - b.y); }
// a * (wr + j*wi); INV: a * (wr - j*wi)
template <bool INV>
__host__ __device__ __forceinline__ double2 cmul(double2 a, double wr, double wi) {
  if (!INV) return make_double2(a.x * wr - a.y * wi, a.x * wi + a.y * wr);
  return make_double2(a.x * wr + a.y * wi, a.y * wr - a.x * wi);
}

// 4-point DFT in place (INV: conjugate kernel, unscaled).
template <bool INV>
__host__ __device__ __forceinline__ void bfly4(double2 &a0, double2 &a1, double2 &a2, double2 &a3) {
  const double2 s02 = cadd(a0, a2), d02 = csub(a0, a2), s13 = cadd(a1, a3), d13 = csub(a1, a3);
  a0 = cadd(s02, s13);
  a2 = csub(s02, s13);
  if (!INV) {   // a1 = d02 - j*d13, a3 = d02 + j*d13
    a1 = make_double2(d02.x + d13.y, d02.y - d13.x);
    a3 = make_double2(d02.x - d13.y, d02.y + d13.x);
  } else {
    a1 = make_double2(d02.x - d13.y, d02.y + d13.x);
    a3 = make_double2(d02.x + d13.y, d02.y - d13.x);
  }
}

// multiply v[4c + b] by W16^(c*b) (INV: conjugate), c, b = 0..3
template <bool INV>
__host__ __device__ __forceinline__ void twiddle16(double2 (&v)[16]) {
  // forward W16^m = cos(pi m/8) - j sin(pi m/8)
  v[5] = cmul<INV>(v[5], kC1, -kS1);                                        // m = 1
  v[7] = cmul<INV>(v[7], kS1, -kC1);                                        // m = 3
  v[13] = cmul<INV>(v[13], kS1, -kC1);                                      // m = 3
  v[15] = cmul<INV>(v[15], -kC1, kS1);                                      // m = 9
  {  // m = 2: R*(1 - j)
    const double2 a = v[6], b = v[9];
    if (!INV) { v[6] = make_double2(kR * (a.x + a.y), kR * (a.y - a.x)); v[9] = make_double2(kR * (b.x + b.y), kR * (b.y - b.x)); }
    else { v[6] = make_double2(kR * (a.x - a.y), kR * (a.y + a.x)); v[9] = make_double2(kR * (b.x - b.y), kR * (b.y + b.x)); }
  }
  {  // m = 6: R*(-1 - j)
    const double2 a = v[11], b = v[14];
    if (!INV) { v[11] = make_double2(kR * (a.y - a.x), -kR * (a.x + a.y)); v[14] = make_double2(kR * (b.y - b.x), -kR * (b.x + b.y)); }
    else { v[11] = make_double2(-kR * (a.x + a.y), kR * (a.x - a.y)); v[14] = make_double2(-kR * (b.x + b.y), kR * (b.x - b.y)); }
  }
  {  // m = 4: -j (INV: +j)
    const double2 a = v[10];
    v[10] = INV ? make_double2(-a.y, a.x) : make_double2(a.y, -a.x);
  }
}

// 16-point DFT, decimation in frequency: input element k in v[k], output element j in v[perm(j)].
template <bool INV>
__host__ __device__ __forceinline__ void dft16_nat2perm(double2 (&v)[16]) {
#pragma unroll
  for (int b = 0; b < 4; b++) bfly4<INV>(v[b], v[4 + b], v[8 + b], v[12 + b]);
  twiddle16<INV>(v);
#pragma unroll
  for (int c = 0; c < 4; c++) bfly4<INV>(v[4 * c], v[4 * c + 1], v[4 * c + 2], v[4 * c + 3]);
}
// the mirrored form: input element j in v[perm(j)], output element k in v[k].
template <bool INV>
__host__ __device__ __forceinline__ void dft16_perm2nat(double2 (&v)[16]) {
#pragma unroll
  for (int c = 0; c < 4; c++) bfly4<INV>(v[4 * c], v[4 * c + 1], v[4 * c + 2], v[4 * c + 3]);
  twiddle16<INV>(v);
#pragma unroll
  for (int b = 0; b < 4; b++) bfly4<INV>(v[b], v[4 + b], v[8 + b], v[12 + b]);
}

// One complex input sample of block `blk`: position `pos` of the block is global sample blk*L - D + pos.
// NP == 2: interleaved IQ pair (I, Q) of one block.  NP == 1: one real channel c0, blocks 2*blk (real part) and
// 2*blk + 1 (imaginary part).  History before the call comes from the carried tail, samples past n read as zero.
template <int NP>
__host__ __device__ __forceinline__ double2 load_point(const Args &a, uint32_t c0, long long blk, int pos) {
  if (NP == 2) {
    const long long g = blk * a.L - a.D + pos;
    int vi = 0, vq = 0;
    if (g < 0) {
      if (g + a.T >= 0) {
        const uint16_t *t16 = (const uint16_t *)a.tail;
        vi = t16[(size_t)(a.T + g)]; vq = t16[(size_t)a.T + (size_t)(a.T + g)];
      }
    } else if ((size_t)g < a.n) {
      const uint32_t w = ld_stream((const uint32_t *)a.x + g);
      vi = (int)(w & 0xFFFF); vq = (int)(w >> 16);
    }
    if (a.xs) { vi = (int)(int16_t)vi; vq = (int)(int16_t)vq; }
    return make_double2((double)vi, (double)vq);
  } else {
    const uint16_t *xc = (const uint16_t *)a.x + (size_t)c0 * a.n;
    const uint16_t *t16 = (const uint16_t *)a.tail + (size_t)c0 * a.T;
    int v[2];
#pragma unroll
    for (int e = 0; e < 2; e++) {
      const long long g = (2 * blk + e) * a.L - a.D + pos;
      int s = 0;
      if (g < 0) { if (g + a.T >= 0) s = t16[(size_t)(a.T + g)]; }
      else if ((size_t)g < a.n) s = ld_stream(xc + g);
      v[e] = a.xs ? (int)(int16_t)s : s;
    }
    return make_double2((double)v[0], (double)v[1]);
  }
}

// Twiddles of a pass, factored: output j = 4 a + b of a radix-16 butterfly is multiplied by W^(x j) = W^(x b) * W^(4 x a),
// x the thread's index in the pass.  A thread reads six table values per pass (tw[0..2] = W^(x b), b = 1..3, and
// tw[3..5] = W^(4 x a), a = 1..3, `stride` elements apart) instead of fifteen: the table of the first pass shrinks from 60 KB
// to 24 KB, which is what lets the spectrum live in shared memory next to both block buffers; nine of the fifteen outputs
// pay a second complex multiplication (FP64 has the headroom, the shared-memory pipe has not).
struct Tw6 { double2 w[6]; };
__host__ __device__ __forceinline__ Tw6 load_tw6(const double2 *tw, int stride, int x) {
  Tw6 t;
#pragma unroll
  for (int i = 0; i < 6; i++) t.w[i] = tw[i * stride + x];
  return t;
}
template <bool INV>
__host__ __device__ __forceinline__ double2 twiddle_j(double2 v, const Tw6 &t, int j) {
  const int a = j >> 2, b = j & 3;
  if (b) v = cmul<INV>(v, t.w[b - 1].x, t.w[b - 1].y);
  if (a) v = cmul<INV>(v, t.w[2 + a].x, t.w[2 + a].y);
  return v;
}

// ---- phase A: samples -> registers, pass 1 (stride 256), twiddle W_4096^(t*j), to shared memory.
// INTERIOR: every sample of the block lies inside this call's input (no history, no stream end): plain strided loads.
// Shared-memory index of position i: pad(i) = i + (i >> 4); for i = tid + 256 j that is tid + (tid >> 4) + 272 j.
template <int NP, bool INTERIOR>
__host__ __device__ __forceinline__ void phase_a(const Args &a, const double2 *tw1, uint32_t c0, long long blk, int tid, double2 *sm) {
  double2 v[16];
  if (INTERIOR && NP == 2) {
    const uint32_t *p = (const uint32_t *)a.x + (blk * a.L - a.D + tid);
#pragma unroll
    for (int k = 0; k < 16; k++) {
      const uint32_t w = ld_stream(p + 256 * k);
      const int vi = a.xs ? (int)(int16_t)(w & 0xFFFF) : (int)(w & 0xFFFF);
      const int vq = a.xs ? ((int)w >> 16) : (int)(w >> 16);
      v[k] = make_double2((double)vi, (double)vq);
    }
  } else if (INTERIOR) {
    const uint16_t *p = (const uint16_t *)a.x + (size_t)c0 * a.n + (2 * blk * a.L - a.D + tid);
#pragma unroll
    for (int k = 0; k < 16; k++) {
      const int s0 = ld_stream(p + 256 * k), s1 = ld_stream(p + a.L + 256 * k);
      v[k] = make_double2((double)(a.xs ? (int)(int16_t)s0 : s0), (double)(a.xs ? (int)(int16_t)s1 : s1));
    }
  } else {
#pragma unroll
    for (int k = 0; k < 16; k++) v[k] = load_point<NP>(a, c0, blk, tid + 256 * k);
  }
  dft16_nat2perm<false>(v);
  OVS_FENCE();
  const Tw6 t = load_tw6(tw1, 256, tid);
  double2 *s0 = sm + tid + (tid >> 4);
#pragma unroll
  for (int j = 0; j < 16; j++) s0[272 * j] = twiddle_j<false>(v[perm(j)], t, j);
}

// ---- phase B: pass 2 (stride 16 inside each block of 256), twiddle W_256^(u*j).
// position 256 b + u + 16 k -> shared-memory index 272 b + u + 17 k
__host__ __device__ __forceinline__ void phase_b(const double2 *tw2, int tid, double2 *sm) {
  const int u = tid & 15;
  double2 *s0 = sm + (tid >> 4) * 272 + u;
  double2 v[16];
#pragma unroll
  for (int k = 0; k < 16; k++) v[k] = s0[17 * k];
  dft16_nat2perm<false>(v);
  OVS_FENCE();
  const Tw6 t = load_tw6(tw2, 16, u);
#pragma unroll
  for (int j = 0; j < 16; j++) s0[17 * j] = twiddle_j<false>(v[perm(j)], t, j);
}

// ---- phase C: pass 3 (16 consecutive points), times H, first backward pass, in registers.
// hsm: this channel's spectrum in shared memory, [16][256], value for position 16 tid + j at hsm[256 j + tid].
__host__ __device__ __forceinline__ void phase_c(const double2 *hsm, int tid, double2 *sm) {
  double2 v[16];
  double2 *s0 = sm + 17 * tid;
#pragma unroll
  for (int k = 0; k < 16; k++) v[k] = s0[k];
  dft16_nat2perm<false>(v);
#pragma unroll
  for (int q = 0; q < 16; q += 4) {
    OVS_FENCE();
#pragma unroll
    for (int j = q; j < q + 4; j++) {
      const double2 h = hsm[256 * j + tid];
      v[perm(j)] = cmul<false>(v[perm(j)], h.x, h.y);
    }
  }
  OVS_FENCE();
  dft16_perm2nat<true>(v);
#pragma unroll
  for (int k = 0; k < 16; k++) s0[k] = v[k];
}

// ---- phase D: backward pass 2
__host__ __device__ __forceinline__ void phase_d(const double2 *tw2, int tid, double2 *sm) {
  const int u = tid & 15;
  double2 *s0 = sm + (tid >> 4) * 272 + u;
  const Tw6 t = load_tw6(tw2, 16, u);
  double2 v[16];
#pragma unroll
  for (int q = 0; q < 16; q += 4) {
#pragma unroll
    for (int j = q; j < q + 4; j++) v[perm(j)] = twiddle_j<true>(s0[17 * j], t, j);
    OVS_FENCE();
  }
  dft16_perm2nat<true>(v);
#pragma unroll
  for (int k = 0; k < 16; k++) s0[17 * k] = v[k];
}

// ---- phase E: backward pass 1 into registers: v[k] = block position tid + 256*k (first D positions are discarded)
__host__ __device__ __forceinline__ void phase_e(const double2 *tw1, int tid, const double2 *sm, double2 (&v)[16]) {
  const double2 *s0 = sm + tid + (tid >> 4);
  const Tw6 t = load_tw6(tw1, 256, tid);
#pragma unroll
  for (int q = 0; q < 16; q += 4) {
#pragma unroll
    for (int j = q; j < q + 4; j++) v[perm(j)] = twiddle_j<true>(s0[272 * j], t, j);
    OVS_FENCE();
  }
  dft16_perm2nat<true>(v);
}

// Is every sample the block reads, and every output it writes, inside this call's arrays?
template <int NP>
__host__ __device__ __forceinline__ bool block_interior(const Args &a, long long blk) {
  const long long first = (NP == 2 ? blk : 2 * blk) * a.L - a.D;
  const long long last = (NP == 2 ? blk : 2 * blk + 1) * a.L - a.D + kN;
  return first >= 0 && (size_t)last <= a.n;
}

// frequency index held at position p of the spectrum the forward passes leave (base-16 digit reversal)
__host__ __device__ constexpr int spectrum_index(int p) { return (p >> 8) + 16 * ((p >> 4) & 15) + 256 * (p & 15); }

}  // namespace ovs
}  // namespace b2d
