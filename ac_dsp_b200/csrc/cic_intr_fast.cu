// cic_intr_fast.cu -- HBM-bound CIC interpolator for 16-bit samples and <= 32-bit lossless state
// (BASELINE config 5, first stage: ac_cic_intr_full R=4 N=3 ac_fixed<16,1> -> <20,5>).
//
// What the reference computes (ac_cic_intr_full.h:150-215, ac_cic_full_core.h:143-160,198-255): comb^N at the input
// rate, zero-stuffing by R, N integrators at the output rate, first N-1 outputs dropped; all modulo 2^intW.  The
// cascade as a whole is the FIR filter h = boxcar(R*M)^(*N) applied to the zero-stuffed input, i.e. the polyphase form
//     out[k*R + ph] = sum_m h[ph + R*m] * x[k - m]          (0 <= ph < R, 0 <= m < T, T = ceil(len(h) / R) <= N*M)
// which needs no recursion, no run-in and no carried integrator state: an output depends on T consecutive inputs.
// h is a compile-time table, so every multiply is an IMAD with an immediate; arithmetic is modulo 2^32, which
// contains the reference's modulo-2^intW arithmetic because intW <= 32.
//
// Data movement.  Per input sample: 2 B read, R*4 B written -- the stores are the traffic.  A tile is TO = 4096
// consecutive outputs of one channel, aligned to the 16-byte grid of THIS call's output array; the call's first
// output is generally not the first phase of an input period (the reference's stream-edge rule leaves a period
// half-emitted between run() calls), so the periods are computed into a shared-memory staging row in their natural
// alignment and copied out with a word offset, fully coalesced 128-bit stores either way.
#include <cstdlib>

#include "kernels.h"

namespace b2d {

constexpr int kIntrThreads = 256;
constexpr int kIntrTileOut = 4096;

template <int R, int N, int M>
struct IntrTaps {
  static constexpr int LEN = N * (R * M - 1) + 1;
  static constexpr int T = (LEN + R - 1) / R;
  unsigned v[T * R];
};

template <int R, int N, int M>
__host__ __device__ constexpr IntrTaps<R, N, M> make_intr_taps() {
  IntrTaps<R, N, M> t{};
  for (int i = 0; i < IntrTaps<R, N, M>::T * R; i++) t.v[i] = 0;
  unsigned cur[IntrTaps<R, N, M>::T * R] = {};
  cur[0] = 1;
  int len = 1;
  for (int s = 0; s < N; s++) {            // convolve with boxcar(R*M), N times
    unsigned nxt[IntrTaps<R, N, M>::T * R] = {};
    for (int i = 0; i < len; i++)
      for (int j = 0; j < R * M; j++) nxt[i + j] += cur[i];
    len += R * M - 1;
    for (int i = 0; i < len; i++) cur[i] = nxt[i];
  }
  for (int i = 0; i < len; i++) t.v[i] = cur[i];
  return t;
}

struct CicIntrArgs {
  const int16_t *x;         // inputs of this call
  void *y;                  // outputs, planar, stride n_out
  const int16_t *tail;      // [C][H] previous inputs
  size_t n, n_out;
  long long n_seen, out_first;
  int H;
  uint32_t C;
  int interleaved, intW, ident, out_bytes;
  Fmt in, out;
  long long ntiles;
};

__device__ __noinline__ void cic_intr_store_converted(const CicIntrArgs &a, uint32_t c, size_t j, uint32_t raw) {
  const int64_t w = wrap_bits((int64_t)raw, a.intW, 1);
  store_raw(a.y, (size_t)c * a.n_out + j, a.out_bytes, a.ident ? w : convert((i128)w, a.in.F(), a.out));
}

template <int R, int N, int M>
__global__ void __launch_bounds__(kIntrThreads) cic_intr_fast_kernel(CicIntrArgs a) {
  typedef IntrTaps<R, N, M> Taps;
  constexpr Taps taps = make_intr_taps<R, N, M>();
  constexpr int T = Taps::T;
  constexpr int TI = kIntrTileOut / R;               // whole input periods per tile (+1 when the tile starts mid-period)
  __shared__ int xs[TI + 1 + T];                     // xs[i] = x[k_start - (T-1) + i]
  __shared__ __align__(16) uint32_t ys[(TI + 1) * R + 4];

  const uint32_t c = blockIdx.y;
  const int16_t *xc = a.interleaved ? a.x + c : a.x + (size_t)c * a.n;
  const size_t xstride = a.interleaved ? a.C : 1;

  // inputs k0-(T-1) .. k0+TI of a tile, NL per thread, fetched one tile ahead so that their latency hides behind the
  // arithmetic and the stores of the current tile (history for k < n_seen, zero past the end of the call)
  constexpr int NL = (TI + T + kIntrThreads - 1) / kIntrThreads;
  int pre[NL];
  auto fetch = [&](long long tile) {
    const long long k0 = (a.out_first + tile * kIntrTileOut) / R;
#pragma unroll
    for (int q = 0; q < NL; q++) {
      const int i = threadIdx.x + q * kIntrThreads;
      const long long li = k0 - (T - 1) + i - a.n_seen;          // index into this call's input
      int v = 0;
      if (i < TI + T) {
        if (li >= 0) { if ((size_t)li < a.n) v = xc[(size_t)li * xstride]; }
        else if (li >= -(long long)a.H) v = a.tail[(size_t)c * a.H + (size_t)(a.H + li)];
      }
      pre[q] = v;
    }
  };
  if ((long long)blockIdx.x < a.ntiles) fetch(blockIdx.x);

  for (long long tile = blockIdx.x; tile < a.ntiles; tile += gridDim.x) {
    const long long j0 = tile * kIntrTileOut;                    // first local output of the tile
    const long long o0 = a.out_first + j0;                       // its global output index
    const long long k0 = o0 / R;                                 // first input period touched
    const int off = (int)(o0 - k0 * R);                          // word offset of the tile inside that period
#pragma unroll
    for (int q = 0; q < NL; q++) {
      const int i = threadIdx.x + q * kIntrThreads;
      if (i < TI + T) xs[i] = pre[q];
    }
    __syncthreads();
    if (tile + gridDim.x < a.ntiles) fetch(tile + gridDim.x);
    // ---- one thread per input period: R outputs from T inputs, all taps immediates
    for (int p = threadIdx.x; p < TI + 1; p += kIntrThreads) {
      uint32_t xv[T];
#pragma unroll
      for (int m = 0; m < T; m++) xv[m] = (uint32_t)xs[p + T - 1 - m];   // x[k - m]
      uint32_t o[R];
#pragma unroll
      for (int ph = 0; ph < R; ph++) {
        uint32_t acc = 0;
#pragma unroll
        for (int m = 0; m < T; m++)
          if (taps.v[ph + R * m] != 0) acc += taps.v[ph + R * m] * xv[m];
        o[ph] = acc;
      }
      if (R % 4 == 0) {
#pragma unroll
        for (int ph = 0; ph < R; ph += 4) *(uint4 *)(ys + p * R + ph) = make_uint4(o[ph], o[ph + 1], o[ph + 2], o[ph + 3]);
      } else {
#pragma unroll
        for (int ph = 0; ph < R; ph++) ys[p * R + ph] = o[ph];
      }
    }
    __syncthreads();
    // ---- copy out: local outputs j0 .. j0+TO-1  <-  ys[off ..]
    const long long remain = (long long)a.n_out - j0;
    const int cnt = remain < kIntrTileOut ? (int)remain : kIntrTileOut;
    const bool vec = a.ident && a.out_bytes == 4 && (a.n_out % 4 == 0 || c == 0) && ((uintptr_t)a.y & 15) == 0;
    if (vec) {
      int32_t *yo = (int32_t *)a.y + (size_t)c * a.n_out + j0;
      const int groups = cnt / 4;
      const int sub = off & 3;                                   // CTA-uniform misalignment inside a 16-byte smem word
      const uint4 *s4 = (const uint4 *)(ys + (off & ~3));
      for (int g = threadIdx.x; g < groups; g += kIntrThreads) {
        const uint4 lo = s4[g];
        uint4 w = lo;
        if (sub) {
          const uint4 hi = s4[g + 1];
          if (sub == 1) w = make_uint4(lo.y, lo.z, lo.w, hi.x);
          else if (sub == 2) w = make_uint4(lo.z, lo.w, hi.x, hi.y);
          else w = make_uint4(lo.w, hi.x, hi.y, hi.z);
        }
        *(uint4 *)(yo + 4 * g) = w;   // exact in 32 bits (lossless intW <= 32): already sign-extended
      }
      for (int j = 4 * groups + threadIdx.x; j < cnt; j += kIntrThreads) yo[j] = (int32_t)ys[off + j];
    } else {
      for (int j = threadIdx.x; j < cnt; j += kIntrThreads) cic_intr_store_converted(a, c, (size_t)(j0 + j), ys[off + j]);
    }
    __syncthreads();
  }
}

// ------------------------------------------------------------------------------------------ R = 4, no staging
// BASELINE config 5 geometry.  One thread = one 16-byte group of 4 consecutive outputs of THIS call's array, i.e.
// phases A..A+3 of one or two input periods, A = out_first mod 4 being uniform over the launch (template parameter).
// A warp walks U x 32 consecutive groups; the T (+1) input samples a thread needs are its own one plus those of its
// left neighbours, taken with warp shuffles from the current and the previous iteration's registers -- no shared
// memory, one 2-byte load and one 128-bit store per group.  Outputs are exact in 32 bits (the lossless width of
// find_inter_type_cic_intr is <= 32), so no sign-extension pass is needed.
constexpr int kDirThreads = 256;
constexpr int kDirU = 8;            // groups per thread

template <int N, int M, int A>
__global__ void __launch_bounds__(kDirThreads) cic_intr4_direct_kernel(CicIntrArgs a) {
  constexpr int R = 4;
  typedef IntrTaps<R, N, M> Taps;
  constexpr Taps taps = make_intr_taps<R, N, M>();
  constexpr int T = Taps::T;
  constexpr int D = A ? 1 : 0;                 // the group reaches into the next period
  constexpr int NX = T + D;                    // samples per group: x[kl], x[kl-1], .., x[kl-NX+1]
  static_assert(NX <= 32, "window exceeds a warp");
  const uint32_t c = blockIdx.y;
  const int lane = threadIdx.x & 31;
  const int16_t *xc = a.interleaved ? a.x + c : a.x + (size_t)c * a.n;
  const size_t xstride = a.interleaved ? a.C : 1;
  const int16_t *tl = a.tail + (size_t)c * a.H;
  int32_t *yc = (int32_t *)a.y + (size_t)c * a.n_out;
  const long long ngroups = (long long)(a.n_out / 4);
  const long long kfirst0 = a.out_first / 4;   // period of output group 0

  auto load_x = [&](long long k) -> uint32_t {   // sample with global index k; history / zero outside this call
    const long long li = k - a.n_seen;
    int v = 0;
    if (li >= 0) { if ((size_t)li < a.n) v = xc[(size_t)li * xstride]; }
    else if (li >= -(long long)a.H) v = tl[a.H + li];
    return (uint32_t)v;
  };

  const long long warps_total = (long long)gridDim.x * (kDirThreads / 32);
  const long long warp_id = (long long)blockIdx.x * (kDirThreads / 32) + (threadIdx.x >> 5);
  for (long long g0 = warp_id * (32 * kDirU); g0 < ngroups; g0 += warps_total * (32 * kDirU)) {
    const long long kl0 = kfirst0 + g0 + D;                       // sample index of lane 0, iteration 0
    // unchecked loads when the whole chunk (with its left halo) lies inside this call's input
    const bool interior = kl0 - 32 >= a.n_seen && (size_t)(kl0 + 32 * kDirU - a.n_seen) <= a.n;
    const int16_t *xp = xc + (size_t)(kl0 - a.n_seen) * xstride;  // only dereferenced when interior
    uint32_t prev = interior ? (uint32_t)(int)xp[((long long)lane - 32) * (long long)xstride] : load_x(kl0 - 32 + lane);
    uint32_t cur[kDirU];
#pragma unroll
    for (int u = 0; u < kDirU; u++)
      cur[u] = interior ? (uint32_t)(int)xp[(size_t)(32 * u + lane) * xstride] : load_x(kl0 + 32 * u + lane);
#pragma unroll
    for (int u = 0; u < kDirU; u++) {
      uint32_t xv[NX];
      xv[0] = cur[u];
#pragma unroll
      for (int m = 1; m < NX; m++) {
        // x[k - m] sits m lanes to the left: in this iteration's registers, or (for the first m lanes) in the last m
        // lanes of the previous iteration's
        const uint32_t send = lane >= 32 - m ? prev : cur[u];
        xv[m] = __shfl_sync(0xffffffffu, send, (lane - m) & 31);
      }
      prev = cur[u];
      uint32_t o[4];
#pragma unroll
      for (int i = 0; i < 4; i++) {
        const int ph = (A + i) % 4;
        const int d = D - (A + i) / 4;           // samples between this output's period and x[kl]
        uint32_t acc = 0;
#pragma unroll
        for (int m = 0; m < T; m++)
          if (taps.v[ph + R * m] != 0) acc += taps.v[ph + R * m] * xv[d + m];
        o[i] = acc;
      }
      const long long g = g0 + 32 * u + lane;
      if (g < ngroups) *(uint4 *)(yc + 4 * g) = make_uint4(o[0], o[1], o[2], o[3]);
    }
  }
  // the last n_out % 4 outputs of the call (one thread each)
  const int rem = (int)(a.n_out & 3);
  if (blockIdx.x == 0 && (int)threadIdx.x < rem) {
    const long long j = ngroups * 4 + threadIdx.x, o = a.out_first + j, k = o / 4;
    const int ph = (int)(o - 4 * k);
    uint32_t acc = 0;
    for (int m = 0; m < T; m++) acc += taps.v[ph + R * m] * load_x(k - m);
    yc[j] = (int32_t)acc;
  }
}

template <int N, int M>
static cudaError_t launch_intr4_direct(const CicIntrArgs &a, cudaStream_t st) {
  const long long ngroups = (long long)(a.n_out / 4);
  const long long per_cta = (long long)kDirThreads * kDirU;
  long long gx = (ngroups + per_cta - 1) / per_cta;
  if (gx < 1) gx = 1;
  const long long cap = (148LL * 16 + a.C - 1) / a.C;
  if (gx > cap) gx = cap;
  dim3 grid((unsigned)gx, a.C);
  switch ((int)(a.out_first & 3)) {
    case 0: cic_intr4_direct_kernel<N, M, 0><<<grid, kDirThreads, 0, st>>>(a); break;
    case 1: cic_intr4_direct_kernel<N, M, 1><<<grid, kDirThreads, 0, st>>>(a); break;
    case 2: cic_intr4_direct_kernel<N, M, 2><<<grid, kDirThreads, 0, st>>>(a); break;
    default: cic_intr4_direct_kernel<N, M, 3><<<grid, kDirThreads, 0, st>>>(a); break;
  }
  return cudaGetLastError();
}

template <int R, int N, int M>
static cudaError_t launch_intr(CicIntrArgs a, cudaStream_t st) {
  a.ntiles = (long long)((a.n_out + kIntrTileOut - 1) / kIntrTileOut);
  long long gx = a.ntiles;
  const long long cap = (148LL * 8 + a.C - 1) / a.C;
  if (gx > cap) gx = cap;
  dim3 grid((unsigned)gx, a.C);
  cic_intr_fast_kernel<R, N, M><<<grid, kIntrThreads, 0, st>>>(a);
  return cudaGetLastError();
}

#define B2D_CIC_INTR_CASES(X) \
  X(4, 3, 1) X(4, 4, 1) X(4, 3, 2) X(2, 3, 1) X(2, 4, 1) X(8, 3, 1) X(8, 4, 1) X(8, 5, 1) X(8, 4, 2) X(16, 3, 1) X(16, 4, 1)

bool cic_intr_fast_supported(const CicLaunch &p) {
  if (!p.intr) return false;
  if (p.fin.W > 16 || p.intW > 32) return false;
  if (!p.fin.S && p.fin.W == 16) return false;   // samples are sign-extended from their int16 container
#define X(r, n, m) if (p.R == r && p.N == n && p.M == m) return true;
  B2D_CIC_INTR_CASES(X)
#undef X
  return false;
}

cudaError_t launch_cic_intr_fast(const CicLaunch &p, cudaStream_t st) {
  if (p.n_out == 0) return cudaSuccess;
  CicIntrArgs a;
  a.x = (const int16_t *)p.in; a.y = p.out; a.tail = (const int16_t *)p.tail; a.n = p.n; a.n_out = p.n_out;
  a.n_seen = (long long)p.n_seen; a.out_first = (long long)p.out_first; a.H = p.H; a.C = p.C; a.interleaved = p.interleaved && p.C > 1;
  a.intW = p.intW; a.in = p.fin; a.out = p.fout; a.out_bytes = container_bytes(p.fout.W);
  a.ident = (p.fout.F() == p.fin.F() && p.fout.W == p.intW && p.fout.S == 1) ? 1 : 0;
  a.ntiles = 0;
  const bool direct_ok = p.R == 4 && a.ident && a.out_bytes == 4 && (a.n_out % 4 == 0 || a.C == 1) && ((uintptr_t)a.y & 15) == 0 &&
                         !(getenv("B2D_CIC_INTR_STAGED") && *getenv("B2D_CIC_INTR_STAGED") == '1');
  if (direct_ok) {
    if (p.N == 3 && p.M == 1) return launch_intr4_direct<3, 1>(a, st);
    if (p.N == 4 && p.M == 1) return launch_intr4_direct<4, 1>(a, st);
    if (p.N == 3 && p.M == 2) return launch_intr4_direct<3, 2>(a, st);
  }
#define X(r, n, m) if (p.R == r && p.N == n && p.M == m) return launch_intr<r, n, m>(a, st);
  B2D_CIC_INTR_CASES(X)
#undef X
  return cudaErrorNotSupported;
}

}  // namespace b2d
