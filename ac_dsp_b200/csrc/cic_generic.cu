// cic_generic.cu -- any-width CIC decimator / interpolator (state up to 64 bits, run-time R, M, N).
//
// What the reference computes (include/ac_dsp/ac_cic_full_core.h, all on the lossless INT_TYPE with
// AC_TRN / AC_WRAP, i.e. modular integer arithmetic of width intW):
//   decimator   ac_cic_dec_full.h:163-222   u[m] = (S^N x)[m R - (N-1)]   (intStage :80-87 is a pipelined
//               N-stage integrator: N-1 samples of latency; decIntgCore :110-135 forwards inputs 0, R, 2R, ..)
//               out[m] = (Delta_M^N u)[m]                                  (comb / diffStage :228-255)
//   interpolator ac_cic_intr_full.h:150-215  c = Delta_M^N x;  z[kR] = c[k], 0 elsewhere  (intrIntgCore :143-160)
//               out[j] = (S^N z)[j]      (first N-1 integrator outputs dropped, :209-213)
// S = inclusive running sum, Delta_M w[m] = w[m] - w[m-M]; everything before the stream start is 0.
//
// Parallel form.  The cascade as a whole is an FIR filter (boxcar(RM)^N), so an output depends on a finite
// window of inputs, although the integrators alone have unbounded memory.  A thread therefore restarts the
// recursion from an all-zero state N*M low-rate samples before its first output: the error this introduces in
// the integrator outputs is a polynomial of degree < N in the low-rate index, which the N combs annihilate
// exactly (also modulo 2^64).  No cross-thread scan, no carried integrator registers: the only state a handle
// carries between run() calls is the last H inputs and the input count.
#include "kernels.h"

namespace b2d {

constexpr int kMaxN = 16;
constexpr int kMaxNM = 64;

int cic_history_len(int intr, int R, int M, int N) { return intr ? N * M + N + 2 : N * M * R + N - 1; }

struct CicGenArgs {
  Fmt in, out;
  int intW, R, M, N;
  uint32_t C;
  int interleaved, in_bytes, out_bytes, H, K;  // K = outputs per thread
  const void *x;
  void *y;
  size_t n, n_out;
  long long n_seen, out_first;
  const void *tail;
};

// input sample with global index g (>= 0): from this call's buffer or the carried history
__device__ __forceinline__ uint64_t cic_sample(const CicGenArgs &a, uint32_t c, long long g) {
  const long long i = g - a.n_seen;
  if (i >= 0) return (uint64_t)load_raw(a.x, elem_index((size_t)i, c, a.n, a.C, a.interleaved), a.in_bytes, a.in.S);
  if (i < -(long long)a.H) return 0;
  return (uint64_t)load_raw(a.tail, (size_t)c * a.H + (size_t)(a.H + i), a.in_bytes, a.in.S);
}

__device__ __forceinline__ uint64_t comb_chain(uint64_t v, uint64_t *d, int N, int M) {
  for (int k = 0; k < N; k++) {
    uint64_t *dl = d + k * M;
    const uint64_t o = v - dl[M - 1];
    for (int i = M - 1; i > 0; i--) dl[i] = dl[i - 1];
    dl[0] = v;
    v = o;
  }
  return v;
}

__device__ __forceinline__ void cic_store(const CicGenArgs &a, uint32_t c, size_t j, uint64_t v) {
  const int64_t w = wrap_bits((int64_t)v, a.intW, 1);
  store_raw(a.y, (size_t)c * a.n_out + j, a.out_bytes, convert((i128)w, a.in.F(), a.out));
}

__global__ void __launch_bounds__(128) cic_dec_generic_kernel(CicGenArgs a) {
  const size_t per_ch = (a.n_out + a.K - 1) / a.K;
  const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= per_ch * a.C) return;
  const uint32_t c = (uint32_t)(t / per_ch);
  const size_t j0 = (t % per_ch) * a.K;
  const size_t j1 = (j0 + a.K < a.n_out) ? j0 + a.K : a.n_out;
  const long long m0 = a.out_first + (long long)j0;
  uint64_t r[kMaxN], d[kMaxNM];
  for (int i = 0; i < a.N; i++) r[i] = 0;
  for (int i = 0; i < a.N * a.M; i++) d[i] = 0;
  long long ms = m0 - (long long)a.N * a.M;  // first low-rate index fed to the combs
  if (ms < 0) ms = 0;
  long long g = ms * a.R - (a.N - 1);        // first input fed to the integrators
  if (g < 0) g = 0;
  for (long long m = ms; m < a.out_first + (long long)j1; m++) {
    const long long gend = m * a.R - (a.N - 1);  // u[m] = (S^N x)[gend]
    for (; g <= gend; g++) {
      uint64_t v = cic_sample(a, c, g);
      for (int i = 0; i < a.N; i++) { r[i] += v; v = r[i]; }
    }
    const uint64_t w = comb_chain(gend < 0 ? 0 : r[a.N - 1], d, a.N, a.M);
    if (m >= m0) cic_store(a, c, (size_t)(m - a.out_first), w);
  }
}

__global__ void __launch_bounds__(128) cic_intr_generic_kernel(CicGenArgs a) {
  const size_t per_ch = (a.n_out + a.K - 1) / a.K;
  const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= per_ch * a.C) return;
  const uint32_t c = (uint32_t)(t / per_ch);
  const size_t j0 = (t % per_ch) * a.K;
  const size_t j1 = (j0 + a.K < a.n_out) ? j0 + a.K : a.n_out;
  const long long o0 = a.out_first + (long long)j0, o1 = a.out_first + (long long)j1;
  uint64_t r[kMaxN], d[kMaxNM];
  for (int i = 0; i < a.N; i++) r[i] = 0;
  for (int i = 0; i < a.N * a.M; i++) d[i] = 0;
  // the filter spans N(RM-1)+1 high-rate samples: restart N*M (+1) inputs before the first output's input
  long long k = o0 / a.R - (long long)a.N * a.M - 1;
  if (k < 0) k = 0;
  for (;; k++) {
    const uint64_t cv = comb_chain(cic_sample(a, c, k), d, a.N, a.M);
    for (int ph = 0; ph < a.R; ph++) {
      const long long j = k * a.R + ph;
      if (j >= o1) return;
      uint64_t v = ph == 0 ? cv : 0;
      for (int i = 0; i < a.N; i++) { r[i] += v; v = r[i]; }
      if (j >= o0) cic_store(a, c, (size_t)(j - a.out_first), v);
    }
  }
}

cudaError_t launch_cic_generic(const CicLaunch &p, cudaStream_t st) {
  if (p.n_out == 0) return cudaSuccess;
  CicGenArgs a;
  a.in = p.fin; a.out = p.fout; a.intW = p.intW; a.R = p.R; a.M = p.M; a.N = p.N;
  a.C = p.C; a.interleaved = p.interleaved;
  a.in_bytes = container_bytes(p.fin.W); a.out_bytes = container_bytes(p.fout.W);
  a.H = p.H; a.x = p.in; a.y = p.out; a.n = p.n; a.n_out = p.n_out;
  a.n_seen = (long long)p.n_seen; a.out_first = (long long)p.out_first; a.tail = p.tail;
  const int nm = p.N * p.M;
  a.K = p.intr ? 8 * nm * p.R : 8 * nm;
  if (a.K < 32) a.K = 32;
  const size_t per_ch = (p.n_out + a.K - 1) / a.K;
  const size_t threads = per_ch * p.C;
  const unsigned blocks = (unsigned)((threads + 127) / 128);
  if (p.intr) cic_intr_generic_kernel<<<blocks, 128, 0, st>>>(a);
  else cic_dec_generic_kernel<<<blocks, 128, 0, st>>>(a);
  return cudaGetLastError();
}

}  // namespace b2d
