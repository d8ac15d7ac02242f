// cic_fast.cu -- HBM-bound CIC decimator for 16-bit samples and <= 32-bit lossless state
// (BASELINE config 3: ac_cic_dec_full R=8 N=4 ac_fixed<16,1> -> <28,13>, interleaved IQ).
//
// Same mathematics as cic_generic.cu (reference ac_cic_full_core.h:80-87,110-135,198-255, ac_cic_dec_full.h:163-222):
// every thread restarts the N-stage integrator from zero N*M low-rate samples before its first output and lets the
// N combs annihilate the resulting degree-(N-1) polynomial error; all arithmetic is modulo 2^32, which contains the
// reference's modulo-2^intW arithmetic because intW <= 32.
//
// Layout.  A tile is ROWS rows of L = K*R consecutive samples (per channel); thread (row, channel) owns the K outputs
// of its row and runs its integrators through the last N*M*R + N-1 samples of the previous row first.  Because the
// pipelined integrator of the reference emits on inputs 0, R, 2R, .. the row boundaries coincide with the emission
// cadence and every shared-memory offset is a compile-time immediate.  Rows are staged from HBM with 16-byte cp.async
// (coalesced, no register round trip) into rows padded by 16 bytes so that the 128-bit row reads of the 8 threads of
// a quarter-warp fall into distinct bank groups; CTAs are persistent and double-buffered (tile i+1 streams in while
// tile i integrates).  Per input sample and channel: N adds + 1 unpack + 1/4 (1/8) LDS.128 -- far below the issue
// budget at the HBM rate, so the kernel is bandwidth-bound by construction.
#include "kernels.h"

namespace b2d {

constexpr int kCicThreads = 128;

struct CicFastArgs {
  const unsigned char *x;   // input samples
  void *y;                  // outputs, planar, stride n_out
  const int16_t *tail;      // [C][H] history
  size_t n, n_out;
  int H;
  uint32_t C;
  int intW;
  Fmt in, out;
  int out_bytes, ident;
  long long ntiles;
};

__device__ __forceinline__ void cp_async16(void *smem_dst, const void *gsrc) {
  const uint32_t d = (uint32_t)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

template <int R, int N, int M, int K, int CT>
struct CicGeom {
  static constexpr int L = K * R;                    // samples per row per channel
  static constexpr int ROWB = L * CT * 2;            // bytes per row
  static constexpr int S = ROWB + 16;                // padded row stride
  static constexpr int ROWS = kCicThreads / CT;      // rows per tile
  static constexpr int H = N * M * R + N - 1;        // run-in samples taken from the previous row
  static constexpr int SPC = 8 / CT;                 // samples (per channel) per 16-byte chunk
  static constexpr int BUF = (ROWS + 1) * S;         // bytes per buffer (row -1 .. ROWS-1)
  static_assert(ROWB % 16 == 0 && (ROWB / 16) % 2 == 0, "row must be an even number of 16-byte chunks");
  static_assert(L >= H, "run-in must fit in one row");
};

// stage rows -1 .. ROWS-1 of `tile` (channel block c0) into buf
template <int R, int N, int M, int K, int CT>
__device__ __noinline__ void cic_stage_edge(const CicFastArgs &a, long long tile, uint32_t c0, unsigned char *buf);

template <int R, int N, int M, int K, int CT>
__device__ __forceinline__ void cic_stage(const CicFastArgs &a, long long tile, uint32_t c0, unsigned char *buf) {
  typedef CicGeom<R, N, M, K, CT> G;
  constexpr int CPR = G::ROWB / 16;                  // chunks per row
  const long long s0 = (tile * G::ROWS - 1) * (long long)G::L;   // first sample (per channel) of row -1
  const unsigned char *gbase = a.x + (CT == 2 ? 0 : (size_t)c0 * a.n * 2);
  const bool full = tile > 0 && (size_t)((tile + 1) * G::ROWS * (long long)G::L) <= a.n;
  if (full) {
    const unsigned char *g = gbase + (size_t)s0 * CT * 2;
    for (int q = threadIdx.x; q < (G::ROWS + 1) * CPR; q += kCicThreads)
      cp_async16(buf + (q / CPR) * G::S + (q % CPR) * 16, g + (size_t)q * 16);
    return;
  }
  cic_stage_edge<R, N, M, K, CT>(a, tile, c0, buf);
}

// first tile of a call (history row from the carried tail) and the ragged last tile: guarded, element-wise where needed
template <int R, int N, int M, int K, int CT>
__device__ __noinline__ void cic_stage_edge(const CicFastArgs &a, long long tile, uint32_t c0, unsigned char *buf) {
  typedef CicGeom<R, N, M, K, CT> G;
  constexpr int CPR = G::ROWB / 16;
  const long long s0 = (tile * G::ROWS - 1) * (long long)G::L;
  const unsigned char *gbase = a.x + (CT == 2 ? 0 : (size_t)c0 * a.n * 2);
  for (int q = threadIdx.x; q < (G::ROWS + 1) * CPR; q += kCicThreads) {
    const long long s = s0 + (long long)q * G::SPC;  // first sample of this chunk
    unsigned char *dst = buf + (q / CPR) * G::S + (q % CPR) * 16;
    if (s >= 0 && (size_t)(s + G::SPC) <= a.n) {
      cp_async16(dst, gbase + (size_t)s * CT * 2);
    } else {
      int16_t *d16 = (int16_t *)dst;
      for (int e = 0; e < G::SPC; e++)
        for (int ch = 0; ch < CT; ch++) {
          const long long g = s + e;
          int16_t v = 0;
          if (g >= 0 && (size_t)g < a.n) v = ((const int16_t *)gbase)[(size_t)g * CT + ch];
          else if (g < 0 && g >= -(long long)a.H) v = a.tail[(size_t)(c0 + ch) * a.H + (size_t)(a.H + g)];
          d16[e * CT + ch] = v;
        }
    }
  }
}

// sample e (0 .. SPC-1) of this thread's channel out of a 16-byte chunk, sign-extended with one PRMT
// (selector nibble bit 3 = replicate the sign of the selected byte); `sel` picks the low or high half for CT == 2.
template <int CT>
__device__ __forceinline__ int cic_pick(const uint4 &v, int e, uint32_t sel) {
  const int wi = CT == 2 ? e : e / 2;
  const uint32_t w = wi == 0 ? v.x : wi == 1 ? v.y : wi == 2 ? v.z : v.w;
  int d;
  if (CT == 2) asm("prmt.b32 %0, %1, %1, %2;" : "=r"(d) : "r"(w), "r"(sel));
  else if (e & 1) d = (int)w >> 16;
  else asm("prmt.b32 %0, %1, %1, 0x9910;" : "=r"(d) : "r"(w));
  return d;
}

// converting epilogue (OUT_TYPE != lossless INT_TYPE): rare, kept out of line so the unrolled hot loop stays small
__device__ __noinline__ void cic_store_converted(const CicFastArgs &a, uint32_t c, size_t j, uint32_t raw) {
  const int64_t w = wrap_bits((int64_t)raw, a.intW, 1);
  store_raw(a.y, (size_t)c * a.n_out + j, a.out_bytes, a.ident ? w : convert((i128)w, a.in.F(), a.out));
}

template <int R, int N, int M, int K, int CT>
__global__ void __launch_bounds__(kCicThreads, 3) cic_dec_fast_kernel(CicFastArgs a) {
  typedef CicGeom<R, N, M, K, CT> G;
  extern __shared__ __align__(16) unsigned char smem[];
  const uint32_t c0 = CT == 2 ? 0 : blockIdx.y;
  const int row = threadIdx.x / CT, ch = threadIdx.x % CT;

  long long tile = blockIdx.x;
  if (tile < a.ntiles) cic_stage<R, N, M, K, CT>(a, tile, c0, smem);
  cp_async_commit();
  for (int it = 0; tile < a.ntiles; tile += gridDim.x, it++) {
    unsigned char *buf = smem + (it & 1) * G::BUF;
    const long long next = tile + gridDim.x;
    if (next < a.ntiles) cic_stage<R, N, M, K, CT>(a, next, c0, smem + ((it + 1) & 1) * G::BUF);
    cp_async_commit();
    cp_async_wait<1>();
    __syncthreads();

    const long long mo0 = (tile * G::ROWS + row) * (long long)K;   // first output of this thread
    if ((size_t)mo0 < a.n_out) {
      uint32_t r[N], d[N][M], res[K];
      const uint32_t sel = ch ? 0xBB32u : 0x9910u;
#pragma unroll
      for (int i = 0; i < N; i++) {
        r[i] = 0;
#pragma unroll
        for (int j = 0; j < M; j++) d[i][j] = 0;
      }
      // phase 0: run-in over the tail of the previous row (buffer row `row`), phase 1: own row (buffer row `row + 1`)
#pragma unroll
      for (int ph = 0; ph < 2; ph++) {
        const unsigned char *p = buf + (row + ph) * G::S;
        constexpr int kFirstChunk0 = (G::L - G::H) / G::SPC;
        const int q_begin = ph == 0 ? kFirstChunk0 : 0;
#pragma unroll
        for (int q = 0; q < G::ROWB / 16; q++) {
          if (q < q_begin) continue;
          const uint4 v = *(const uint4 *)(p + q * 16);
#pragma unroll
          for (int e = 0; e < G::SPC; e++) {
            const int col = q * G::SPC + e;
            if (ph == 0 && col < G::L - G::H) continue;
            const uint32_t xs = (uint32_t)cic_pick<CT>(v, e, sel);
            // pipelined integrator (intStage, ac_cic_full_core.h:80-87): stage i adds stage i-1's OLD value
#pragma unroll
            for (int i = N - 1; i > 0; i--) r[i] += r[i - 1];
            r[0] += xs;
            if (col % R == 0 && (ph == 1 || col >= G::L - N * M * R)) {
              // comb chain (diffStage, ac_cic_full_core.h:246-255)
              uint32_t w = r[N - 1];
#pragma unroll
              for (int k = 0; k < N; k++) {
                const uint32_t o = w - d[k][M - 1];
#pragma unroll
                for (int j = M - 1; j > 0; j--) d[k][j] = d[k][j - 1];
                d[k][0] = w;
                w = o;
              }
              if (ph == 1) res[col / R] = w;
            }
          }
        }
      }
      // ---- outputs: K consecutive values of channel c0 + ch
      const uint32_t c = c0 + ch;
      if (a.ident && a.out_bytes == 4 && (size_t)(mo0 + K) <= a.n_out && (a.n_out % 4 == 0 || c == 0)) {
        int4 *yo = (int4 *)((int32_t *)a.y + (size_t)c * a.n_out + mo0);
        const int sh = 32 - a.intW;
#pragma unroll
        for (int i = 0; i < K; i += 4) {
          int4 o;
          o.x = (int)(res[i] << sh) >> sh; o.y = (int)(res[i + 1] << sh) >> sh;
          o.z = (int)(res[i + 2] << sh) >> sh; o.w = (int)(res[i + 3] << sh) >> sh;
          yo[i / 4] = o;
        }
      } else {
#pragma unroll
        for (int i = 0; i < K; i++)
          if ((size_t)(mo0 + i) < a.n_out) cic_store_converted(a, c, (size_t)(mo0 + i), res[i]);
      }
    }
    __syncthreads();
  }
  cp_async_wait<0>();
}

// --------------------------------------------------------------------------------------------- dispatch
template <int R, int N, int M, int K, int CT>
static cudaError_t launch_dec(const CicFastArgs &a, uint32_t C, cudaStream_t st) {
  typedef CicGeom<R, N, M, K, CT> G;
  CicFastArgs b = a;
  const size_t per_tile = (size_t)G::ROWS * G::L;
  b.ntiles = (long long)((a.n + per_tile - 1) / per_tile);
  const size_t smem = 2 * (size_t)G::BUF;
  cudaError_t e = cudaFuncSetAttribute(cic_dec_fast_kernel<R, N, M, K, CT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  long long gx = 148 * 3;
  if (CT == 1 && C > 1) gx = (gx + C - 1) / C;
  if (gx > b.ntiles) gx = b.ntiles;
  dim3 grid((unsigned)gx, CT == 2 ? 1 : C);
  cic_dec_fast_kernel<R, N, M, K, CT><<<grid, kCicThreads, smem, st>>>(b);
  return cudaGetLastError();
}

#define B2D_CIC_DEC_CASES(X) \
  X(8, 4, 1, 16) X(8, 4, 2, 16) X(8, 3, 1, 16) X(8, 5, 1, 16) X(4, 3, 1, 32) X(4, 4, 1, 32) X(16, 4, 1, 8) X(16, 3, 1, 8) X(2, 3, 1, 64)

static bool dec_case_exists(int R, int N, int M) {
#define X(r, n, m, k) if (R == r && N == n && M == m) return true;
  B2D_CIC_DEC_CASES(X)
#undef X
  return false;
}

bool cic_fast_supported(const CicLaunch &p) {
  if (p.intr) return false;
  if (p.fin.W > 16 || p.intW > 32) return false;
  if (!p.fin.S && p.fin.W == 16) return false;   // samples are sign-extended from their int16 container
  if (p.interleaved && p.C != 2 && p.C != 1) return false;
  return dec_case_exists(p.R, p.N, p.M);
}

cudaError_t launch_cic_fast(const CicLaunch &p, cudaStream_t st) {
  // the emission cadence must coincide with the row grid: calls that start mid-period take the generic kernel
  const bool aligned = (p.n_seen % (unsigned long long)p.R) == 0 && (((uintptr_t)p.in | (uintptr_t)p.out) & 15) == 0 &&
                       (p.interleaved || p.C == 1 || (p.n % 8) == 0);
  if (!aligned) return launch_cic_generic(p, st);
  if (p.n_out == 0) return cudaSuccess;
  CicFastArgs a;
  a.x = (const unsigned char *)p.in; a.y = p.out; a.tail = (const int16_t *)p.tail; a.n = p.n; a.n_out = p.n_out;
  a.H = p.H; a.C = p.C; a.intW = p.intW; a.in = p.fin; a.out = p.fout; a.out_bytes = container_bytes(p.fout.W);
  a.ident = (p.fout.F() == p.fin.F() && p.fout.W == p.intW && p.fout.S == 1) ? 1 : 0;
  a.ntiles = 0;
  const int ct = (p.interleaved && p.C == 2) ? 2 : 1;
#define X(r, n, m, k)                                                                     \
  if (p.R == r && p.N == n && p.M == m)                                                   \
    return ct == 2 ? launch_dec<r, n, m, k, 2>(a, p.C, st) : launch_dec<r, n, m, k, 1>(a, p.C, st);
  B2D_CIC_DEC_CASES(X)
#undef X
  return cudaErrorNotSupported;
}

}  // namespace b2d
