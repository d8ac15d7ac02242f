// cic_fast.cu -- HBM-bound CIC decimator for 16-bit samples and <= 32-bit lossless state
// (BASELINE config 3: ac_cic_dec_full R=8 N=4 ac_fixed<16,1> -> <28,13>, interleaved IQ).
//
// Same mathematics as cic_generic.cu (reference ac_cic_full_core.h:80-87,110-135,198-255, ac_cic_dec_full.h:163-222):
// every thread restarts the N-stage integrator from zero N*M low-rate samples before its first output and lets the
// N combs annihilate the resulting degree-(N-1) polynomial error; all arithmetic is modulo 2^32, which contains the
// reference's modulo-2^intW arithmetic because intW <= 32.
//
// BASELINE config 3 itself (R = 8, N = 4, M = 1) does not take this recursion: cic8n4_row below evaluates the same
// filter as three decimating half-band stages (see there); the recursion serves the other instantiated geometries.
//
// Layout.  A tile is ROWS rows of L = K*R consecutive samples (per channel); thread (row, channel) owns the K outputs
// of its row and runs its integrators through the last N*M*R + N-1 samples of the previous row first.  Because the
// pipelined integrator of the reference emits on inputs 0, R, 2R, .. the row boundaries coincide with the emission
// cadence and every shared-memory offset is a compile-time immediate.  Rows are staged from HBM with 16-byte cp.async
// (coalesced, no register round trip) into rows padded by 16 bytes so that the 128-bit row reads of the 8 threads of
// a quarter-warp fall into distinct bank groups; CTAs are persistent and double-buffered (tile i+1 streams in while
// tile i integrates).  Per input sample and channel: N adds + 1 unpack + 1/4 (1/8) LDS.128 -- far below the issue
// budget at the HBM rate, so the kernel is bandwidth-bound by construction.
#include <cstdlib>

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
  int recursive;            // force the integrator / comb form where the non-recursive one exists (tests)
  int nstage;               // tile buffers in the cp.async ring (2..4)
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

// ---- R = 8, N = 4, M = 1 without recursion -------------------------------------------------------------------------
// boxcar(8)^4 = (1+z^-1)^4 (1+z^-2)^4 (1+z^-4)^4: three cascaded [1 4 6 4 1] half-band stages, each followed by a
// decimation by 2.  With v0 = x and  v_{s+1}[q] = sum_t b[t] * v_s[2q + phi_s - t],  phi = (1, 0, -1), the third stage
// is exactly the reference's output  out[m] = (boxcar(8)^4 * x)[8m - 3]  (integrate x4, keep every 8th from the pipelined
// integrator, comb x4: ac_cic_full_core.h:80-135,198-255) -- no integrator state, no run-in, no comb pass, no wrap
// (the lossless width is <= 32).  The first stage runs on the packed 16-bit samples with DP2A (two taps per
// instruction, taps as bytes {1,4},{6,4},{1,0}); stages two and three are 4 integer ops per value at 1/4 and 1/8 of the
// input rate: about 3.75 instructions per input sample against 8.5 for the recursive form, which is what frees the
// kernel from its instruction ceiling.  Words: CT == 2: word(t) = (I_t, Q_t);  CT == 1: word(j) = (x[2j], x[2j+1]).
__device__ __forceinline__ int cic_dp2a(uint32_t a, uint32_t b, int c) {
  int d; asm("dp2a.lo.s32.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c)); return d;
}
__device__ __forceinline__ uint32_t cic_half(uint32_t a, uint32_t b, uint32_t c, uint32_t d, uint32_t e) {
  return a + e + (c << 1) + ((b + d + c) << 2);        // a + 4b + 6c + 4d + e
}

template <int CT, int K>
__device__ __forceinline__ void cic8n4_row(const unsigned char *prev_row, const unsigned char *own_row, int ch, uint32_t (&res)[K]) {
  constexpr int L = K * 8;                             // samples per row and channel
  constexpr int WPR = CT == 2 ? L : L / 2;             // 32-bit words per row
  constexpr int WPI = CT == 2 ? 8 : 4;                 // words consumed per output
  const uint32_t sel = CT == 2 ? (ch ? 0x7632u : 0x5410u) : 0x5432u;
  // P[q] = (x[2q-1], x[2q]) packed, V1[p], V2[q]: static indices, everything below is fully unrolled
  uint32_t P[4 * K + 16], V1[4 * K + 15], V2[2 * K + 6];
  const uint4 *g0 = (const uint4 *)prev_row, *g1 = (const uint4 *)own_row;
  uint32_t last = 0;                                   // the word just before the current group
#pragma unroll
  for (int i = -4; i < K - 1; i++) {
    // words i*WPI .. i*WPI + WPI-1 (negative indices live at the end of the previous row)
    uint32_t w[WPI];
#pragma unroll
    for (int gq = 0; gq < WPI / 4; gq++) {
      const int wi = i * WPI + 4 * gq;                 // first word of this 16-byte group
      const uint4 v = wi < 0 ? g0[(WPR + wi) / 4] : g1[wi / 4];
      w[4 * gq] = v.x; w[4 * gq + 1] = v.y; w[4 * gq + 2] = v.z; w[4 * gq + 3] = v.w;
    }
    // pairs P[4i .. 4i+3]
#pragma unroll
    for (int e = 0; e < 4; e++) {
      const int q = 4 * i + e;
      uint32_t A, B;
      if (CT == 2) { A = e == 0 ? last : w[2 * e - 1]; B = w[2 * e]; }
      else { A = e == 0 ? last : w[e - 1]; B = w[e]; }
      if (q >= -15) P[q + 16] = __byte_perm(A, B, sel);
    }
    last = w[WPI - 1];
    // first stage: v1[p] = x[2p-3] + 4 x[2p-2] + 6 x[2p-1] + 4 x[2p] + x[2p+1],  p = 4i-1 .. 4i+2
#pragma unroll
    for (int e = 0; e < 4; e++) {
      const int p = 4 * i - 1 + e;
      if (p >= -14) V1[p + 14] = (uint32_t)cic_dp2a(P[p + 17], 0x0001u, cic_dp2a(P[p + 16], 0x0406u, cic_dp2a(P[p + 15], 0x0401u, 0)));
    }
    // second stage: v2[q] from v1[2q-4 .. 2q],  q = 2i, 2i+1
#pragma unroll
    for (int e = 0; e < 2; e++) {
      const int q = 2 * i + e;
      if (q >= -5) V2[q + 5] = cic_half(V1[2 * q - 4 + 14], V1[2 * q - 3 + 14], V1[2 * q - 2 + 14], V1[2 * q - 1 + 14], V1[2 * q + 14]);
    }
    // third stage: out[i+1] from v2[2(i+1)-5 .. 2(i+1)-1]
    if (i + 1 >= 0) {
      const int m = i + 1;
      res[m] = cic_half(V2[2 * m - 5 + 5], V2[2 * m - 4 + 5], V2[2 * m - 3 + 5], V2[2 * m - 2 + 5], V2[2 * m - 1 + 5]);
    }
  }
}

template <int R, int N, int M, int K, int CT>
__global__ void __launch_bounds__(kCicThreads, 3) cic_dec_fast_kernel(CicFastArgs a) {
  typedef CicGeom<R, N, M, K, CT> G;
  extern __shared__ __align__(16) unsigned char smem[];
  const uint32_t c0 = CT == 2 ? 0 : blockIdx.y;
  const int row = threadIdx.x / CT, ch = threadIdx.x % CT;

  // ns-deep ring of tile buffers: tiles it+1 .. it+ns-1 stream in while tile it is evaluated (one commit group per tile)
  const int ns = a.nstage;
  long long tile = blockIdx.x;
  for (int s = 0; s < ns - 1; s++) {
    const long long t = tile + (long long)s * gridDim.x;
    if (t < a.ntiles) cic_stage<R, N, M, K, CT>(a, t, c0, smem + s * G::BUF);
    cp_async_commit();
  }
  int slot = 0;
  for (; tile < a.ntiles; tile += gridDim.x) {
    unsigned char *buf = smem + slot * G::BUF;
    const long long next = tile + (long long)(ns - 1) * gridDim.x;
    const int nslot = slot == 0 ? ns - 1 : slot - 1;             // the buffer freed by the previous iteration
    if (next < a.ntiles) cic_stage<R, N, M, K, CT>(a, next, c0, smem + nslot * G::BUF);
    cp_async_commit();
    if (ns == 2) cp_async_wait<1>(); else if (ns == 3) cp_async_wait<2>(); else cp_async_wait<3>();
    __syncthreads();
    slot = slot + 1 == ns ? 0 : slot + 1;

    const long long mo0 = (tile * G::ROWS + row) * (long long)K;   // first output of this thread
    if ((size_t)mo0 < a.n_out) {
      uint32_t res[K];
      if (R == 8 && N == 4 && M == 1 && !a.recursive) {
        cic8n4_row<CT, K>(buf + row * G::S, buf + (row + 1) * G::S, ch, res);
      } else {
      uint32_t r[N], d[N][M];
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
  // ring depth: the non-recursive form is light enough that 2 CTAs per SM with two tiles in flight each win
  // (94.6 % of the measured HBM peak against 87 % with 3 CTAs x 1 tile); the recursive form wants the third CTA
  const int ns_default = (R == 8 && N == 4 && M == 1 && !b.recursive) ? 3 : 2;
  { const char *ev = getenv("B2D_CIC_STAGES"); b.nstage = ev ? atoi(ev) : ns_default; }
  if (b.nstage < 2 || b.nstage > 4) b.nstage = ns_default;
  const size_t smem = (size_t)b.nstage * G::BUF;
  cudaError_t e = cudaFuncSetAttribute(cic_dec_fast_kernel<R, N, M, K, CT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  int per_sm = 1;
  cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, cic_dec_fast_kernel<R, N, M, K, CT>, kCicThreads, smem);
  if (per_sm < 1) per_sm = 1;
  long long gx = 148LL * per_sm;
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
  { const char *e = getenv("B2D_CIC_RECURSIVE"); a.recursive = (e && *e == '1') ? 1 : 0; }
  const int ct = (p.interleaved && p.C == 2) ? 2 : 1;
#define X(r, n, m, k)                                                                     \
  if (p.R == r && p.N == n && p.M == m)                                                   \
    return ct == 2 ? launch_dec<r, n, m, k, 2>(a, p.C, st) : launch_dec<r, n, m, k, 1>(a, p.C, st);
  B2D_CIC_DEC_CASES(X)
#undef X
  return cudaErrorNotSupported;
}

}  // namespace b2d
