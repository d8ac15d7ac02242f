// intg_dump.cu -- integrate-and-dump over CHN interleaved channels: ac_intg_dump (reference
// include/ac_dsp/ac_intg_dump.h:84-151).
//
// What the reference computes.  Per frame it reads one n_sample token, adds samples j = 1 .. NS of every channel into
// temp[i] (ACC_TYPE, re-quantised at every add, :100) and at j == n_sample writes the CHN sums (converted to OUT_TYPE)
// and clears them.  A token outside 1 .. NS never matches, so such a frame consumes NS samples per channel, writes
// nothing and leaves temp[] to carry into the next frame (:133-147).  The host turns the token sequence into SEGMENTS
// of the per-channel sample axis (that is control flow, not arithmetic); a kernel sums each segment:
//     out[s][c] = OUT_TYPE( wrap_ACC( carry[c]*(s == 0) + sum_{k in segment s} q(x[k*CHN + c]) ) )
// With Q in {AC_TRN, AC_RND} and O = AC_WRAP the per-add re-quantisation is q(x) = floor((x + rnd) / 2^d) for
// d = F_in - F_acc > 0 (an exact shift otherwise) and the sum is modular, hence order-free; other ACC modes are summed
// sequentially in the reference's order by one thread per (segment, channel).  HBM-bound: 2 B read per sample.
//   intgdump_warp_kernel    CHN divides 32: a warp streams one segment with coalesced loads (a lane always sees the same
//                           channel), then the lanes of a channel are folded with xor-shuffles.
//   intgdump_thread_kernel  any CHN / any ACC mode: one thread per (segment, channel).
//   intgdump_vec_kernel     16-bit samples, CHN in {1,2,4,8}, equal segments that are whole 16-byte chunks: 128-bit
//                           loads (a chunk always starts at channel 0), per-lane channel sums with DP2A, and a
//                           transposing butterfly -- log2(CHN) exchange steps that halve the values a lane holds, then
//                           one shuffle per remaining step -- across the lanes of a segment.
#include "kernels.h"
#include <algorithm>

namespace b2d {

struct IdArgs {
  Fmt in, acc, out;
  int chn, in_bytes, out_bytes, fast;
  int d;                       // F_in - F_acc
  long long rnd;
  const void *x;               // samples of this call, interleaved over CHN
  void *y;                     // [nseg_out][CHN] outputs
  const int64_t *carry;        // [CHN] temp[] on entry (ACC raw)
  int64_t *carry_next;         // [CHN] temp[] on exit
  const unsigned long long *table;   // [nseg + 1] segment boundaries (per-channel sample index) or null when regular
  unsigned long long n_reg;    // regular: every segment has n_reg samples per channel
  size_t nseg_out;             // dumping segments
  int has_tail;                // one more segment follows that is not dumped: its sum becomes carry_next
  unsigned long long tail_end; // regular mode: end of the tail segment
};

__device__ __forceinline__ void id_bounds(const IdArgs &a, size_t s, unsigned long long &b, unsigned long long &e) {
  if (a.table) { b = a.table[s]; e = a.table[s + 1]; }
  else { b = s * a.n_reg; e = s < a.nseg_out ? b + a.n_reg : a.tail_end; }
}

__device__ __forceinline__ int64_t id_term(const IdArgs &a, int64_t x) {
  return a.d > 0 ? (x + a.rnd) >> a.d : (int64_t)((uint64_t)x << (-a.d));
}

__device__ __forceinline__ void id_finish(const IdArgs &a, size_t s, int c, int64_t sum) {
  if (s < a.nseg_out) store_raw(a.y, s * a.chn + c, a.out_bytes, convert((i128)sum, a.acc.F(), a.out));
  else a.carry_next[c] = sum;
}

__global__ void __launch_bounds__(256) intgdump_warp_kernel(IdArgs a) {
  const int lane = threadIdx.x & 31;
  const size_t nseg = a.nseg_out + (a.has_tail ? 1 : 0);
  const size_t warp0 = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarps = ((size_t)gridDim.x * blockDim.x) >> 5;
  const int c = lane % a.chn;                                  // CHN divides 32: element e of a segment is channel e % CHN
  for (size_t s = warp0; s < nseg; s += nwarps) {
    unsigned long long b, e;
    id_bounds(a, s, b, e);
    const unsigned long long e0 = b * a.chn, e1 = e * a.chn;   // element range of the segment
    // start the lanes at a multiple of 32 elements so that lane <-> channel stays fixed and loads stay aligned
    int64_t sum = 0;
    unsigned long long p = (e0 & ~31ULL) + lane;
    if (p < e0) p += 32;
    if (a.in_bytes == 2) {
      const int16_t *x = (const int16_t *)a.x;
      const uint16_t *xu = (const uint16_t *)a.x;
      for (; p + 7 * 32 < e1; p += 8 * 32) {                   // 8 loads in flight per lane
        int64_t v[8];
#pragma unroll
        for (int u = 0; u < 8; u++) v[u] = a.in.S ? (int64_t)x[p + 32 * u] : (int64_t)xu[p + 32 * u];
#pragma unroll
        for (int u = 0; u < 8; u++) sum += id_term(a, v[u]);
      }
      for (; p < e1; p += 32) sum += id_term(a, a.in.S ? (int64_t)x[p] : (int64_t)xu[p]);
    } else {
      for (; p < e1; p += 32) sum += id_term(a, load_raw(a.x, p, a.in_bytes, a.in.S));
    }
    for (int off = a.chn; off < 32; off <<= 1) sum += __shfl_xor_sync(0xffffffffu, sum, off);
    if (lane < a.chn) {
      if (s == 0) sum += a.carry[c];
      id_finish(a, s, c, wrap_bits(sum, a.acc.W, a.acc.S));
    }
  }
}

__global__ void __launch_bounds__(256) intgdump_thread_kernel(IdArgs a) {
  const size_t nseg = a.nseg_out + (a.has_tail ? 1 : 0);
  const size_t total = nseg * a.chn;
  for (size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (size_t)gridDim.x * blockDim.x) {
    const size_t s = t / a.chn;
    const int c = (int)(t % a.chn);
    unsigned long long b, e;
    id_bounds(a, s, b, e);
    int64_t acc = s == 0 ? a.carry[c] : 0;
    if (a.fast) {
      for (unsigned long long k = b; k < e; k++) acc += id_term(a, load_raw(a.x, k * a.chn + c, a.in_bytes, a.in.S));
      acc = wrap_bits(acc, a.acc.W, a.acc.S);
    } else {
      for (unsigned long long k = b; k < e; k++) acc = macc(acc, a.acc, (i128)load_raw(a.x, k * a.chn + c, a.in_bytes, a.in.S), a.in.F());
    }
    id_finish(a, s, c, acc);
  }
}

// ---- 128-bit path ------------------------------------------------------------------------------------------------
// Element e of a 16-byte chunk belongs to channel e % CHN (segments start on chunk boundaries and CHN divides 8).
template <int CHN, bool SGN, bool DPOS>
__device__ __forceinline__ void id_chunk_sums(const uint4 &q, int d, int rnd, int (&acc)[CHN]) {
  const unsigned w[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
  for (int k = 0; k < 4; k++) {
    if (DPOS) {                                   // per-term floor((x + rnd) / 2^d), 0 < d <= 15
      const int lo = SGN ? (int)(short)(w[k] & 0xffffu) : (int)(w[k] & 0xffffu);
      const int hi = SGN ? ((int)w[k] >> 16) : (int)(w[k] >> 16);
      acc[(2 * k) % CHN] += (lo + rnd) >> d;
      acc[(2 * k + 1) % CHN] += (hi + rnd) >> d;
    } else if (CHN == 1) {
      acc[0] = SGN ? __dp2a_lo((int)w[k], 0x0101, acc[0]) : (int)__dp2a_lo(w[k], 0x0101u, (unsigned)acc[0]);
    } else {
      const int c0 = (2 * k) % CHN, c1 = (2 * k + 1) % CHN;
      acc[c0] = SGN ? __dp2a_lo((int)w[k], 0x0001, acc[c0]) : (int)__dp2a_lo(w[k], 0x0001u, (unsigned)acc[c0]);
      acc[c1] = SGN ? __dp2a_lo((int)w[k], 0x0100, acc[c1]) : (int)__dp2a_lo(w[k], 0x0100u, (unsigned)acc[c1]);
    }
  }
}

// Sum v[c] over the G (power of two, >= CHN) lanes of a group; on return every lane holds the total of channel `ch`.
template <int CHN, typename T>
__device__ __forceinline__ T id_group_reduce(T (&v)[CHN], int lane, int G, int &ch) {
  ch = 0;
  if (CHN >= 2) {
    const bool up = lane & 1;
#pragma unroll
    for (int i = 0; i < CHN / 2; i++) {
      const T send = up ? v[i] : v[i + CHN / 2], keep = up ? v[i + CHN / 2] : v[i];
      v[i] = keep + __shfl_xor_sync(0xffffffffu, send, 1);
    }
    ch += up ? CHN / 2 : 0;
  }
  if (CHN >= 4) {
    const bool up = lane & 2;
#pragma unroll
    for (int i = 0; i < CHN / 4; i++) {
      const T send = up ? v[i] : v[i + CHN / 4], keep = up ? v[i + CHN / 4] : v[i];
      v[i] = keep + __shfl_xor_sync(0xffffffffu, send, 2);
    }
    ch += up ? CHN / 4 : 0;
  }
  if (CHN >= 8) {
    const bool up = lane & 4;
    const T send = up ? v[0] : v[1], keep = up ? v[1] : v[0];
    v[0] = keep + __shfl_xor_sync(0xffffffffu, send, 4);
    ch += up ? 1 : 0;
  }
  for (int off = CHN; off < G; off <<= 1) v[0] += __shfl_xor_sync(0xffffffffu, v[0], off);
  return v[0];
}

__device__ __forceinline__ void id_vec_finish(const IdArgs &a, size_t s, int c, int64_t sum) {
  if (a.d < 0) sum = (int64_t)((uint64_t)sum << (-a.d));
  if (s == 0) sum += a.carry[c];
  id_finish(a, s, c, wrap_bits(sum, a.acc.W, a.acc.S));
}

// L = 16-byte chunks per segment.  L <= 32 (a power of two): a warp row of 32 chunks holds 32/L segments, ROWS rows in
// flight per warp.  L > 32: a warp walks one segment.
template <int CHN, bool SGN, bool DPOS, bool WIDE>
__global__ void __launch_bounds__(256, 4) intgdump_vec_kernel(IdArgs a, unsigned long long L) {
  constexpr int ROWS = 4;
  const int lane = threadIdx.x & 31;
  const size_t warp0 = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarps = ((size_t)gridDim.x * blockDim.x) >> 5;
  const uint4 *x4 = (const uint4 *)a.x;
  const int rnd = (int)a.rnd;
  if (L <= 32) {
    const unsigned long long nchunks = a.nseg_out * L, nrows = (nchunks + 31) / 32;
    const int G = (int)L, lg = 31 - __clz((int)L);
    for (unsigned long long r0 = warp0 * ROWS; r0 < nrows; r0 += nwarps * ROWS) {
      uint4 q[ROWS];
#pragma unroll
      for (int u = 0; u < ROWS; u++) {
        const unsigned long long ck = (r0 + u) * 32 + lane;
        q[u] = ck < nchunks ? __ldg(x4 + ck) : make_uint4(0, 0, 0, 0);
      }
#pragma unroll
      for (int u = 0; u < ROWS; u++) {
        int acc[CHN];
#pragma unroll
        for (int c = 0; c < CHN; c++) acc[c] = 0;
        id_chunk_sums<CHN, SGN, DPOS>(q[u], a.d, rnd, acc);
        int ch;
        const int tot = id_group_reduce<CHN, int>(acc, lane, G, ch);
        const unsigned long long ck = (r0 + u) * 32 + lane;
        if ((lane & (G - 1)) < CHN && ck < nchunks) id_vec_finish(a, (size_t)(ck >> lg), ch, (int64_t)tot);
      }
    }
  } else {
    for (size_t s = warp0; s < a.nseg_out; s += nwarps) {
      const uint4 *xs = x4 + s * L;
      int ch;
      long long tot;
      if (WIDE) {
        long long wide[CHN];
#pragma unroll
        for (int c = 0; c < CHN; c++) wide[c] = 0;
        unsigned long long k = lane;
        while (k < L) {
          int acc[CHN];
#pragma unroll
          for (int c = 0; c < CHN; c++) acc[c] = 0;
          // at most 8 * 2^16 per chunk: 2048 chunks stay inside int32
          for (int it = 0; it < 512 && k < L; it++, k += ROWS * 32) {
            uint4 q[ROWS];
#pragma unroll
            for (int u = 0; u < ROWS; u++) q[u] = k + 32 * u < L ? __ldg(xs + k + 32 * u) : make_uint4(0, 0, 0, 0);
#pragma unroll
            for (int u = 0; u < ROWS; u++) id_chunk_sums<CHN, SGN, DPOS>(q[u], a.d, rnd, acc);
          }
#pragma unroll
          for (int c = 0; c < CHN; c++) wide[c] += acc[c];
        }
        tot = id_group_reduce<CHN, long long>(wide, lane, 32, ch);
      } else {                                  // the whole segment sums inside int32 (n_reg < 2^15)
        int acc[CHN];
#pragma unroll
        for (int c = 0; c < CHN; c++) acc[c] = 0;
        const unsigned Ls = (unsigned)L;
        for (unsigned k = lane; k < Ls; k += ROWS * 32) {
          uint4 q[ROWS];
#pragma unroll
          for (int u = 0; u < ROWS; u++) q[u] = k + 32 * u < Ls ? __ldg(xs + k + 32 * u) : make_uint4(0, 0, 0, 0);
#pragma unroll
          for (int u = 0; u < ROWS; u++) id_chunk_sums<CHN, SGN, DPOS>(q[u], a.d, rnd, acc);
        }
        tot = id_group_reduce<CHN, int>(acc, lane, 32, ch);
      }
      if (lane < CHN) id_vec_finish(a, s, ch, tot);
    }
  }
}

template <int CHN>
static void launch_vec(const IdArgs &a, unsigned long long L, unsigned blocks, cudaStream_t st) {
  const bool sgn = a.in.S != 0, dpos = a.d > 0, wide = a.n_reg >= 32768;
#define B2D_ID_VEC(S_, D_, W_) intgdump_vec_kernel<CHN, S_, D_, W_><<<blocks, 256, 0, st>>>(a, L)
  if (wide) {
    if (sgn && !dpos) B2D_ID_VEC(true, false, true);
    else if (sgn && dpos) B2D_ID_VEC(true, true, true);
    else if (!sgn && !dpos) B2D_ID_VEC(false, false, true);
    else B2D_ID_VEC(false, true, true);
  } else {
    if (sgn && !dpos) B2D_ID_VEC(true, false, false);
    else if (sgn && dpos) B2D_ID_VEC(true, true, false);
    else if (!sgn && !dpos) B2D_ID_VEC(false, false, false);
    else B2D_ID_VEC(false, true, false);
  }
#undef B2D_ID_VEC
}

// Which kernel a launch takes (b2d_intgdump_path reports it).
const char *intgdump_path(const IdLaunch &p) {
  const bool fast = p.facc.O == B2D_WRAP && (p.facc.Q == B2D_TRN || p.facc.Q == B2D_RND);
  const int d = p.fin.F() - p.facc.F();
  if (!fast || d > 62 || d < -62 || p.force_thread) return "intgdump_thread";
  const int chn = p.chn;
  if (!p.table && !p.has_tail && container_bytes(p.fin.W) == 2 && (chn == 1 || chn == 2 || chn == 4 || chn == 8) && d <= 15 && d >= -40 &&
      ((uintptr_t)p.in & 15) == 0 && (p.n_reg * chn) % 8 == 0) {
    const unsigned long long L = p.n_reg * chn / 8;
    if (L > 32 || ((L & (L - 1)) == 0 && L >= (unsigned long long)chn)) return "intgdump_vec";
  }
  return (32 % chn) == 0 ? "intgdump_warp" : "intgdump_thread";
}

cudaError_t launch_intgdump(const IdLaunch &p, cudaStream_t st) {
  IdArgs a;
  a.in = p.fin; a.acc = p.facc; a.out = p.fout; a.chn = p.chn;
  a.in_bytes = container_bytes(p.fin.W); a.out_bytes = container_bytes(p.fout.W);
  a.fast = (p.facc.O == B2D_WRAP && (p.facc.Q == B2D_TRN || p.facc.Q == B2D_RND)) ? 1 : 0;
  a.d = p.fin.F() - p.facc.F();
  if (a.d > 62 || a.d < -62) a.fast = 0;
  a.rnd = (a.d > 0 && p.facc.Q == B2D_RND) ? (1LL << (a.d - 1)) : 0;
  a.x = p.in; a.y = p.out; a.carry = p.carry; a.carry_next = p.carry_next; a.table = p.table; a.n_reg = p.n_reg;
  a.nseg_out = p.nseg_out; a.has_tail = p.has_tail; a.tail_end = p.tail_end;
  const size_t nseg = p.nseg_out + (p.has_tail ? 1 : 0);
  if (nseg == 0) return cudaSuccess;
  const char *path = intgdump_path(p);
  if (path[9] == 'v') {
    const unsigned long long L = p.n_reg * p.chn / 8;
    const unsigned long long warps = L <= 32 ? (nseg * L + 127) / 128 : nseg;
    unsigned blocks = (unsigned)std::min<unsigned long long>((warps + 7) / 8, 148ULL * 4 * 4);
    if (blocks == 0) blocks = 1;
    switch (p.chn) {
      case 1: launch_vec<1>(a, L, blocks, st); break;
      case 2: launch_vec<2>(a, L, blocks, st); break;
      case 4: launch_vec<4>(a, L, blocks, st); break;
      default: launch_vec<8>(a, L, blocks, st); break;
    }
  } else if (path[9] == 'w') {
    size_t blocks = (nseg * 32 + 255) / 256;
    if (blocks > 148 * 8) blocks = 148 * 8;
    intgdump_warp_kernel<<<(unsigned)blocks, 256, 0, st>>>(a);
  } else {
    size_t blocks = (nseg * p.chn + 255) / 256;
    if (blocks > 148 * 16) blocks = 148 * 16;
    intgdump_thread_kernel<<<(unsigned)blocks, 256, 0, st>>>(a);
  }
  return cudaGetLastError();
}

}  // namespace b2d
