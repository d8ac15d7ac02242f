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
#include "kernels.h"

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
  if (a.fast && (32 % p.chn) == 0 && !p.force_thread) {
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
