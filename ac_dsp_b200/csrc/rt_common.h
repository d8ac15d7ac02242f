// rt_common.h -- what the host-runtime translation units (runtime.cu, rt_fir.cu, rt_cic.cu, rt_poly.cu,
// rt_intgdump.cu) share: error reporting, descriptor checks, the ONE pipelined host-buffer loop every run() on host
// memory goes through, the communicator, checkpoint blobs and the two handle types that other families reach into.
#pragma once
#include <algorithm>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>
#include <string>
#include <vector>

#include <nvtx3/nvToolsExt.h>

#include "kernels.h"
#include "nccl_dl.h"

namespace b2d {

// ------------------------------------------------------------------------------------------ errors
int fail(int status, const char *fmt, ...);   // sets the thread-local text of b2d_last_error(), returns status
// NVTX range per load / run entry point (tracing): header-only NVTX 3, a no-op unless a tool is attached
struct TraceRange {
  explicit TraceRange(const char *name) { nvtxRangePushA(name); }
  ~TraceRange() { nvtxRangePop(); }
};
#define CU(expr)                                                                                      \
  do {                                                                                                \
    cudaError_t e__ = (expr);                                                                         \
    if (e__ != cudaSuccess) return ::b2d::fail(B2D_ECUDA, "%s: %s (%s:%d)", #expr, cudaGetErrorString(e__), __FILE__, __LINE__); \
  } while (0)

inline Fmt to_fmt(const b2d_fmt &f) { return Fmt{f.W, f.I, f.S ? 1 : 0, f.Q, f.O}; }
int check_fmt(const b2d_fmt &f, int maxW, const char *what);
int use_device(int dev);
// raw COEFF_TYPE values in their container -> int64, wrapped to the format (what every load() starts with)
void widen_coeffs(const void *raw, size_t n, int c_bytes, const Fmt &fc, int64_t *out);

// ------------------------------------------------------------------------------------ host pipeline
// run() on HOST buffers: the stream is cut into chunks; chunk i+1 is copied in while chunk i computes
// and chunk i-1 is copied out (three streams, three device slots).  Each chunk is an ordinary run_dev()
// call, so the result is the reference's own "several run() calls" behaviour by construction.
struct Pipe {
  static const int S = 3;
  cudaStream_t s_in = nullptr, s_k = nullptr, s_out = nullptr;
  cudaEvent_t e_in[S] = {}, e_k[S] = {}, e_out[S] = {};
  void *d_in[S] = {}, *d_out[S] = {}, *d_pk[S] = {};
  size_t cap_in = 0, cap_out = 0, cap_pk = 0;
  bool ready = false;
  int init();
  int ensure(size_t in_bytes, size_t out_bytes, size_t pk_bytes);
  void destroy();
};

// copy `len` samples per channel starting at time `off` between a full buffer (n_full per channel) and a
// compact chunk buffer (len per channel); `bytes` per element
cudaError_t copy_chunk(void *dst, const void *src, bool to_device, int bytes, uint32_t C, int interleaved,
                       size_t n_full, size_t off, size_t len, cudaStream_t st);

// What differs between the handle families in a host-buffer run().
struct HostRun {
  const void *in;
  void *out;
  size_t n;             // inputs per channel of the whole call
  uint32_t C;
  int il;               // input layout interleaved
  int in_bytes;         // input container
  int out_bytes;        // output container (what the kernels write)
  int wire_bytes;       // bytes per output value on the host link and in `out`: out_bytes, or fewer (B2D_WIRE_PACKED)
  bool out_like_in;     // FIR: one output per input, same layout; otherwise PLANAR outputs with a channel stride of no_total
  size_t no_total;      // outputs per channel of the whole call
  size_t L, Lout;       // chunk length in inputs per channel; output capacity per channel of one chunk
};

// inputs per channel per chunk so that one chunk moves about 24 MiB over the link: the first chunk's H2D and the last
// chunk's kernel + D2H are the part of a call nothing overlaps with (about chunk / link rate: 0.45 ms at 55 GB/s, against
// 1.9 ms with the 96 MiB chunks of round 1), while a chunk still takes ~20x longer on the link than its six CUDA calls take
// to issue.  B2D_PIPE_CHUNK_BYTES overrides the target (the parity tests push many small chunks through the three slots).
inline size_t pipe_chunk(size_t n, double link_bytes_per_input) {
  double target = (double)(24u << 20);
  size_t floor_len = 4096;
  if (const char *e = getenv("B2D_PIPE_CHUNK_BYTES")) {
    const double v = atof(e);
    if (v >= 256) { target = v; floor_len = 16; }
  }
  size_t L = std::max<size_t>((size_t)(target / link_bytes_per_input), floor_len);
  // chunk boundaries on multiples of 4096 samples (>= 8 KiB of every container): with page-aligned caller buffers every
  // DMA piece starts on a page.  Measured on a B200 box (r02, tools/trace_e2e.py, fir256 2^27 IQ samples per call):
  // 3.22-3.29 G IQ samples/s aligned vs 2.81-3.05 unaligned (containers), 5.04-5.22 vs 4.51-4.67 (packed).
  if (L > 16384) L &= ~(size_t)4095;
  return std::min(L, n);
}

// count(len): outputs per channel the next `len` inputs will produce (state-dependent: asked just before the launch);
// launch(d_in, len, d_out, n_out, stream): the family's run_dev body.
template <class Count, class Launch>
int run_host_pipeline(Pipe &P, const HostRun &r, Count count, Launch launch) {
  int st = P.init();
  if (st) return st;
  const bool packed = r.wire_bytes < r.out_bytes;
  if ((st = P.ensure(r.L * r.C * r.in_bytes, r.Lout * r.C * r.out_bytes, packed ? r.Lout * r.C * r.wire_bytes + 16 : 0))) return st;
  if (r.n <= r.L) {
    // one chunk: there is nothing to overlap -- copy in, compute, copy out on the compute stream, one synchronisation.
    // This is the path of the reference-signature calls that move a handful of samples (ac_fir_prog_coeffs::run: one).
    const size_t no = count(r.n);
    CU(copy_chunk(P.d_in[0], r.in, true, r.in_bytes, r.C, r.il, r.n, 0, r.n, P.s_k));
    if ((st = launch(P.d_in[0], r.n, P.d_out[0], no, P.s_k))) return st;
    const void *src = P.d_out[0];
    if (packed && no) {
      CU(launch_pack_wire(P.d_out[0], r.out_bytes, P.d_pk[0], r.wire_bytes, no * r.C, P.s_k));
      src = P.d_pk[0];
    }
    if (no) {
      if (r.out_like_in) CU(copy_chunk(r.out, src, false, r.wire_bytes, r.C, r.il, r.n, 0, r.n, P.s_k));
      else CU(copy_chunk(r.out, src, false, r.wire_bytes, r.C, 0, r.no_total, 0, no, P.s_k));
    }
    CU(cudaStreamSynchronize(P.s_k));
    return B2D_OK;
  }
  // B2D_PIPE_TRACE=1 (diagnosis): timestamps around every copy and launch, printed per call on stderr
  const char *trace_env = getenv("B2D_PIPE_TRACE");
  const bool trace = trace_env && *trace_env == '1';
  std::vector<cudaEvent_t> tev;
  auto stamp = [&](cudaStream_t stq) { if (trace) { cudaEvent_t e; cudaEventCreate(&e); cudaEventRecord(e, stq); tev.push_back(e); } };
  size_t i = 0, off_out = 0;
  for (size_t off = 0; off < r.n; off += r.L, i++) {
    const int s = (int)(i % Pipe::S);
    const size_t len = std::min(r.L, r.n - off);
    const size_t no = count(len);
    if (i >= (size_t)Pipe::S) CU(cudaStreamWaitEvent(P.s_in, P.e_k[s], 0));       // slot's previous kernel has read its input
    stamp(P.s_in);
    CU(copy_chunk(P.d_in[s], r.in, true, r.in_bytes, r.C, r.il, r.n, off, len, P.s_in));
    stamp(P.s_in);
    CU(cudaEventRecord(P.e_in[s], P.s_in));
    CU(cudaStreamWaitEvent(P.s_k, P.e_in[s], 0));
    if (i >= (size_t)Pipe::S) CU(cudaStreamWaitEvent(P.s_k, P.e_out[s], 0));      // slot's previous output has left
    stamp(P.s_k);
    if ((st = launch(P.d_in[s], len, P.d_out[s], no, P.s_k))) return st;
    const void *src = P.d_out[s];
    if (packed && no) {
      CU(launch_pack_wire(P.d_out[s], r.out_bytes, P.d_pk[s], r.wire_bytes, no * r.C, P.s_k));
      src = P.d_pk[s];
    }
    stamp(P.s_k);
    CU(cudaEventRecord(P.e_k[s], P.s_k));
    CU(cudaStreamWaitEvent(P.s_out, P.e_k[s], 0));
    stamp(P.s_out);
    if (no) {
      if (r.out_like_in) CU(copy_chunk(r.out, src, false, r.wire_bytes, r.C, r.il, r.n, off, len, P.s_out));
      else CU(copy_chunk(r.out, src, false, r.wire_bytes, r.C, 0, r.no_total, off_out, no, P.s_out));
    }
    stamp(P.s_out);
    CU(cudaEventRecord(P.e_out[s], P.s_out));
    off_out += no;
  }
  CU(cudaStreamSynchronize(P.s_out));
  CU(cudaStreamSynchronize(P.s_k));
  if (trace) {   // per chunk: [h2d start, h2d end, kernel start, kernel end, d2h start, d2h end] in ms from the first stamp
    for (size_t c = 0; c * 6 + 5 < tev.size(); c++) {
      float t[6];
      for (int k = 0; k < 6; k++) cudaEventElapsedTime(&t[k], tev[0], tev[c * 6 + k]);
      fprintf(stderr, "pipe chunk %3zu  h2d %8.3f-%8.3f  kernel %8.3f-%8.3f  d2h %8.3f-%8.3f\n", c, t[0], t[1], t[2], t[3], t[4], t[5]);
    }
    for (cudaEvent_t e : tev) cudaEventDestroy(e);
  }
  return B2D_OK;
}

// bytes per output value in host memory for a handle's wire format
inline int wire_bytes_of(int W, int wire) { return wire == B2D_WIRE_PACKED ? (W + 7) / 8 : container_bytes(W); }
int check_wire(int32_t wire);

}  // namespace b2d

// ----------------------------------------------------------------------------------------------- comm
struct b2d_comm {
  b2d::ncclComm_t comm = nullptr;
  int rank = 0, world = 1, device = 0;
  cudaStream_t stream = nullptr;
  void *d_buf = nullptr;
  size_t cap = 0;
};

namespace b2d {
// values[0..n) of rank `root` -> every rank (int64 payload), synchronous.
int comm_bcast_i64(b2d_comm *c, int64_t *values, size_t n, int root);

// ---------------------------------------------------------------------------------------- checkpoints
struct StateHdr { uint32_t magic, version; uint64_t n_seen; uint32_t hist, channels, bytes, pad; };
static const uint32_t kFirMagic = 0x46324442u, kCicMagic = 0x43324442u;
static const uint32_t kDecMagic = 0x44324442u, kIntrMagic = 0x49324442u, kDumpMagic = 0x55324442u, kCasMagic = 0x4b324442u;
static const uint32_t kMvAvgMagic = 0x4d324442u;
// StateHdr followed by device arrays copied verbatim.
struct StatePart { void *dev; size_t bytes; };
size_t state_total(const StatePart *parts, int np);
int state_get(const StateHdr &hd, const StatePart *parts, int np, void *blob, size_t bytes);
int state_set(const StateHdr &want, const StatePart *parts, int np, const void *blob, size_t bytes, StateHdr *got);

enum { PATH_GENERIC = 0, PATH_Q15 = 1, PATH_WIDE = 2, PATH_Q24 = 3 };
}  // namespace b2d

// ------------------------------------------------------------------------------------------- handles
// (the cascade owns one FIR and one CIC object and reads their members)
struct b2d_fir {
  b2d_fir_desc d;
  b2d::Fmt fin, fc, fa, fo;
  int device = 0, T = 0, in_bytes = 2, out_bytes = 2, c_bytes = 2;
  int path = b2d::PATH_GENERIC;
  int wire = B2D_WIRE_CONTAINER;
  std::vector<int64_t> h_coeff;   // [C][N] raw, wrapped to COEFF_TYPE
  std::vector<char> ch_loaded;    // per channel
  int64_t *d_coeff64 = nullptr;
  uint32_t *d_coeff_pk = nullptr;
  int pk_words = 0;
  int32_t *d_coeff32 = nullptr;
  int wide_words = 0, wide_mode = 0;
  // overlap-save evaluation of long q15 filters (fir_ovs.cu): twiddle tables, per-channel spectra (prepared before the first
  // long call after a load), the a-priori error bound of the loaded taps, the optional residual monitor
  int ovs_mode = 0;               // 0: off, 1: long calls, 2: every call (B2D_FIR_OVS=2, tests)
  double2 *d_tw = nullptr;        // [6][256] + [6][16]
  double2 *d_hs = nullptr;        // [C][4096]
  std::vector<char> hs_stale;     // per channel
  double ovs_bound = 0.0;
  double *d_resid = nullptr;
  void *d_tail[2] = {nullptr, nullptr};
  int cur = 0;
  b2d_comm *comm = nullptr;
  int root = 0;
  int64_t *d_dl = nullptr;        // REG_SHARE: OUT_TYPE(reg[N_TAPS-1]) per channel
  bool ran = false;               // samples were filtered since create / reset (see the TRANSPOSED rule in b2d_fir_load)
  int64_t *d_pend[2] = {nullptr, nullptr};   // TRANSPOSED: [C][N_TAPS-1] partial sums the previous taps left for the next outputs
  int pcur = 0;
  size_t pend_rem = 0;            // outputs that still start from a pending partial sum (0: d_pend is all zeros / unallocated)
  void *d_win = nullptr;          // run_window scratch: [C][N_TAPS-1] tail, [C] newest samples, [C] outputs, [C][N_TAPS-1] dummy tail
  cudaEvent_t e_hist = nullptr;   // recorded after the history carry of the last launch: the next launch's stream waits on it
  b2d::Pipe pipe;
};

struct b2d_cic {
  b2d_cic_desc d;
  b2d::Fmt fin, fo;
  int device = 0, intW = 0, H = 0, in_bytes = 2, out_bytes = 4;
  int fast = 0;
  int wire = B2D_WIRE_CONTAINER;
  unsigned long long n_seen = 0;
  void *d_tail[2] = {nullptr, nullptr};
  int cur = 0;
  cudaEvent_t e_hist = nullptr;
  b2d::Pipe pipe;
};

namespace b2d {
inline bool all_loaded(const b2d_fir *h) {
  for (char c : h->ch_loaded) if (!c) return false;
  return true;
}
// Order the history ping-pong across streams: a launch reads d_tail[cur], written by the previous launch's
// tail kernel on whatever stream that call used.  hist_wait() makes `st` wait for that write, hist_mark() records the new one.
int hist_wait(cudaEvent_t &ev, cudaStream_t st);
int hist_mark(cudaEvent_t &ev, cudaStream_t st);
}  // namespace b2d
