// rt_cic.cu -- host runtime of the CIC handles (ac_cic_dec_full / ac_cic_intr_full) and of the CIC-interpolator + FIR cascade.
#include "rt_common.h"

using namespace b2d;

// ------------------------------------------------------------------------------------------------ CIC
static int log2_ceil_u128(unsigned __int128 v) {
  int k = 0;
  while ((((unsigned __int128)1) << k) < v) k++;
  return k;
}

// find_inter_type_cic_dec (ac_cic_dec_full.h:116-137): outW = log2_ceil(R^N * M^N) + W + !S
// find_inter_type_cic_intr (ac_cic_intr_full.h:107-127): outW = log2_ceil(R^(N-1) * M^N) + W + !S
static int cic_int_width(const b2d_cic_desc *d, int *outW) {
  unsigned __int128 g = 1;
  const unsigned __int128 lim = ((unsigned __int128)1) << 100;
  const uint32_t nr = d->mode == B2D_CIC_INTR ? d->N - 1 : d->N;
  for (uint32_t i = 0; i < nr; i++) { g *= d->R; if (g > lim) return -1; }
  for (uint32_t i = 0; i < d->N; i++) { g *= d->M; if (g > lim) return -1; }
  *outW = log2_ceil_u128(g) + d->in.W + (d->in.S ? 0 : 1);
  return 0;
}

// Differential delay the reference's comb really has.  diffStage() shifts comb_dly_ln[k][0..M-1] with an ASCENDING copy
// loop (ac_cic_full_core.h:247-251: `if (i != 0) dly[i] = dly[i-1]` for i = 0 .. M-1), so dly[0] smears through the line
// and dly[M-1], read at the next step, is the input of two steps ago: the delay is min(M, 2) while the lossless width
// (cic_int_width above) keeps growing with M.  Bit-exactness means following the code, not the intent; M <= 2 -- every
// reference vector and BASELINE configuration -- is unaffected (pinned by tests/golden/cic_comb_quirk.npz).
static uint32_t cic_comb_delay(uint32_t M) { return M > 2 ? 2u : M; }

static int cic_check(const b2d_cic_desc *d, int *outW) {
  int st;
  if (!d) return fail(B2D_EINVAL, "null descriptor");
  if ((st = check_fmt(d->in, 32, "IN_TYPE"))) return st;
  if ((st = check_fmt(d->out, 64, "OUT_TYPE"))) return st;
  if (d->mode != B2D_CIC_DEC && d->mode != B2D_CIC_INTR) return fail(B2D_EINVAL, "bad mode");
  // rate counters of the reference are 8 bits wide (ac_cic_full_core.h:72-73,91); R = 1 never re-reads in the interpolator
  if (d->R < 1 || d->R > 256) return fail(B2D_EINVAL, "R = %u outside 1..256", d->R);
  if (d->R == 1) return fail(B2D_EUNSUPPORTED, "R = 1 (a pass-through rate; the reference's interpolator never re-reads its input then)");
  if (d->M < 1 || d->N < 1 || d->N > 255) return fail(B2D_EINVAL, "M = %u, N = %u invalid", d->M, d->N);
  if (d->n_channels < 1) return fail(B2D_EINVAL, "n_channels must be >= 1");
  if (d->layout != B2D_PLANAR && d->layout != B2D_INTERLEAVED) return fail(B2D_EINVAL, "bad layout");
  if (cic_int_width(d, outW) || *outW > 64) return fail(B2D_EUNSUPPORTED, "lossless internal width exceeds 64 bits");
  if (d->N > 16 || (uint64_t)d->N * d->M > 64) return fail(B2D_EUNSUPPORTED, "N > 16 or N*M > 64");
  return B2D_OK;
}

extern "C" int b2d_cic_int_width(const b2d_cic_desc *desc, int32_t *outW) {
  if (!outW) return fail(B2D_EINVAL, "null argument");
  int w = 0;
  int st = cic_check(desc, &w);
  if (st) return st;
  *outW = w;
  return B2D_OK;
}

static unsigned long long cic_emitted(const b2d_cic *h, unsigned long long K) {
  const long long R = h->d.R, N = h->d.N;
  if (h->d.mode == B2D_CIC_DEC) return (K + R - 1) / R;  // inputs 0, R, 2R, ... are forwarded
  if (K == 0) return 0;
  const long long e = ((long long)K - 1) * R + 1 - (N - 1);  // integrator steps so far minus the N-1 dropped
  return e > 0 ? (unsigned long long)e : 0;
}

extern "C" int b2d_cic_create(b2d_cic **out, const b2d_cic_desc *desc) {
  if (!out) return fail(B2D_EINVAL, "null argument");
  *out = nullptr;
  int w = 0;
  int st = cic_check(desc, &w);
  if (st) return st;
  int dev = desc->device;
  if (dev < 0) CU(cudaGetDevice(&dev));
  if ((st = use_device(dev))) return st;
  b2d_cic *h = new (std::nothrow) b2d_cic();
  if (!h) return fail(B2D_ENOMEM, "handle");
  h->d = *desc; h->fin = to_fmt(desc->in); h->fo = to_fmt(desc->out); h->device = dev; h->intW = w;
  h->in_bytes = container_bytes(desc->in.W); h->out_bytes = container_bytes(desc->out.W);
  h->H = cic_history_len(desc->mode == B2D_CIC_INTR, desc->R, cic_comb_delay(desc->M), desc->N);
  const size_t tail_bytes = (size_t)h->H * desc->n_channels * h->in_bytes;
  for (int i = 0; i < 2; i++) {
    cudaError_t e = cudaMalloc(&h->d_tail[i], tail_bytes);
    if (e == cudaSuccess) e = cudaMemset(h->d_tail[i], 0, tail_bytes);
    if (e != cudaSuccess) {
      cudaGetLastError();
      b2d_cic_destroy(h);
      return fail(B2D_ECUDA, "b2d_cic_create: %s", cudaGetErrorString(e));
    }
  }
  CicLaunch p{};
  p.fin = h->fin; p.fout = h->fo; p.intW = w; p.R = desc->R; p.M = cic_comb_delay(desc->M); p.N = desc->N;
  p.intr = desc->mode == B2D_CIC_INTR; p.C = desc->n_channels; p.interleaved = desc->layout == B2D_INTERLEAVED;
  h->fast = cic_fast_supported(p) ? 1 : (cic_intr_fast_supported(p) ? 2 : 0);
  const char *force = getenv("B2D_FORCE_GENERIC");
  if ((force && *force == '1') || desc->M > 2) h->fast = 0;   // M > 2: the width is not the one the fast kernels were instantiated for
  *out = h;
  return B2D_OK;
}

extern "C" int b2d_cic_destroy(b2d_cic *h) {
  if (!h) return B2D_OK;
  use_device(h->device);
  cudaDeviceSynchronize();
  h->pipe.destroy();
  if (h->e_hist) cudaEventDestroy(h->e_hist);
  for (int i = 0; i < 2; i++) if (h->d_tail[i]) cudaFree(h->d_tail[i]);
  delete h;
  return B2D_OK;
}

extern "C" const char *b2d_cic_path(b2d_cic *h) { return !h ? "" : (h->fast == 1 ? "cic_fast" : (h->fast == 2 ? "cic_intr_fast" : "cic_generic")); }

extern "C" size_t b2d_cic_max_out(b2d_cic *h, size_t n) {
  if (!h) return 0;
  return h->d.mode == B2D_CIC_DEC ? n / h->d.R + 1 : n * h->d.R;
}

static int cic_launch(b2d_cic *h, const void *d_in, size_t n, void *d_out, size_t n_out, cudaStream_t st) {
  if (n == 0) return B2D_OK;
  CicLaunch p;
  p.fin = h->fin; p.fout = h->fo; p.intW = h->intW; p.R = h->d.R; p.M = cic_comb_delay(h->d.M); p.N = h->d.N;
  p.intr = h->d.mode == B2D_CIC_INTR; p.C = h->d.n_channels; p.interleaved = h->d.layout == B2D_INTERLEAVED;
  p.in = d_in; p.out = d_out; p.n = n; p.n_out = n_out;
  p.n_seen = h->n_seen; p.out_first = cic_emitted(h, h->n_seen);
  p.tail = h->d_tail[h->cur]; p.tail_next = h->d_tail[h->cur ^ 1]; p.H = h->H;
  int hs = hist_wait(h->e_hist, st);
  if (hs) return hs;
  CU(h->fast == 1 ? launch_cic_fast(p, st) : (h->fast == 2 ? launch_cic_intr_fast(p, st) : launch_cic_generic(p, st)));
  CU(launch_cic_tail(p, st));
  if ((hs = hist_mark(h->e_hist, st))) return hs;
  h->cur ^= 1;
  h->n_seen += n;
  return B2D_OK;
}

extern "C" int b2d_cic_run_dev(b2d_cic *h, const void *d_in, size_t n, void *d_out, size_t *n_out, void *cuda_stream) {
  TraceRange trace__("b2d_cic_run_dev");
  if (!h || (n && !d_in)) return fail(B2D_EINVAL, "null argument");
  const size_t no = (size_t)(cic_emitted(h, h->n_seen + n) - cic_emitted(h, h->n_seen));
  if (no && !d_out) return fail(B2D_EINVAL, "null output");
  int st = use_device(h->device);
  if (st) return st;
  if ((st = cic_launch(h, d_in, n, d_out, no, (cudaStream_t)cuda_stream))) return st;
  if (n_out) *n_out = no;
  return B2D_OK;
}

extern "C" int b2d_cic_run(b2d_cic *h, const void *in, size_t n, void *out, size_t *n_out) {
  TraceRange trace__("b2d_cic_run");
  if (!h || (n && !in)) return fail(B2D_EINVAL, "null argument");
  const size_t no_total = (size_t)(cic_emitted(h, h->n_seen + n) - cic_emitted(h, h->n_seen));
  if (no_total && !out) return fail(B2D_EINVAL, "null output");
  if (n_out) *n_out = no_total;
  if (n == 0) return B2D_OK;
  int st = use_device(h->device);
  if (st) return st;
  const bool dec = h->d.mode == B2D_CIC_DEC;
  HostRun r;
  r.in = in; r.out = out; r.n = n; r.C = h->d.n_channels; r.il = h->d.layout == B2D_INTERLEAVED;
  r.in_bytes = h->in_bytes; r.out_bytes = h->out_bytes; r.wire_bytes = wire_bytes_of(h->fo.W, h->wire);
  r.out_like_in = false; r.no_total = no_total;     // outputs are PLANAR with a channel stride of the whole call's output count
  r.L = pipe_chunk(n, r.C * (r.in_bytes + r.wire_bytes * (dec ? 1.0 / h->d.R : (double)h->d.R)));
  r.Lout = dec ? r.L / h->d.R + 1 : r.L * h->d.R;
  return run_host_pipeline(h->pipe, r,
                           [h](size_t len) { return (size_t)(cic_emitted(h, h->n_seen + len) - cic_emitted(h, h->n_seen)); },
                           [h](const void *d_in, size_t len, void *d_out, size_t no, cudaStream_t s) { return cic_launch(h, d_in, len, d_out, no, s); });
}

extern "C" int b2d_cic_set_wire(b2d_cic *h, int32_t wire) {
  if (!h) return fail(B2D_EINVAL, "null handle");
  int st = check_wire(wire);
  if (st) return st;
  h->wire = wire;
  return B2D_OK;
}

extern "C" int b2d_cic_reset(b2d_cic *h) {
  if (!h) return fail(B2D_EINVAL, "null handle");
  int st = use_device(h->device);
  if (st) return st;
  CU(cudaDeviceSynchronize());
  for (int i = 0; i < 2; i++) CU(cudaMemset(h->d_tail[i], 0, (size_t)h->H * h->d.n_channels * h->in_bytes));
  h->n_seen = 0;
  return B2D_OK;
}

extern "C" int b2d_cic_state_bytes(b2d_cic *h, size_t *bytes) {
  if (!h || !bytes) return fail(B2D_EINVAL, "null argument");
  *bytes = sizeof(StateHdr) + (size_t)h->H * h->d.n_channels * h->in_bytes;
  return B2D_OK;
}
extern "C" int b2d_cic_get_state(b2d_cic *h, void *blob, size_t bytes) {
  size_t need = 0;
  if (!h || !blob) return fail(B2D_EINVAL, "null argument");
  b2d_cic_state_bytes(h, &need);
  if (bytes < need) return fail(B2D_EINVAL, "state blob needs %zu bytes", need);
  int st = use_device(h->device);
  if (st) return st;
  CU(cudaDeviceSynchronize());
  StateHdr hd{kCicMagic, 1, h->n_seen, (uint32_t)h->H, h->d.n_channels, (uint32_t)h->in_bytes, 0};
  memcpy(blob, &hd, sizeof(hd));
  CU(cudaMemcpy((char *)blob + sizeof(hd), h->d_tail[h->cur], need - sizeof(hd), cudaMemcpyDeviceToHost));
  return B2D_OK;
}
extern "C" int b2d_cic_set_state(b2d_cic *h, const void *blob, size_t bytes) {
  size_t need = 0;
  if (!h || !blob) return fail(B2D_EINVAL, "null argument");
  b2d_cic_state_bytes(h, &need);
  StateHdr hd;
  if (bytes < need) return fail(B2D_EINVAL, "state blob needs %zu bytes", need);
  memcpy(&hd, blob, sizeof(hd));
  if (hd.magic != kCicMagic || hd.hist != (uint32_t)h->H || hd.channels != h->d.n_channels || hd.bytes != (uint32_t)h->in_bytes)
    return fail(B2D_EINVAL, "state blob does not belong to this filter configuration");
  int st = use_device(h->device);
  if (st) return st;
  CU(cudaDeviceSynchronize());
  CU(cudaMemcpy(h->d_tail[h->cur], (const char *)blob + sizeof(hd), need - sizeof(hd), cudaMemcpyHostToDevice));
  h->n_seen = hd.n_seen;
  return B2D_OK;
}

// -------------------------------------------------------------------------------------------- cascade
// ac_cic_intr_full -> ac_fir_* as one handle: fused polyphase kernel when exact, else the two kernels back to back.
struct b2d_cicfir {
  b2d_cic_desc cd;
  b2d_fir_desc fd;
  int device = 0, fused = 0;
  // fused
  Fmt fa, fo;
  int R = 0, taps_total = 0, planes = 3, words = 0, H = 0, lsh = 0, in_bytes = 2, out_bytes = 8;
  std::vector<int64_t> hcic;
  std::vector<char> ch_loaded;
  uint32_t *d_cw = nullptr;
  void *d_tail[2] = {nullptr, nullptr};
  int cur = 0;
  unsigned long long n_seen = 0;
  // two-stage
  b2d_cic *cic = nullptr;
  b2d_fir *fir = nullptr;
  void *d_mid = nullptr;
  size_t mid_cap = 0;
  int wire = B2D_WIRE_CONTAINER;
  cudaEvent_t e_hist = nullptr;
  Pipe pipe;
};

static unsigned long long intr_emitted(unsigned long long K, long long R, long long N) {
  if (K == 0) return 0;
  const long long e = ((long long)K - 1) * R + 1 - (N - 1);
  return e > 0 ? (unsigned long long)e : 0;
}

// Can the pair be evaluated as one exact integer FIR on the 16-bit input?  (see upfir_q15.cu)
static bool cicfir_fusable(const b2d_cic_desc &cd, const b2d_fir_desc &fd, int intW, int *lsh, int *taps_total, int *planes) {
  const Fmt in = to_fmt(cd.in), mid = to_fmt(cd.out), fc = to_fmt(fd.coeff), fa = to_fmt(fd.acc);
  if (in.W > 16 || (!in.S && in.W == 16)) return false;
  if (!(mid.S && mid.F() == in.F() && mid.W >= intW)) return false;          // the lossless INT_TYPE passes unchanged
  if (fa.O != B2D_WRAP || (fa.Q != B2D_TRN && fa.Q != B2D_RND)) return false;
  if (cd.M > 2) return false;                 // cic_comb_delay: the composite taps below assume delay M; two-stage path instead
  const int s = mid.F() + fc.F() - fa.F();
  if (s > 0 || -s > 40 || -s >= fa.W) return false;
  switch (fd.ftype) {
    case B2D_SHIFT_REG: case B2D_ROTATE_SHIFT: case B2D_C_BUFF: case B2D_TRANSPOSED: case B2D_FOLD_EVEN: break;
    case B2D_FOLD_ODD:
      if (!fa.S || !(fa.F() >= mid.F() && mid.W + 1 + (fa.F() - mid.F()) <= fa.W)) return false;   // mid is signed: an unsigned fold wraps
      break;
    default: return false;
  }
  // composite tap magnitude: |c| <= 2^(Wc-1) * (R*M)^N
  unsigned __int128 g = 1;
  for (uint32_t i = 0; i < cd.N; i++) { g *= (unsigned __int128)cd.R * cd.M; if (g > ((unsigned __int128)1 << 40)) return false; }
  const int bits = fc.W + (fc.S ? 0 : 1) + log2_ceil_u128(g);
  const int total = (int)fd.n_taps + (int)cd.N * ((int)cd.R * (int)cd.M - 1);
  if (!upfir_q15_geometry((int)cd.R, total, bits)) return false;
  *lsh = -s; *taps_total = total; *planes = upfir_q15_planes(bits);
  return true;
}

extern "C" int b2d_cicfir_destroy(b2d_cicfir *h) {
  if (!h) return B2D_OK;
  use_device(h->device);
  cudaDeviceSynchronize();
  h->pipe.destroy();
  if (h->cic) b2d_cic_destroy(h->cic);
  if (h->fir) b2d_fir_destroy(h->fir);
  if (h->d_mid) cudaFree(h->d_mid);
  if (h->d_cw) cudaFree(h->d_cw);
  if (h->e_hist) cudaEventDestroy(h->e_hist);
  for (int i = 0; i < 2; i++) if (h->d_tail[i]) cudaFree(h->d_tail[i]);
  delete h;
  return B2D_OK;
}

extern "C" int b2d_cicfir_create(b2d_cicfir **out, const b2d_cic_desc *cd, const b2d_fir_desc *fd) {
  if (!out || !cd || !fd) return fail(B2D_EINVAL, "null argument");
  *out = nullptr;
  int intW = 0;
  int st = cic_check(cd, &intW);
  if (st) return st;
  if (cd->mode != B2D_CIC_INTR) return fail(B2D_EINVAL, "the cascade takes an interpolator (B2D_CIC_INTR) first stage");
  if (fd->in.W != cd->out.W || fd->in.I != cd->out.I || (fd->in.S != 0) != (cd->out.S != 0))
    return fail(B2D_EINVAL, "fir->in must be the interpolator's OUT_TYPE");
  if (fd->n_channels != cd->n_channels) return fail(B2D_EINVAL, "both stages must have the same n_channels");
  int dev = cd->device;
  if (dev < 0) CU(cudaGetDevice(&dev));
  if ((st = use_device(dev))) return st;
  b2d_cicfir *h = new (std::nothrow) b2d_cicfir();
  if (!h) return fail(B2D_ENOMEM, "handle");
  h->cd = *cd; h->fd = *fd; h->device = dev;
  h->cd.device = dev; h->fd.device = dev; h->fd.layout = B2D_PLANAR;
  h->fa = to_fmt(fd->acc); h->fo = to_fmt(fd->out);
  h->out_bytes = container_bytes(fd->out.W);
  h->R = (int)cd->R;
  const uint32_t C = cd->n_channels;
  h->ch_loaded.assign(C, 0);
  const char *force = getenv("B2D_CICFIR_TWO_STAGE");
  h->fused = cicfir_fusable(*cd, *fd, intW, &h->lsh, &h->taps_total, &h->planes) && !(force && *force == '1');
  // a TRANSPOSED second stage whose taps can change keeps partial sums across the change (b2d_fir_load): that state lives
  // in the FIR object, so only the constant-coefficient class may be folded into the composite-tap kernel
  if (fd->ftype == B2D_TRANSPOSED && fd->kind != B2D_FIR_CONST) h->fused = 0;
  // the FIR descriptor is validated by creating the second-stage object in either mode (it also serves reset/load checks)
  if ((st = b2d_fir_create(&h->fir, &h->fd))) { b2d_cicfir_destroy(h); return st; }
  if (h->fused) {
    // boxcar(R*M)^N, ac_cic_intr_full.h:150-215 as one FIR (cic_intr_fast.cu)
    h->hcic.assign(1, 1);
    for (uint32_t s = 0; s < cd->N; s++) {
      std::vector<int64_t> nx(h->hcic.size() + cd->R * cd->M - 1, 0);
      for (size_t i = 0; i < h->hcic.size(); i++)
        for (uint32_t j = 0; j < cd->R * cd->M; j++) nx[i + j] += h->hcic[i];
      h->hcic.swap(nx);
    }
    h->words = upfir_q15_words(h->R, h->taps_total, h->planes);
    h->H = (h->taps_total + h->R - 1) / h->R + 2;
    cudaError_t e = cudaMalloc(&h->d_cw, (size_t)C * h->words * sizeof(uint32_t));
    for (int i = 0; i < 2 && e == cudaSuccess; i++) {
      e = cudaMalloc(&h->d_tail[i], (size_t)h->H * C * 2);
      if (e == cudaSuccess) e = cudaMemset(h->d_tail[i], 0, (size_t)h->H * C * 2);
    }
    if (e != cudaSuccess) { cudaGetLastError(); b2d_cicfir_destroy(h); return fail(B2D_ECUDA, "b2d_cicfir_create: %s", cudaGetErrorString(e)); }
  } else {
    if ((st = b2d_cic_create(&h->cic, &h->cd))) { b2d_cicfir_destroy(h); return st; }
  }
  *out = h;
  return B2D_OK;
}

extern "C" const char *b2d_cicfir_path(b2d_cicfir *h) { return !h ? "" : (h->fused ? "cicfir_fused" : "cicfir_two_stage"); }
extern "C" size_t b2d_cicfir_max_out(b2d_cicfir *h, size_t n) { return h ? n * h->cd.R : 0; }

extern "C" int b2d_cicfir_load(b2d_cicfir *h, const void *coeff_raw, size_t n, int32_t channel) {
  TraceRange trace__("b2d_cicfir_load");
  if (!h) return fail(B2D_EINVAL, "null handle");
  int st = b2d_fir_load(h->fir, coeff_raw, n, channel);     // validation, wrapping to COEFF_TYPE, const-kind rule
  if (st || !h->fused) return st;
  const size_t N = h->fd.n_taps;
  const uint32_t C = h->cd.n_channels;
  for (uint32_t c = 0; c < C; c++) {
    if (channel >= 0 && (uint32_t)channel != c) continue;
    const int64_t *g = h->fir->h_coeff.data() + c * N;
    std::vector<int64_t> eff(g, g + N);
    if (h->fd.ftype == B2D_FOLD_EVEN) {
      std::fill(eff.begin(), eff.end(), 0);
      for (size_t i = 0; i < N / 2; i++) { eff[i] = g[i]; eff[N - 1 - i] = g[i]; }
    } else if (h->fd.ftype == B2D_FOLD_ODD) {
      std::fill(eff.begin(), eff.end(), 0);
      for (size_t i = 0; i < (N - 1) / 2 + 1; i++) { eff[i] = g[i]; if (i != (N - 1) / 2) eff[N - 1 - i] = g[i]; }
    }
    std::vector<int64_t> comp((size_t)h->taps_total, 0);
    for (size_t i = 0; i < h->hcic.size(); i++)
      for (size_t j = 0; j < N; j++) comp[i + j] += h->hcic[i] * eff[j];
    std::vector<uint32_t> pk((size_t)h->words, 0);
    upfir_q15_pack(comp.data(), h->taps_total, h->R, h->planes, pk.data());
    CU(cudaDeviceSynchronize());
    CU(cudaMemcpy(h->d_cw + (size_t)c * h->words, pk.data(), (size_t)h->words * sizeof(uint32_t), cudaMemcpyHostToDevice));
    h->ch_loaded[c] = 1;
  }
  return B2D_OK;
}

static size_t cicfir_count(const b2d_cicfir *h, size_t n) {
  const unsigned long long seen = h->fused ? h->n_seen : h->cic->n_seen;
  return (size_t)(intr_emitted(seen + n, h->cd.R, h->cd.N) - intr_emitted(seen, h->cd.R, h->cd.N));
}

static int cicfir_launch(b2d_cicfir *h, const void *d_in, size_t n, void *d_out, size_t n_out, cudaStream_t st) {
  if (n == 0) return B2D_OK;
  if (h->fused) {
    UpLaunch p;
    p.facc = h->fa; p.fout = h->fo; p.R = h->R; p.taps_total = h->taps_total; p.planes = h->planes; p.lsh = h->lsh;
    p.C = h->cd.n_channels; p.interleaved = h->cd.layout == B2D_INTERLEAVED;
    p.in = d_in; p.out = d_out; p.n = n; p.n_out = n_out;
    p.n_seen = h->n_seen; p.out_first = intr_emitted(h->n_seen, h->cd.R, h->cd.N);
    p.tail = h->d_tail[h->cur]; p.H = h->H; p.cw = h->d_cw;
    int hs = hist_wait(h->e_hist, st);
    if (hs) return hs;
    CU(launch_upfir_q15(p, st));
    CicLaunch t{};
    t.fin = to_fmt(h->cd.in); t.C = p.C; t.interleaved = p.interleaved; t.in = d_in; t.n = n;
    t.tail = h->d_tail[h->cur]; t.tail_next = h->d_tail[h->cur ^ 1]; t.H = h->H;
    CU(launch_cic_tail(t, st));
    if ((hs = hist_mark(h->e_hist, st))) return hs;
    h->cur ^= 1;
    h->n_seen += n;
    return B2D_OK;
  }
  const uint32_t C = h->cd.n_channels;
  const size_t mid_bytes = (size_t)container_bytes(h->cd.out.W) * C * std::max<size_t>(n_out, 1);
  if (mid_bytes > h->mid_cap) {
    if (h->d_mid) cudaFree(h->d_mid);
    h->d_mid = nullptr; h->mid_cap = 0;
    if (cudaMalloc(&h->d_mid, mid_bytes) != cudaSuccess) { cudaGetLastError(); return fail(B2D_ENOMEM, "cudaMalloc(%zu)", mid_bytes); }
    h->mid_cap = mid_bytes;
  }
  size_t n_mid = 0, n_fir = 0;
  int s = b2d_cic_run_dev(h->cic, d_in, n, h->d_mid, &n_mid, st);
  if (s) return s;
  if (n_mid != n_out) return fail(B2D_ESTATE, "cascade count mismatch");
  return b2d_fir_run_dev(h->fir, h->d_mid, n_mid, d_out, &n_fir, st);
}

static int cicfir_ready(b2d_cicfir *h) {
  if (h->fused) { for (char c : h->ch_loaded) if (!c) return fail(B2D_ESTATE, "run() before the coefficients of every channel were loaded"); }
  else if (!all_loaded(h->fir)) return fail(B2D_ESTATE, "run() before the coefficients of every channel were loaded");
  return B2D_OK;
}

extern "C" int b2d_cicfir_run_dev(b2d_cicfir *h, const void *d_in, size_t n, void *d_out, size_t *n_out, void *cuda_stream) {
  TraceRange trace__("b2d_cicfir_run_dev");
  if (!h || (n && !d_in)) return fail(B2D_EINVAL, "null argument");
  int st = cicfir_ready(h);
  if (st) return st;
  const size_t no = cicfir_count(h, n);
  if (no && !d_out) return fail(B2D_EINVAL, "null output");
  if ((st = use_device(h->device))) return st;
  if ((st = cicfir_launch(h, d_in, n, d_out, no, (cudaStream_t)cuda_stream))) return st;
  if (n_out) *n_out = no;
  return B2D_OK;
}

extern "C" int b2d_cicfir_run(b2d_cicfir *h, const void *in, size_t n, void *out, size_t *n_out) {
  TraceRange trace__("b2d_cicfir_run");
  if (!h || (n && !in)) return fail(B2D_EINVAL, "null argument");
  int st = cicfir_ready(h);
  if (st) return st;
  const size_t no_total = cicfir_count(h, n);
  if (no_total && !out) return fail(B2D_EINVAL, "null output");
  if (n_out) *n_out = no_total;
  if (n == 0) return B2D_OK;
  if ((st = use_device(h->device))) return st;
  HostRun r;
  r.in = in; r.out = out; r.n = n; r.C = h->cd.n_channels; r.il = h->cd.layout == B2D_INTERLEAVED;
  r.in_bytes = container_bytes(h->cd.in.W); r.out_bytes = h->out_bytes; r.wire_bytes = wire_bytes_of(h->fo.W, h->wire);
  r.out_like_in = false; r.no_total = no_total;
  r.L = pipe_chunk(n, r.C * (r.in_bytes + (double)r.wire_bytes * h->cd.R));
  r.Lout = r.L * h->cd.R;
  return run_host_pipeline(h->pipe, r, [h](size_t len) { return cicfir_count(h, len); },
                           [h](const void *d_in, size_t len, void *d_out, size_t no, cudaStream_t s) { return cicfir_launch(h, d_in, len, d_out, no, s); });
}

extern "C" int b2d_cicfir_set_wire(b2d_cicfir *h, int32_t wire) {
  if (!h) return fail(B2D_EINVAL, "null handle");
  int st = check_wire(wire);
  if (st) return st;
  h->wire = wire;
  return B2D_OK;
}

extern "C" int b2d_cicfir_reset(b2d_cicfir *h) {
  if (!h) return fail(B2D_EINVAL, "null handle");
  int st = use_device(h->device);
  if (st) return st;
  CU(cudaDeviceSynchronize());
  if (h->fused) {
    for (int i = 0; i < 2; i++) CU(cudaMemset(h->d_tail[i], 0, (size_t)h->H * h->cd.n_channels * 2));
    h->n_seen = 0;
    return B2D_OK;
  }
  if ((st = b2d_cic_reset(h->cic))) return st;
  return b2d_fir_reset(h->fir);
}

// Checkpoint of the cascade: fused = input history + count; two-stage = the two stage blobs back to back.
extern "C" int b2d_cicfir_state_bytes(b2d_cicfir *h, size_t *bytes) {
  if (!h || !bytes) return fail(B2D_EINVAL, "null argument");
  if (h->fused) { *bytes = sizeof(StateHdr) + (size_t)h->H * h->cd.n_channels * 2; return B2D_OK; }
  size_t a = 0, b = 0;
  int st;
  if ((st = b2d_cic_state_bytes(h->cic, &a)) || (st = b2d_fir_state_bytes(h->fir, &b))) return st;
  *bytes = sizeof(StateHdr) + a + b;
  return B2D_OK;
}
extern "C" int b2d_cicfir_get_state(b2d_cicfir *h, void *blob, size_t bytes) {
  if (!h || !blob) return fail(B2D_EINVAL, "null argument");
  int st = use_device(h->device);
  if (st) return st;
  if (h->fused) {
    const StatePart parts[1] = {{h->d_tail[h->cur], (size_t)h->H * h->cd.n_channels * 2}};
    return state_get(StateHdr{kCasMagic, 1, h->n_seen, (uint32_t)h->H, h->cd.n_channels, 2, 1}, parts, 1, blob, bytes);
  }
  size_t need = 0, a = 0;
  if ((st = b2d_cicfir_state_bytes(h, &need))) return st;
  if (bytes < need) return fail(B2D_EINVAL, "state blob needs %zu bytes", need);
  b2d_cic_state_bytes(h->cic, &a);
  const StateHdr hd{kCasMagic, 1, 0, 0, h->cd.n_channels, 2, 0};
  memcpy(blob, &hd, sizeof(hd));
  if ((st = b2d_cic_get_state(h->cic, (char *)blob + sizeof(hd), a))) return st;
  return b2d_fir_get_state(h->fir, (char *)blob + sizeof(hd) + a, need - sizeof(hd) - a);
}
extern "C" int b2d_cicfir_set_state(b2d_cicfir *h, const void *blob, size_t bytes) {
  if (!h || !blob) return fail(B2D_EINVAL, "null argument");
  int st = use_device(h->device);
  if (st) return st;
  if (h->fused) {
    const StatePart parts[1] = {{h->d_tail[h->cur], (size_t)h->H * h->cd.n_channels * 2}};
    StateHdr got;
    if ((st = state_set(StateHdr{kCasMagic, 1, 0, (uint32_t)h->H, h->cd.n_channels, 2, 1}, parts, 1, blob, bytes, &got))) return st;
    if (got.pad != 1) return fail(B2D_EINVAL, "state blob was taken from a two-stage cascade");
    h->n_seen = got.n_seen;
    return B2D_OK;
  }
  size_t need = 0, a = 0;
  if ((st = b2d_cicfir_state_bytes(h, &need))) return st;
  if (bytes < need) return fail(B2D_EINVAL, "state blob needs %zu bytes", need);
  StateHdr hd;
  memcpy(&hd, blob, sizeof(hd));
  if (hd.magic != kCasMagic || hd.pad != 0 || hd.channels != h->cd.n_channels) return fail(B2D_EINVAL, "state blob does not belong to this cascade");
  b2d_cic_state_bytes(h->cic, &a);
  if ((st = b2d_cic_set_state(h->cic, (const char *)blob + sizeof(hd), a))) return st;
  return b2d_fir_set_state(h->fir, (const char *)blob + sizeof(hd) + a, need - sizeof(hd) - a);
}
