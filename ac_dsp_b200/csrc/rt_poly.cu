// rt_poly.cu -- host runtime of the polyphase handles (ac_poly_dec, ac_poly_intr).
#include "rt_common.h"

using namespace b2d;

// -------------------------------------------------------------------------------------------- ac_poly_dec
struct b2d_polydec {
  b2d_polydec_desc d;
  Fmt fin, fc, fa, fo;
  int device = 0, in_bytes = 2, out_bytes = 8, c_bytes = 2, wide = 0, T = 0;
  std::vector<char> ch_loaded;
  int64_t *d_coeff64 = nullptr;
  int32_t *d_coeff32 = nullptr;
  uint32_t *d_coeff_pk = nullptr;
  void *d_tail[2] = {nullptr, nullptr};
  int cur = 0;
  unsigned long long n_seen = 0;
  int wire = B2D_WIRE_CONTAINER;
  cudaEvent_t e_hist = nullptr;
  Pipe pipe;
};

extern "C" int b2d_polydec_destroy(b2d_polydec *h) {
  if (!h) return B2D_OK;
  use_device(h->device);
  cudaDeviceSynchronize();
  h->pipe.destroy();
  if (h->e_hist) cudaEventDestroy(h->e_hist);
  if (h->d_coeff64) cudaFree(h->d_coeff64);
  if (h->d_coeff32) cudaFree(h->d_coeff32);
  if (h->d_coeff_pk) cudaFree(h->d_coeff_pk);
  for (int i = 0; i < 2; i++) if (h->d_tail[i]) cudaFree(h->d_tail[i]);
  delete h;
  return B2D_OK;
}

extern "C" int b2d_polydec_create(b2d_polydec **out, const b2d_polydec_desc *desc) {
  if (!out || !desc) return fail(B2D_EINVAL, "null argument");
  *out = nullptr;
  int st;
  if ((st = check_fmt(desc->in, 32, "IN_TYPE"))) return st;
  if ((st = check_fmt(desc->coeff, 32, "COEFF_TYPE"))) return st;
  if ((st = check_fmt(desc->acc, 64, "ACC_TYPE"))) return st;
  if ((st = check_fmt(desc->out, 64, "OUT_TYPE"))) return st;
  if (desc->n_taps < 1 || desc->df < 1 || (uint64_t)desc->n_taps * desc->df > (1u << 20)) return fail(B2D_EINVAL, "NTAPS = %u, DF = %u invalid", desc->n_taps, desc->df);
  if (desc->n_channels < 1) return fail(B2D_EINVAL, "n_channels must be >= 1");
  if (desc->layout != B2D_PLANAR && desc->layout != B2D_INTERLEAVED) return fail(B2D_EINVAL, "bad layout");
  const Fmt fin = to_fmt(desc->in), fc = to_fmt(desc->coeff), fa = to_fmt(desc->acc), fo = to_fmt(desc->out);
  {  // bit budget of the 128-bit generic evaluation (as for the FIR classes)
    const int Fp = fin.F() + fc.F(), Wp = fin.W + fc.W + 2, rF = std::max(Fp, fa.F());
    if (fa.W + (rF - fa.F()) > 125 || Wp + (rF - Fp) > 125 || fa.W + std::max(0, fo.F() - fa.F()) > 125)
      return fail(B2D_EUNSUPPORTED, "format combination exceeds the 128-bit intermediate budget");
  }
  int dev = desc->device;
  if (dev < 0) CU(cudaGetDevice(&dev));
  if ((st = use_device(dev))) return st;
  b2d_polydec *h = new (std::nothrow) b2d_polydec();
  if (!h) return fail(B2D_ENOMEM, "handle");
  h->d = *desc; h->fin = fin; h->fc = fc; h->fa = fa; h->fo = fo; h->device = dev;
  h->in_bytes = container_bytes(fin.W); h->out_bytes = container_bytes(fo.W); h->c_bytes = container_bytes(fc.W);
  const uint32_t C = desc->n_channels;
  const size_t L = (size_t)desc->n_taps * desc->df;
  h->T = (int)L - 1;
  h->ch_loaded.assign(C, 0);
  h->wide = polydec_wide_mode(fin, fc, fa, (int)desc->n_taps, (int)desc->df) >= 0;
  const char *force = getenv("B2D_FORCE_GENERIC");
  if (h->wide && polydec_q15_supported(fin, fc, fa, (int)desc->n_taps, (int)desc->df) && !(force && *force == '2')) h->wide = 2;
  if (force && *force == '1') h->wide = 0;
  cudaError_t e = cudaMalloc(&h->d_coeff64, C * L * sizeof(int64_t));
  if (e == cudaSuccess && h->wide == 2) e = cudaMalloc(&h->d_coeff_pk, (size_t)C * polydec_q15_words((int)desc->n_taps, (int)desc->df) * sizeof(uint32_t));
  if (e == cudaSuccess && h->wide == 1) e = cudaMalloc(&h->d_coeff32, (size_t)C * desc->df * polydec_words((int)desc->n_taps) * sizeof(int32_t));
  const size_t tail_bytes = std::max<size_t>((size_t)h->T * C * h->in_bytes, 16);
  for (int i = 0; i < 2 && e == cudaSuccess; i++) {
    e = cudaMalloc(&h->d_tail[i], tail_bytes);
    if (e == cudaSuccess) e = cudaMemset(h->d_tail[i], 0, tail_bytes);
  }
  if (e != cudaSuccess) { cudaGetLastError(); b2d_polydec_destroy(h); return fail(B2D_ECUDA, "b2d_polydec_create: %s", cudaGetErrorString(e)); }
  *out = h;
  return B2D_OK;
}

extern "C" const char *b2d_polydec_path(b2d_polydec *h) { return !h ? "" : (h->wide == 2 ? "polydec_q15" : (h->wide ? "polydec_wide" : "polydec_generic")); }
extern "C" size_t b2d_polydec_max_out(b2d_polydec *h, size_t n) { return h ? n / h->d.df + 1 : 0; }

extern "C" int b2d_polydec_load(b2d_polydec *h, const void *coeff_raw, size_t n, int32_t channel) {
  TraceRange trace__("b2d_polydec_load");
  if (!h || !coeff_raw) return fail(B2D_EINVAL, "null argument");
  const size_t L = (size_t)h->d.n_taps * h->d.df;
  const uint32_t C = h->d.n_channels;
  if (n != L) return fail(B2D_EINVAL, "expected %zu coefficients (NTAPS * DF), got %zu", L, n);
  if (channel < -1 || channel >= (int32_t)C) return fail(B2D_EINVAL, "channel %d outside -1..%u", channel, C - 1);
  int st = use_device(h->device);
  if (st) return st;
  std::vector<int64_t> v(L);
  widen_coeffs(coeff_raw, L, h->c_bytes, h->fc, v.data());
  CU(cudaDeviceSynchronize());
  const int words = polydec_words((int)h->d.n_taps);
  std::vector<int32_t> pk;
  std::vector<uint32_t> pq;
  if (h->wide == 1) { pk.assign((size_t)h->d.df * words, 0); polydec_pack(v.data(), (int)h->d.n_taps, (int)h->d.df, pk.data()); }
  if (h->wide == 2) { pq.assign((size_t)polydec_q15_words((int)h->d.n_taps, (int)h->d.df), 0); polydec_q15_pack(h->fc, v.data(), (int)h->d.n_taps, (int)h->d.df, pq.data()); }
  for (uint32_t c = 0; c < C; c++) {
    if (channel >= 0 && (uint32_t)channel != c) continue;
    CU(cudaMemcpy(h->d_coeff64 + c * L, v.data(), L * sizeof(int64_t), cudaMemcpyHostToDevice));
    if (h->wide == 1) CU(cudaMemcpy(h->d_coeff32 + (size_t)c * pk.size(), pk.data(), pk.size() * sizeof(int32_t), cudaMemcpyHostToDevice));
    if (h->wide == 2) CU(cudaMemcpy(h->d_coeff_pk + (size_t)c * pq.size(), pq.data(), pq.size() * sizeof(uint32_t), cudaMemcpyHostToDevice));
    h->ch_loaded[c] = 1;
  }
  return B2D_OK;
}

static size_t polydec_count(const b2d_polydec *h, size_t n) {
  return (size_t)((h->n_seen + n) / h->d.df - h->n_seen / h->d.df);
}

static int polydec_launch(b2d_polydec *h, const void *d_in, size_t n, void *d_out, size_t n_out, cudaStream_t st) {
  if (n == 0) return B2D_OK;
  DecLaunch p;
  p.fin = h->fin; p.fcoeff = h->fc; p.facc = h->fa; p.fout = h->fo;
  p.nt = (int)h->d.n_taps; p.df = (int)h->d.df; p.wide = h->wide; p.C = h->d.n_channels; p.interleaved = h->d.layout == B2D_INTERLEAVED;
  p.in = d_in; p.out = d_out; p.n = n; p.n_out = n_out; p.n_seen = h->n_seen; p.tail = h->d_tail[h->cur];
  p.coeff64 = h->d_coeff64; p.coeff32 = h->d_coeff32; p.coeff_pk = h->d_coeff_pk;
  int hs = hist_wait(h->e_hist, st);
  if (hs) return hs;
  CU(launch_polydec(p, st));
  FirLaunch t{};                    // history carry: the last NTAPS*DF - 1 samples, exactly as for an FIR of that length
  t.fin = h->fin; t.n_taps = h->T + 1; t.C = p.C; t.interleaved = p.interleaved; t.in = d_in; t.n = n;
  t.tail = h->d_tail[h->cur]; t.tail_next = h->d_tail[h->cur ^ 1];
  CU(launch_fir_tail(t, st));
  if ((hs = hist_mark(h->e_hist, st))) return hs;
  h->cur ^= 1;
  h->n_seen += n;
  return B2D_OK;
}

static int polydec_ready(const b2d_polydec *h) {
  for (char c : h->ch_loaded) if (!c) return fail(B2D_ESTATE, "run() before the coefficients of every channel were loaded");
  return B2D_OK;
}

extern "C" int b2d_polydec_run_dev(b2d_polydec *h, const void *d_in, size_t n, void *d_out, size_t *n_out, void *cuda_stream) {
  TraceRange trace__("b2d_polydec_run_dev");
  if (!h || (n && !d_in)) return fail(B2D_EINVAL, "null argument");
  int st = polydec_ready(h);
  if (st) return st;
  const size_t no = polydec_count(h, n);
  if (no && !d_out) return fail(B2D_EINVAL, "null output");
  if ((st = use_device(h->device))) return st;
  if ((st = polydec_launch(h, d_in, n, d_out, no, (cudaStream_t)cuda_stream))) return st;
  if (n_out) *n_out = no;
  return B2D_OK;
}

extern "C" int b2d_polydec_run(b2d_polydec *h, const void *in, size_t n, void *out, size_t *n_out) {
  TraceRange trace__("b2d_polydec_run");
  if (!h || (n && !in)) return fail(B2D_EINVAL, "null argument");
  int st = polydec_ready(h);
  if (st) return st;
  const size_t no_total = polydec_count(h, n);
  if (no_total && !out) return fail(B2D_EINVAL, "null output");
  if (n_out) *n_out = no_total;
  if (n == 0) return B2D_OK;
  if ((st = use_device(h->device))) return st;
  HostRun r;
  r.in = in; r.out = out; r.n = n; r.C = h->d.n_channels; r.il = h->d.layout == B2D_INTERLEAVED;
  r.in_bytes = h->in_bytes; r.out_bytes = h->out_bytes; r.wire_bytes = wire_bytes_of(h->fo.W, h->wire);
  r.out_like_in = false; r.no_total = no_total;
  r.L = pipe_chunk(n, r.C * (r.in_bytes + (double)r.wire_bytes / h->d.df));
  r.Lout = r.L / h->d.df + 1;
  return run_host_pipeline(h->pipe, r, [h](size_t len) { return polydec_count(h, len); },
                           [h](const void *d_in, size_t len, void *d_out, size_t no, cudaStream_t s) { return polydec_launch(h, d_in, len, d_out, no, s); });
}

extern "C" int b2d_polydec_set_wire(b2d_polydec *h, int32_t wire) {
  if (!h) return fail(B2D_EINVAL, "null handle");
  int st = check_wire(wire);
  if (st) return st;
  h->wire = wire;
  return B2D_OK;
}

extern "C" int b2d_polydec_reset(b2d_polydec *h) {
  if (!h) return fail(B2D_EINVAL, "null handle");
  int st = use_device(h->device);
  if (st) return st;
  CU(cudaDeviceSynchronize());
  const size_t tail_bytes = std::max<size_t>((size_t)h->T * h->d.n_channels * h->in_bytes, 16);
  for (int i = 0; i < 2; i++) CU(cudaMemset(h->d_tail[i], 0, tail_bytes));
  h->n_seen = 0;
  return B2D_OK;
}

extern "C" int b2d_polydec_state_bytes(b2d_polydec *h, size_t *bytes) {
  if (!h || !bytes) return fail(B2D_EINVAL, "null argument");
  *bytes = sizeof(StateHdr) + (size_t)h->T * h->d.n_channels * h->in_bytes;
  return B2D_OK;
}
extern "C" int b2d_polydec_get_state(b2d_polydec *h, void *blob, size_t bytes) {
  if (!h) return fail(B2D_EINVAL, "null handle");
  int st = use_device(h->device);
  if (st) return st;
  const StatePart parts[1] = {{h->d_tail[h->cur], (size_t)h->T * h->d.n_channels * h->in_bytes}};
  return state_get(StateHdr{kDecMagic, 1, h->n_seen, (uint32_t)h->T, h->d.n_channels, (uint32_t)h->in_bytes, 0}, parts, 1, blob, bytes);
}
extern "C" int b2d_polydec_set_state(b2d_polydec *h, const void *blob, size_t bytes) {
  if (!h) return fail(B2D_EINVAL, "null handle");
  int st = use_device(h->device);
  if (st) return st;
  const StatePart parts[1] = {{h->d_tail[h->cur], (size_t)h->T * h->d.n_channels * h->in_bytes}};
  StateHdr got;
  if ((st = state_set(StateHdr{kDecMagic, 1, 0, (uint32_t)h->T, h->d.n_channels, (uint32_t)h->in_bytes, 0}, parts, 1, blob, bytes, &got))) return st;
  h->n_seen = got.n_seen;
  return B2D_OK;
}

// -------------------------------------------------------------------------------------------- ac_poly_intr
struct b2d_polyintr {
  b2d_polyintr_desc d;
  Fmt fin, fc, fa, fo;
  int device = 0, in_bytes = 2, out_bytes = 8, c_bytes = 2, csz = 0, H = 0;
  int mode = 0;                      // 0 generic, 1 wide (64-bit modular), 2 q15 (upfir_lane DP2A kernel)
  int lsh = 0, planes = 2, words = 0;
  bool init = false;                 // folded forms: a step has been taken (ac_poly_intr.h:165)
  std::vector<char> ch_loaded;
  int64_t *d_coeff64 = nullptr;
  uint32_t *d_cw = nullptr;
  uint8_t *d_sign = nullptr, *d_corr = nullptr;
  int64_t *d_carry[2] = {nullptr, nullptr};
  void *d_tail[2] = {nullptr, nullptr};
  int cur = 0, ccur = 0;
  unsigned long long n_seen = 0;
  int wire = B2D_WIRE_CONTAINER;
  cudaEvent_t e_hist = nullptr;
  Pipe pipe;
};

extern "C" int b2d_polyintr_destroy(b2d_polyintr *h) {
  if (!h) return B2D_OK;
  use_device(h->device);
  cudaDeviceSynchronize();
  h->pipe.destroy();
  if (h->e_hist) cudaEventDestroy(h->e_hist);
  if (h->d_coeff64) cudaFree(h->d_coeff64);
  if (h->d_cw) cudaFree(h->d_cw);
  if (h->d_sign) cudaFree(h->d_sign);
  if (h->d_corr) cudaFree(h->d_corr);
  for (int i = 0; i < 2; i++) { if (h->d_tail[i]) cudaFree(h->d_tail[i]); if (h->d_carry[i]) cudaFree(h->d_carry[i]); }
  delete h;
  return B2D_OK;
}

static int polyintr_csz(uint32_t nt, uint32_t ifac, int ftype) {
  return (int)(ifac * (ftype == B2D_PI_FOLD_EVEN ? nt / 2 : (ftype == B2D_PI_FOLD_ODD ? nt / 2 + 1 : nt)));
}

// FOLD_ANTI on 16-bit operands with an exact wrapping accumulator is the plain polyphase FIR of upfir_q15.cu
static bool polyintr_q15_ok(const b2d_polyintr_desc &d, int *lsh) {
  const Fmt in = to_fmt(d.in), fc = to_fmt(d.coeff), fa = to_fmt(d.acc);
  if (d.ftype != B2D_PI_FOLD_ANTI) return false;
  if (in.W > 16 || (!in.S && in.W == 16) || fc.W > 16 || (!fc.S && fc.W == 16)) return false;
  if (fa.O != B2D_WRAP || (fa.Q != B2D_TRN && fa.Q != B2D_RND)) return false;
  const int s = in.F() + fc.F() - fa.F();
  if (s > 0 || -s > 40 || -s >= fa.W) return false;
  if (!upfir_q15_geometry((int)d.intr_factor, (int)(d.n_taps * d.intr_factor), 16)) return false;
  *lsh = -s;
  return true;
}

extern "C" int b2d_polyintr_create(b2d_polyintr **out, const b2d_polyintr_desc *desc) {
  if (!out || !desc) return fail(B2D_EINVAL, "null argument");
  *out = nullptr;
  int st;
  if ((st = check_fmt(desc->in, 32, "IN_TYPE"))) return st;
  if ((st = check_fmt(desc->coeff, 32, "COEFF_TYPE"))) return st;
  if ((st = check_fmt(desc->acc, 64, "ACC_TYPE"))) return st;
  if ((st = check_fmt(desc->out, 64, "OUT_TYPE"))) return st;
  if (desc->n_taps < 1 || desc->n_taps > (1u << 16)) return fail(B2D_EINVAL, "NTAPS = %u outside 1..65536", desc->n_taps);
  if (desc->intr_factor < 1 || desc->intr_factor > 255) return fail(B2D_EINVAL, "IF = %u outside 1..255", desc->intr_factor);
  if (desc->ftype < B2D_PI_FOLD_EVEN || desc->ftype > B2D_PI_FOLD_ANTI) return fail(B2D_EINVAL, "bad ftype");
  if (desc->n_channels < 1) return fail(B2D_EINVAL, "n_channels must be >= 1");
  if (desc->layout != B2D_PLANAR && desc->layout != B2D_INTERLEAVED) return fail(B2D_EINVAL, "bad layout");
  const Fmt fin = to_fmt(desc->in), fc = to_fmt(desc->coeff), fa = to_fmt(desc->acc), fo = to_fmt(desc->out);
  {  // bit budget of the 128-bit generic evaluation: the folded forms multiply COEFF_TYPE by an ACC_TYPE fold
    const bool folded = desc->ftype != B2D_PI_FOLD_ANTI;
    const int Fp = (folded ? fa.F() : fin.F()) + fc.F(), Wp = (folded ? fa.W : fin.W) + fc.W + 2, rF = std::max(Fp, fa.F());
    if (fa.W + (rF - fa.F()) > 125 || Wp + (rF - Fp) > 125 || fa.W + 2 + std::max(0, fo.F() - fa.F()) > 125 ||
        fin.W + 2 + std::max(0, fa.F() - fin.F()) > 125)
      return fail(B2D_EUNSUPPORTED, "format combination exceeds the 128-bit intermediate budget");
  }
  int dev = desc->device;
  if (dev < 0) CU(cudaGetDevice(&dev));
  if ((st = use_device(dev))) return st;
  b2d_polyintr *h = new (std::nothrow) b2d_polyintr();
  if (!h) return fail(B2D_ENOMEM, "handle");
  h->d = *desc; h->fin = fin; h->fc = fc; h->fa = fa; h->fo = fo; h->device = dev;
  h->in_bytes = container_bytes(fin.W); h->out_bytes = container_bytes(fo.W); h->c_bytes = container_bytes(fc.W);
  h->csz = polyintr_csz(desc->n_taps, desc->intr_factor, desc->ftype);
  const uint32_t C = desc->n_channels, IF = desc->intr_factor;
  h->H = (int)desc->n_taps + 2;
  h->ch_loaded.assign(C, 0);
  h->mode = polyintr_fast_supported(fin, fc, fa, desc->ftype) ? 1 : 0;
  const char *force = getenv("B2D_FORCE_GENERIC");
  if (h->mode && polyintr_q15_ok(*desc, &h->lsh) && !(force && *force == '2')) h->mode = 2;
  if (force && *force == '1') h->mode = 0;
  if (h->mode == 2) { h->planes = 2; h->words = upfir_q15_words((int)IF, (int)(desc->n_taps * IF), h->planes); }
  cudaError_t e = cudaMalloc(&h->d_coeff64, (size_t)C * std::max(h->csz, 1) * sizeof(int64_t));
  if (e == cudaSuccess && h->mode == 2) e = cudaMalloc(&h->d_cw, (size_t)C * h->words * sizeof(uint32_t));
  if (e == cudaSuccess) e = cudaMalloc(&h->d_sign, (size_t)C * IF);
  if (e == cudaSuccess) e = cudaMalloc(&h->d_corr, (size_t)C * IF);
  const size_t tail_bytes = std::max<size_t>((size_t)h->H * C * h->in_bytes, 16);
  for (int i = 0; i < 2 && e == cudaSuccess; i++) {
    e = cudaMalloc(&h->d_tail[i], tail_bytes);
    if (e == cudaSuccess) e = cudaMemset(h->d_tail[i], 0, tail_bytes);
    if (e == cudaSuccess) e = cudaMalloc(&h->d_carry[i], (size_t)C * IF * sizeof(int64_t));
    if (e == cudaSuccess) e = cudaMemset(h->d_carry[i], 0, (size_t)C * IF * sizeof(int64_t));
  }
  if (e != cudaSuccess) { cudaGetLastError(); b2d_polyintr_destroy(h); return fail(B2D_ECUDA, "b2d_polyintr_create: %s", cudaGetErrorString(e)); }
  *out = h;
  return B2D_OK;
}

extern "C" const char *b2d_polyintr_path(b2d_polyintr *h) { return !h ? "" : (h->mode == 2 ? "polyintr_q15" : (h->mode ? "polyintr_wide" : "polyintr_generic")); }
extern "C" size_t b2d_polyintr_coeffsz(b2d_polyintr *h) { return h ? (size_t)h->csz : 0; }
extern "C" size_t b2d_polyintr_max_out(b2d_polyintr *h, size_t n) { return h ? n * h->d.intr_factor : 0; }

extern "C" int b2d_polyintr_load(b2d_polyintr *h, const void *coeff_raw, size_t n, const uint8_t *sign, const uint8_t *corr, int32_t channel) {
  TraceRange trace__("b2d_polyintr_load");
  if (!h || (!coeff_raw && h->csz)) return fail(B2D_EINVAL, "null argument");
  const uint32_t C = h->d.n_channels, IF = h->d.intr_factor;
  const size_t L = (size_t)h->csz;
  if (n != L) return fail(B2D_EINVAL, "expected %zu coefficients, got %zu", L, n);
  if (channel < -1 || channel >= (int32_t)C) return fail(B2D_EINVAL, "channel %d outside -1..%u", channel, C - 1);
  std::vector<uint8_t> sg(IF, 1), cr(IF);
  for (uint32_t j = 0; j < IF; j++) {
    cr[j] = corr ? corr[j] : (uint8_t)j;
    if (sign) sg[j] = sign[j] ? 1 : 0;
    if (cr[j] >= IF) return fail(B2D_EINVAL, "corr[%u] = %u outside 0..IF-1 (the reference would index acc_a / acc_b out of range)", j, cr[j]);
  }
  int st = use_device(h->device);
  if (st) return st;
  std::vector<int64_t> v(std::max<size_t>(L, 1));
  widen_coeffs(coeff_raw, L, h->c_bytes, h->fc, v.data());
  CU(cudaDeviceSynchronize());
  std::vector<uint32_t> pk;
  if (h->mode == 2) {      // composite taps of the polyphase form: c[ph + IF*m] = coeffs[m + NTAPS*ph]
    const int NT = (int)h->d.n_taps;
    std::vector<int64_t> comp((size_t)NT * IF);
    for (uint32_t ph = 0; ph < IF; ph++)
      for (int m = 0; m < NT; m++) comp[ph + (size_t)IF * m] = v[m + (size_t)NT * ph];
    pk.assign((size_t)h->words, 0);
    upfir_q15_pack(comp.data(), NT * (int)IF, (int)IF, h->planes, pk.data());
  }
  for (uint32_t c = 0; c < C; c++) {
    if (channel >= 0 && (uint32_t)channel != c) continue;
    if (L) CU(cudaMemcpy(h->d_coeff64 + c * L, v.data(), L * sizeof(int64_t), cudaMemcpyHostToDevice));
    if (h->mode == 2) CU(cudaMemcpy(h->d_cw + (size_t)c * h->words, pk.data(), pk.size() * sizeof(uint32_t), cudaMemcpyHostToDevice));
    CU(cudaMemcpy(h->d_sign + (size_t)c * IF, sg.data(), IF, cudaMemcpyHostToDevice));
    CU(cudaMemcpy(h->d_corr + (size_t)c * IF, cr.data(), IF, cudaMemcpyHostToDevice));
    h->ch_loaded[c] = 1;
  }
  return B2D_OK;
}

static size_t polyintr_rows(const b2d_polyintr *h, size_t n) {
  if (h->d.ftype == B2D_PI_FOLD_ANTI || h->init) return n;
  return n ? n - 1 : 0;
}

static int polyintr_launch(b2d_polyintr *h, const void *d_in, size_t n, void *d_out, size_t n_rows, cudaStream_t st) {
  if (n == 0) return B2D_OK;
  const uint32_t C = h->d.n_channels;
  const bool il = h->d.layout == B2D_INTERLEAVED;
  int hs = hist_wait(h->e_hist, st);
  if (hs) return hs;
  if (h->mode == 2) {
    UpLaunch p;
    p.facc = h->fa; p.fout = h->fo; p.R = (int)h->d.intr_factor; p.taps_total = (int)(h->d.n_taps * h->d.intr_factor);
    p.planes = h->planes; p.lsh = h->lsh; p.C = C; p.interleaved = il;
    p.in = d_in; p.out = d_out; p.n = n; p.n_out = n * h->d.intr_factor;
    p.n_seen = h->n_seen; p.out_first = h->n_seen * h->d.intr_factor;
    p.tail = h->d_tail[h->cur]; p.H = h->H; p.cw = h->d_cw;
    CU(launch_upfir_q15(p, st));
  } else {
    PiLaunch p;
    p.fin = h->fin; p.fcoeff = h->fc; p.facc = h->fa; p.fout = h->fo;
    p.nt = (int)h->d.n_taps; p.ifac = (int)h->d.intr_factor; p.ftype = h->d.ftype; p.csz = h->csz; p.fast = h->mode == 1;
    p.C = C; p.interleaved = il; p.in = d_in; p.out = d_out; p.n = n; p.n_rows = n_rows;
    p.row_shift = (h->d.ftype != B2D_PI_FOLD_ANTI && h->init) ? 1 : 0;
    p.tail = h->d_tail[h->cur]; p.H = h->H; p.coeff64 = h->d_coeff64; p.sign = h->d_sign; p.corr = h->d_corr;
    p.carry = h->d_carry[h->ccur]; p.carry_next = h->d_carry[h->ccur ^ 1];
    CU(launch_polyintr(p, st));
    if (h->d.ftype != B2D_PI_FOLD_ANTI) h->ccur ^= 1;
  }
  CicLaunch t{};
  t.fin = h->fin; t.C = C; t.interleaved = il; t.in = d_in; t.n = n;
  t.tail = h->d_tail[h->cur]; t.tail_next = h->d_tail[h->cur ^ 1]; t.H = h->H;
  CU(launch_cic_tail(t, st));
  if ((hs = hist_mark(h->e_hist, st))) return hs;
  h->cur ^= 1;
  h->n_seen += n;
  h->init = true;
  return B2D_OK;
}

static int polyintr_ready(const b2d_polyintr *h) {
  for (char c : h->ch_loaded) if (!c) return fail(B2D_ESTATE, "run() before the control / coefficient structures of every channel were loaded");
  return B2D_OK;
}

extern "C" int b2d_polyintr_run_dev(b2d_polyintr *h, const void *d_in, size_t n, void *d_out, size_t *n_out, void *cuda_stream) {
  TraceRange trace__("b2d_polyintr_run_dev");
  if (!h || (n && !d_in)) return fail(B2D_EINVAL, "null argument");
  int st = polyintr_ready(h);
  if (st) return st;
  const size_t rows = polyintr_rows(h, n);
  if (rows && !d_out) return fail(B2D_EINVAL, "null output");
  if ((st = use_device(h->device))) return st;
  if ((st = polyintr_launch(h, d_in, n, d_out, rows, (cudaStream_t)cuda_stream))) return st;
  if (n_out) *n_out = rows * h->d.intr_factor;
  return B2D_OK;
}

extern "C" int b2d_polyintr_run(b2d_polyintr *h, const void *in, size_t n, void *out, size_t *n_out) {
  TraceRange trace__("b2d_polyintr_run");
  if (!h || (n && !in)) return fail(B2D_EINVAL, "null argument");
  int st = polyintr_ready(h);
  if (st) return st;
  const uint32_t IF = h->d.intr_factor;
  const size_t no_total = polyintr_rows(h, n) * IF;
  if (no_total && !out) return fail(B2D_EINVAL, "null output");
  if (n_out) *n_out = no_total;
  if (n == 0) return B2D_OK;
  if ((st = use_device(h->device))) return st;
  HostRun r;
  r.in = in; r.out = out; r.n = n; r.C = h->d.n_channels; r.il = h->d.layout == B2D_INTERLEAVED;
  r.in_bytes = h->in_bytes; r.out_bytes = h->out_bytes; r.wire_bytes = wire_bytes_of(h->fo.W, h->wire);
  r.out_like_in = false; r.no_total = no_total;
  r.L = pipe_chunk(n, r.C * (r.in_bytes + (double)r.wire_bytes * IF));
  r.Lout = r.L * IF;
  return run_host_pipeline(h->pipe, r, [h, IF](size_t len) { return polyintr_rows(h, len) * IF; },
                           [h, IF](const void *d_in, size_t len, void *d_out, size_t no, cudaStream_t s) { return polyintr_launch(h, d_in, len, d_out, no / IF, s); });
}

extern "C" int b2d_polyintr_set_wire(b2d_polyintr *h, int32_t wire) {
  if (!h) return fail(B2D_EINVAL, "null handle");
  int st = check_wire(wire);
  if (st) return st;
  h->wire = wire;
  return B2D_OK;
}

extern "C" int b2d_polyintr_reset(b2d_polyintr *h) {
  if (!h) return fail(B2D_EINVAL, "null handle");
  int st = use_device(h->device);
  if (st) return st;
  CU(cudaDeviceSynchronize());
  const size_t tail_bytes = std::max<size_t>((size_t)h->H * h->d.n_channels * h->in_bytes, 16);
  for (int i = 0; i < 2; i++) {
    CU(cudaMemset(h->d_tail[i], 0, tail_bytes));
    CU(cudaMemset(h->d_carry[i], 0, (size_t)h->d.n_channels * h->d.intr_factor * sizeof(int64_t)));
  }
  h->n_seen = 0;
  h->init = false;
  return B2D_OK;
}

static void polyintr_parts(b2d_polyintr *h, StatePart *parts) {
  parts[0] = StatePart{h->d_tail[h->cur], (size_t)h->H * h->d.n_channels * h->in_bytes};
  parts[1] = StatePart{h->d_carry[h->ccur], (size_t)h->d.n_channels * h->d.intr_factor * sizeof(int64_t)};
}
extern "C" int b2d_polyintr_state_bytes(b2d_polyintr *h, size_t *bytes) {
  if (!h || !bytes) return fail(B2D_EINVAL, "null argument");
  StatePart parts[2];
  polyintr_parts(h, parts);
  *bytes = state_total(parts, 2);
  return B2D_OK;
}
// the delay line, the parked accumulators of the last step (ac_poly_intr.h:108-110) and `init`
extern "C" int b2d_polyintr_get_state(b2d_polyintr *h, void *blob, size_t bytes) {
  if (!h) return fail(B2D_EINVAL, "null handle");
  int st = use_device(h->device);
  if (st) return st;
  StatePart parts[2];
  polyintr_parts(h, parts);
  return state_get(StateHdr{kIntrMagic, 1, h->n_seen, (uint32_t)h->H, h->d.n_channels, (uint32_t)h->in_bytes, h->init ? 1u : 0u}, parts, 2, blob, bytes);
}
extern "C" int b2d_polyintr_set_state(b2d_polyintr *h, const void *blob, size_t bytes) {
  if (!h) return fail(B2D_EINVAL, "null handle");
  int st = use_device(h->device);
  if (st) return st;
  StatePart parts[2];
  polyintr_parts(h, parts);
  StateHdr got;
  if ((st = state_set(StateHdr{kIntrMagic, 1, 0, (uint32_t)h->H, h->d.n_channels, (uint32_t)h->in_bytes, 0}, parts, 2, blob, bytes, &got))) return st;
  h->n_seen = got.n_seen;
  h->init = got.pad != 0;
  return B2D_OK;
}
