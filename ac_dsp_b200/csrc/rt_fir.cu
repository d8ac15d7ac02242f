// rt_fir.cu -- host runtime of the FIR handles (ac_fir_const_coeffs / _load_coeffs / _prog_coeffs / ac_fir_reg_share):
// descriptor validation, kernel-family choice, coefficient load incl. the NCCL broadcast and the TRANSPOSED partial-sum
// carry, history ping-pong, checkpoints.
#include "rt_common.h"

using namespace b2d;

// ------------------------------------------------------------------------------------------------ FIR
extern "C" int b2d_fir_create(b2d_fir **out, const b2d_fir_desc *desc) {
  if (!out || !desc) return fail(B2D_EINVAL, "null argument");
  *out = nullptr;
  int st;
  if ((st = check_fmt(desc->in, 32, "IN_TYPE"))) return st;
  if ((st = check_fmt(desc->coeff, 32, "COEFF_TYPE"))) return st;
  if ((st = check_fmt(desc->acc, 64, "ACC_TYPE"))) return st;
  if ((st = check_fmt(desc->out, 64, "OUT_TYPE"))) return st;
  if (desc->n_taps < 1 || desc->n_taps > (1u << 20)) return fail(B2D_EINVAL, "n_taps %u outside 1..2^20", desc->n_taps);
  if (desc->n_channels < 1) return fail(B2D_EINVAL, "n_channels must be >= 1");
  if (desc->layout != B2D_PLANAR && desc->layout != B2D_INTERLEAVED) return fail(B2D_EINVAL, "bad layout");
  if (desc->kind < B2D_FIR_CONST || desc->kind > B2D_FIR_REG_SHARE) return fail(B2D_EINVAL, "bad kind");
  if (desc->ftype < B2D_SHIFT_REG || desc->ftype > B2D_FOLD_ODD_ANTI) return fail(B2D_EINVAL, "bad ftype");
  const bool anti_ft = desc->ftype == B2D_FOLD_EVEN_ANTI || desc->ftype == B2D_FOLD_ODD_ANTI;
  if (desc->kind != B2D_FIR_REG_SHARE && anti_ft)
    return fail(B2D_EUNSUPPORTED, "the const / load / prog FIR classes do not dispatch the _ANTI architectures (output left unwritten)");
  if (desc->kind == B2D_FIR_REG_SHARE && (desc->ftype == B2D_ROTATE_SHIFT || desc->ftype == B2D_C_BUFF || desc->ftype == B2D_TRANSPOSED))
    return fail(B2D_EUNSUPPORTED, "ac_fir_reg_share does not dispatch this architecture (output left unwritten)");
  Fmt fin = to_fmt(desc->in), fc = to_fmt(desc->coeff), fa = to_fmt(desc->acc), fo = to_fmt(desc->out);
  {  // bit budget of the 128-bit generic evaluation
    const int Fp = desc->ftype == B2D_FOLD_ODD ? fc.F() + fa.F() : fin.F() + fc.F();
    const int Wp = desc->ftype == B2D_FOLD_ODD ? fc.W + fa.W : fin.W + fc.W + 2;
    const int rF = std::max(Fp, fa.F());
    if (fa.W + (rF - fa.F()) > 125 || Wp + (rF - Fp) > 125 || fa.W + std::max(0, fo.F() - fa.F()) > 125 ||
        (desc->ftype == B2D_FOLD_ODD && fin.W + 1 + std::max(0, fa.F() - fin.F()) > 125))
      return fail(B2D_EUNSUPPORTED, "format combination exceeds the 128-bit intermediate budget");
  }
  int dev = desc->device;
  if (dev < 0) CU(cudaGetDevice(&dev));
  if ((st = use_device(dev))) return st;
  b2d_fir *h = new (std::nothrow) b2d_fir();
  if (!h) return fail(B2D_ENOMEM, "handle");
  h->d = *desc; h->fin = fin; h->fc = fc; h->fa = fa; h->fo = fo; h->device = dev;
  h->T = (int)desc->n_taps - 1;
  h->in_bytes = container_bytes(fin.W); h->out_bytes = container_bytes(fo.W); h->c_bytes = container_bytes(fc.W);
  const uint32_t C = desc->n_channels;
  const size_t N = desc->n_taps;
  h->h_coeff.assign(C * N, 0);
  h->ch_loaded.assign(C, 0);
  h->path = fir_q15_supported(fin, fc, fa, fo, (int)N, desc->ftype) ? PATH_Q15
            : (fir_q24_supported(fin, fc, fa, fo, (int)N, desc->ftype) ? PATH_Q24
            : (fir_wide_supported(fin, fc, fa, fo, (int)N, desc->ftype) ? PATH_WIDE : PATH_GENERIC));
  const char *force = getenv("B2D_FORCE_GENERIC");
  if (force && *force == '1') h->path = PATH_GENERIC;
  if (force && *force == '2' && fir_wide_supported(fin, fc, fa, fo, (int)N, desc->ftype)) h->path = PATH_WIDE;
  cudaError_t e = cudaMalloc(&h->d_coeff64, C * N * sizeof(int64_t));
  const size_t tail_bytes = std::max<size_t>((size_t)h->T * C * h->in_bytes, 16);
  for (int i = 0; i < 2 && e == cudaSuccess; i++) {
    e = cudaMalloc(&h->d_tail[i], tail_bytes);
    if (e == cudaSuccess) e = cudaMemset(h->d_tail[i], 0, tail_bytes);
  }
  if (e == cudaSuccess && (h->path == PATH_Q15 || h->path == PATH_Q24)) {
    h->pk_words = h->path == PATH_Q15 ? fir_q15_pk_words((int)N, desc->ftype) : fir_q24_pk_words((int)N);
    e = cudaMalloc(&h->d_coeff_pk, (size_t)C * h->pk_words * sizeof(uint32_t));
  }
  if (e == cudaSuccess && h->path == PATH_Q15 && fir_ovs_geometry((int)N, C, desc->layout == B2D_INTERLEAVED)) {
    const char *ov = getenv("B2D_FIR_OVS");        // 0: never, 2: every call whatever its length (tests), default: long calls
    h->ovs_mode = ov ? (*ov == '0' ? 0 : (*ov == '2' ? 2 : 1)) : 1;
    if (h->ovs_mode) {
      std::vector<double2> tw(6 * 256 + 6 * 16);
      fir_ovs_tables(tw.data(), tw.data() + 6 * 256);
      e = cudaMalloc(&h->d_tw, tw.size() * sizeof(double2));
      if (e == cudaSuccess) e = cudaMemcpy(h->d_tw, tw.data(), tw.size() * sizeof(double2), cudaMemcpyHostToDevice);
      if (e == cudaSuccess) e = cudaMalloc(&h->d_hs, (size_t)C * 4096 * sizeof(double2));
      const char *rs = getenv("B2D_OVS_RESID");
      if (e == cudaSuccess && rs && *rs == '1') {
        e = cudaMalloc(&h->d_resid, sizeof(double));
        if (e == cudaSuccess) e = cudaMemset(h->d_resid, 0, sizeof(double));
      }
      h->hs_stale.assign(C, 1);
    }
  }
  if (e == cudaSuccess && desc->kind == B2D_FIR_REG_SHARE) {
    e = cudaMalloc(&h->d_dl, C * sizeof(int64_t));
    if (e == cudaSuccess) e = cudaMemset(h->d_dl, 0, C * sizeof(int64_t));
  }
  if (e == cudaSuccess && h->path == PATH_WIDE) {
    h->wide_words = fir_wide_words((int)N);
    h->wide_mode = fir_wide_mode(fin, fc, fa, (int)N, desc->ftype);
    e = cudaMalloc(&h->d_coeff32, (size_t)C * h->wide_words * sizeof(int32_t));
  }
  if (e != cudaSuccess) {
    cudaGetLastError();
    b2d_fir_destroy(h);
    return fail(e == cudaErrorMemoryAllocation ? B2D_ENOMEM : B2D_ECUDA, "b2d_fir_create: %s", cudaGetErrorString(e));
  }
  *out = h;
  return B2D_OK;
}

extern "C" int b2d_fir_destroy(b2d_fir *h) {
  if (!h) return B2D_OK;
  use_device(h->device);
  cudaDeviceSynchronize();
  h->pipe.destroy();
  if (h->d_coeff64) cudaFree(h->d_coeff64);
  if (h->d_coeff_pk) cudaFree(h->d_coeff_pk);
  if (h->d_coeff32) cudaFree(h->d_coeff32);
  if (h->d_dl) cudaFree(h->d_dl);
  if (h->d_tw) cudaFree(h->d_tw);
  if (h->d_hs) cudaFree(h->d_hs);
  if (h->d_resid) cudaFree(h->d_resid);
  if (h->d_win) cudaFree(h->d_win);
  if (h->e_hist) cudaEventDestroy(h->e_hist);
  for (int i = 0; i < 2; i++) { if (h->d_tail[i]) cudaFree(h->d_tail[i]); if (h->d_pend[i]) cudaFree(h->d_pend[i]); }
  delete h;
  return B2D_OK;
}

// Overlap-save is armed when every channel's loaded taps keep the FP64 error bound under 1/2 (and, for an interleaved IQ
// pair, both channels carry the same taps: the pair is transformed as one complex sequence).
static bool fir_ovs_armed(const b2d_fir *h) {
  if (!h->ovs_mode || !h->d_hs) return false;
  const size_t N = h->d.n_taps;
  const uint32_t C = h->d.n_channels;
  for (uint32_t c = 0; c < C; c++) if (!h->ch_loaded[c]) return false;
  if (h->d.layout == B2D_INTERLEAVED && C == 2 && !std::equal(h->h_coeff.begin(), h->h_coeff.begin() + N, h->h_coeff.begin() + N)) return false;
  return h->ovs_bound < 0.49;
}
// Is a call of n samples per channel worth the overlap-save evaluation?  Kernel durations measured on a B200
// (profiles/r02_ovs_crossover.txt, microseconds): the DP2A kernel takes 5.5 + 0.0095 taps for its first wave plus its MACs at
// 17.8 T/s; overlap-save comes in rounds of one block per CTA half, 13 + 9.1 per round of 296 blocks (IQ pair) or of
// 2 * (148 / C) block pairs per real channel.  256 taps on an IQ pair: from about 5 * 10^5 samples per call; 1024 taps on
// eight real channels: from about 2.5 * 10^4 per channel.  B2D_FIR_OVS=2 takes every call (tests).
static bool fir_ovs_worth(const b2d_fir *h, size_t n) {
  if (h->ovs_mode == 2) return n > 0;
  const size_t L = 4096 - (size_t)fir_ovs_discard((int)h->d.n_taps);
  if (n < L) return false;
  const uint32_t C = h->d.n_channels;
  const double taps = (double)h->d.n_taps;
  const double t_dp2a = 5.5 + 0.0095 * taps + (double)n * C * taps / 17.8e6;
  const size_t blocks = (n + L - 1) / L;
  size_t rounds;
  if (h->d.layout == B2D_INTERLEAVED && C == 2) rounds = (blocks + 295) / 296;
  else {
    const size_t halves = 2 * std::max<size_t>(1, 148 / C);          // per channel
    const size_t waves = (C + 147) / 148;                            // more channels than SMs: the grid rows queue up
    rounds = (((blocks + 1) / 2 + halves - 1) / halves) * waves;
  }
  return 13.0 + 9.1 * (double)rounds < t_dp2a;
}

// Spectra of channels whose taps changed since the last long call (host, extended precision; fir_ovs_spectrum).
static int fir_ovs_prepare(b2d_fir *h, cudaStream_t st) {
  const size_t N = h->d.n_taps;
  bool any = false;
  for (uint32_t c = 0; c < h->d.n_channels; c++) any = any || h->hs_stale[c];
  if (!any) return B2D_OK;
  CU(cudaStreamSynchronize(st));       // earlier launches on this stream may still read the spectra
  std::vector<int64_t> eff(N);
  std::vector<double2> hs(4096);
  for (uint32_t c = 0; c < h->d.n_channels; c++) {
    if (!h->hs_stale[c]) continue;
    if (c && std::equal(h->h_coeff.begin() + (size_t)c * N, h->h_coeff.begin() + (size_t)(c + 1) * N, h->h_coeff.begin() + (size_t)(c - 1) * N) && !h->hs_stale[c - 1]) {
      CU(cudaMemcpy(h->d_hs + (size_t)c * 4096, h->d_hs + (size_t)(c - 1) * 4096, 4096 * sizeof(double2), cudaMemcpyDeviceToDevice));
    } else {
      fir_effective_taps(h->h_coeff.data() + (size_t)c * N, (int)N, h->d.ftype, eff.data());
      fir_ovs_spectrum(eff.data(), (int)N, hs.data());
      CU(cudaMemcpy(h->d_hs + (size_t)c * 4096, hs.data(), 4096 * sizeof(double2), cudaMemcpyHostToDevice));
    }
    h->hs_stale[c] = 0;
  }
  return B2D_OK;
}

static cudaError_t fir_dispatch(const b2d_fir *h, const FirLaunch &p, cudaStream_t st, bool allow_ovs = false) {
  if (allow_ovs && h->path == PATH_Q15 && fir_ovs_worth(h, p.n) && fir_ovs_armed(h))
    return launch_fir_ovs(p, h->d_tw, h->d_hs, h->d_resid, st);
  switch (h->path) {
    case PATH_Q15: return launch_fir_q15(p, st);
    case PATH_Q24: return launch_fir_q24(p, st);
    case PATH_WIDE: return launch_fir_wide(p, st);
    default: return launch_fir_generic(p, st);
  }
}

extern "C" const char *b2d_fir_path(b2d_fir *h) {
  static const char *names[] = {"fir_generic", "fir_q15", "fir_wide", "fir_q24"};
  if (h && h->path == PATH_Q15 && fir_ovs_armed(h)) return "fir_ovs";   // calls long enough to be worth it (fir_ovs_worth); shorter ones: fir_q15
  return !h ? "" : names[h->path];
}

// Overlap-save diagnostics: the a-priori bound of |FP64 result - exact sum| for the loaded taps (the path is armed below
// 0.49) and, with B2D_OVS_RESID=1 in the environment at create time, the largest distance from an integer seen so far.
extern "C" int b2d_fir_ovs_margin(b2d_fir *h, double *bound, double *resid) {
  if (!h) return fail(B2D_EINVAL, "null handle");
  if (bound) *bound = h->ovs_bound;
  if (resid) {
    *resid = -1.0;
    if (h->d_resid) {
      int st = use_device(h->device);
      if (st) return st;
      CU(cudaDeviceSynchronize());
      CU(cudaMemcpy(resid, h->d_resid, sizeof(double), cudaMemcpyDeviceToHost));
    }
  }
  return B2D_OK;
}

extern "C" int b2d_fir_set_comm(b2d_fir *h, b2d_comm *comm, int32_t root) {
  if (!h) return fail(B2D_EINVAL, "null handle");
  if (comm && (root < 0 || root >= comm->world)) return fail(B2D_EINVAL, "root %d outside the communicator", root);
  h->comm = comm; h->root = root;
  return B2D_OK;
}

extern "C" int b2d_fir_load(b2d_fir *h, const void *coeff_raw, size_t n, int32_t channel) {
  TraceRange trace__("b2d_fir_load");
  if (!h) return fail(B2D_EINVAL, "null handle");
  const size_t N = h->d.n_taps;
  const uint32_t C = h->d.n_channels;
  if (n != N) return fail(B2D_EINVAL, "expected %zu coefficients, got %zu", N, n);
  if (channel < -1 || channel >= (int32_t)C) return fail(B2D_EINVAL, "channel %d outside -1..%u", channel, C - 1);
  const bool have_local = !h->comm || h->comm->rank == h->root;
  if (h->d.kind == B2D_FIR_CONST) {  // ac_fir_const_coeffs: the pointer is bound once, at construction
    for (uint32_t c = 0; c < C; c++)
      if ((channel < 0 || (uint32_t)channel == c) && h->ch_loaded[c])
        return fail(B2D_ESTATE, "constant-coefficient filter: coefficients are fixed at construction");
  }
  int st = use_device(h->device);
  if (st) return st;
  // v[N] is a status word: an argument error that only the root can see (null pointer) is broadcast with the payload,
  // so every rank of the communicator fails the same way instead of the others waiting in ncclBroadcast
  std::vector<int64_t> v(N + 1, 0);
  if (have_local) {
    if (coeff_raw) widen_coeffs(coeff_raw, N, h->c_bytes, h->fc, v.data());
    else v[N] = 1;
  }
  if (h->comm && (st = comm_bcast_i64(h->comm, v.data(), N + 1, h->root))) return st;
  if (v[N]) return fail(B2D_EINVAL, "null coefficient pointer%s", have_local ? "" : " on the root rank");
  v.resize(N);
  // the coefficient set may be swapped between run() calls while earlier launches are still in flight
  CU(cudaDeviceSynchronize());
  // TRANSPOSED keeps ACC_TYPE partial sums as its state (reg_trans[], ac_fir_load_coeffs.h:265-278,
  // ac_fir_prog_coeffs.h:232-247): after a coefficient change the next N_TAPS-1 outputs of the reference are old-tap
  // partial sums plus new-tap products.  The engine carries samples, so a CHANGE of taps on a filter that has consumed
  // samples converts its history into those pending sums first (old taps, still on the device, over the history followed
  // by zeros), clears the history of the channel, and the next N_TAPS-1 outputs start from them (fir_launch).  Re-loading
  // equal taps (ac_fir_prog_coeffs passes its array on every call) is not a change.
  if (h->d.ftype == B2D_TRANSPOSED && h->ran && h->T > 0) {
    for (uint32_t c = 0; c < C; c++) {
      if (channel >= 0 && (uint32_t)channel != c) continue;
      if (!h->ch_loaded[c] || std::equal(v.begin(), v.end(), h->h_coeff.begin() + (size_t)c * N)) continue;
      const size_t pend_bytes = (size_t)C * h->T * sizeof(int64_t);
      for (int i = 0; i < 2; i++)
        if (!h->d_pend[i]) {
          CU(cudaMalloc(&h->d_pend[i], pend_bytes));
          CU(cudaMemset(h->d_pend[i], 0, pend_bytes));
        }
      FirLaunch p;
      p.fin = h->fin; p.fcoeff = h->fc; p.facc = h->fa; p.fout = h->fo;
      p.n_taps = (int)N; p.ftype = B2D_TRANSPOSED; p.ascending = 0; p.C = 1; p.interleaved = 0;
      p.in = nullptr; p.out = nullptr; p.n = (size_t)h->T;
      p.tail = (const char *)h->d_tail[h->cur] + (size_t)c * h->T * h->in_bytes; p.tail_next = nullptr;
      p.coeff64 = h->d_coeff64 + (size_t)c * N; p.coeff_pk = nullptr; p.pk_words = 0; p.coeff32 = nullptr;
      int64_t *pend = h->d_pend[h->pcur] + (size_t)c * h->T;
      CU(launch_fir_pending(p, (size_t)h->T, pend, pend, nullptr));      // in place: thread i reads and writes element i only
      CU(cudaMemsetAsync((char *)h->d_tail[h->cur] + (size_t)c * h->T * h->in_bytes, 0, (size_t)h->T * h->in_bytes, nullptr));
      h->pend_rem = (size_t)h->T;
    }
    CU(cudaDeviceSynchronize());
  }
  for (uint32_t c = 0; c < C; c++) {
    if (channel >= 0 && (uint32_t)channel != c) continue;
    // re-loading the taps a channel already has (ac_fir_prog_coeffs passes its array on every call) keeps its spectrum
    const bool same = h->ch_loaded[c] && std::equal(v.begin(), v.end(), h->h_coeff.begin() + c * N);
    std::copy(v.begin(), v.end(), h->h_coeff.begin() + c * N);
    h->ch_loaded[c] = 1;
    CU(cudaMemcpy(h->d_coeff64 + c * N, v.data(), N * sizeof(int64_t), cudaMemcpyHostToDevice));
    if (h->path == PATH_Q15 || h->path == PATH_Q24) {
      std::vector<uint32_t> pk(h->pk_words, 0);
      if (h->path == PATH_Q15) fir_q15_pack(h->fc, v.data(), (int)N, h->d.ftype, pk.data(), h->pk_words);
      else fir_q24_pack(v.data(), (int)N, h->d.ftype, pk.data(), h->pk_words);
      CU(cudaMemcpy(h->d_coeff_pk + (size_t)c * h->pk_words, pk.data(), h->pk_words * sizeof(uint32_t), cudaMemcpyHostToDevice));
    }
    if (h->d_hs && !same) h->hs_stale[c] = 1;
    if (h->path == PATH_WIDE) {
      std::vector<int32_t> w(h->wide_words, 0);
      fir_wide_pack(v.data(), (int)N, h->d.ftype, h->wide_mode, w.data(), h->wide_words);
      CU(cudaMemcpy(h->d_coeff32 + (size_t)c * h->wide_words, w.data(), h->wide_words * sizeof(int32_t), cudaMemcpyHostToDevice));
    }
  }
  if (h->d_hs) {   // error bound of the overlap-save evaluation over the taps now loaded (largest 1-norm of any channel)
    double l1max = 0.0;
    std::vector<int64_t> eff(N);
    for (uint32_t c = 0; c < C; c++) {
      if (!h->ch_loaded[c]) continue;
      fir_effective_taps(h->h_coeff.data() + (size_t)c * N, (int)N, h->d.ftype, eff.data());
      double l1 = 0.0;
      for (size_t i = 0; i < N; i++) l1 += std::fabs((double)eff[i]);
      l1max = std::max(l1max, l1);
    }
    h->ovs_bound = fir_ovs_error_bound(h->fin, l1max);
  }
  return B2D_OK;
}

static int fir_launch(b2d_fir *h, const void *d_in, size_t n, void *d_out, cudaStream_t st) {
  if (n == 0) return B2D_OK;
  FirLaunch p;
  p.fin = h->fin; p.fcoeff = h->fc; p.facc = h->fa; p.fout = h->fo;
  p.n_taps = (int)h->d.n_taps; p.ftype = h->d.ftype; p.C = h->d.n_channels; p.interleaved = h->d.layout == B2D_INTERLEAVED;
  p.ascending = h->d.kind == B2D_FIR_REG_SHARE;
  p.in = d_in; p.out = d_out; p.n = n;
  p.tail = h->d_tail[h->cur]; p.tail_next = h->d_tail[h->cur ^ 1];
  p.coeff64 = h->d_coeff64; p.coeff_pk = h->d_coeff_pk; p.pk_words = h->pk_words; p.coeff32 = h->d_coeff32;
  int hs = hist_wait(h->e_hist, st);
  if (hs) return hs;
  if (h->path == PATH_Q15 && fir_ovs_worth(h, n) && fir_ovs_armed(h) && (hs = fir_ovs_prepare(h, st))) return hs;
  CU(fir_dispatch(h, p, st, true));
  if (h->pend_rem) {   // TRANSPOSED after a coefficient change: the first outputs start from the old taps' partial sums
    const size_t m = std::min(n, h->pend_rem);
    CU(launch_fir_pending(p, m, h->d_pend[h->pcur], nullptr, st));
    CU(launch_fir_pending_shift(h->d_pend[h->pcur], h->d_pend[h->pcur ^ 1], m, h->T, h->d.n_channels, st));
    h->pcur ^= 1;
    h->pend_rem -= m;
  }
  if (h->d_dl) CU(launch_fir_delay_out(p, h->d_dl, st));
  CU(launch_fir_tail(p, st));
  if ((hs = hist_mark(h->e_hist, st))) return hs;
  h->cur ^= 1;
  h->ran = true;
  return B2D_OK;
}

extern "C" int b2d_fir_run_dev(b2d_fir *h, const void *d_in, size_t n, void *d_out, size_t *n_out, void *cuda_stream) {
  TraceRange trace__("b2d_fir_run_dev");
  if (!h || (n && (!d_in || !d_out))) return fail(B2D_EINVAL, "null argument");
  if (!all_loaded(h)) return fail(B2D_ESTATE, "run() before the coefficients of every channel were loaded");
  int st = use_device(h->device);
  if (st) return st;
  if ((st = fir_launch(h, d_in, n, d_out, (cudaStream_t)cuda_stream))) return st;
  if (n_out) *n_out = n;
  return B2D_OK;
}

extern "C" int b2d_fir_run(b2d_fir *h, const void *in, size_t n, void *out, size_t *n_out) {
  TraceRange trace__("b2d_fir_run");
  if (!h || (n && (!in || !out))) return fail(B2D_EINVAL, "null argument");
  if (!all_loaded(h)) return fail(B2D_ESTATE, "run() before the coefficients of every channel were loaded");
  if (n_out) *n_out = n;
  if (n == 0) return B2D_OK;
  int st = use_device(h->device);
  if (st) return st;
  HostRun r;
  r.in = in; r.out = out; r.n = n; r.C = h->d.n_channels; r.il = h->d.layout == B2D_INTERLEAVED;
  r.in_bytes = h->in_bytes; r.out_bytes = h->out_bytes; r.wire_bytes = wire_bytes_of(h->fo.W, h->wire);
  r.out_like_in = true; r.no_total = n;
  r.L = r.Lout = pipe_chunk(n, (double)r.C * (r.in_bytes + r.wire_bytes));
  return run_host_pipeline(h->pipe, r, [](size_t len) { return len; },
                           [h](const void *d_in, size_t len, void *d_out, size_t, cudaStream_t s) { return fir_launch(h, d_in, len, d_out, s); });
}

// Host-link format of b2d_fir_run's output array (b200dsp.h: b2d_wire).
extern "C" int b2d_fir_set_wire(b2d_fir *h, int32_t wire) {
  if (!h) return fail(B2D_EINVAL, "null handle");
  int st = check_wire(wire);
  if (st) return st;
  h->wire = wire;
  return B2D_OK;
}

extern "C" int b2d_fir_reset(b2d_fir *h) {
  if (!h) return fail(B2D_EINVAL, "null handle");
  int st = use_device(h->device);
  if (st) return st;
  CU(cudaDeviceSynchronize());
  const size_t tail_bytes = std::max<size_t>((size_t)h->T * h->d.n_channels * h->in_bytes, 16);
  for (int i = 0; i < 2; i++) CU(cudaMemset(h->d_tail[i], 0, tail_bytes));
  if (h->d_dl) CU(cudaMemset(h->d_dl, 0, h->d.n_channels * sizeof(int64_t)));
  for (int i = 0; i < 2; i++) if (h->d_pend[i]) CU(cudaMemset(h->d_pend[i], 0, (size_t)h->d.n_channels * h->T * sizeof(int64_t)));
  h->pend_rem = 0;
  h->ran = false;
  return B2D_OK;
}

extern "C" int b2d_fir_load_blocked(b2d_fir *h, const void *ram, size_t n_ram, uint32_t mww, uint32_t bs, uint32_t bo, int32_t channel) {
  TraceRange trace__("b2d_fir_load_blocked");
  if (!h) return fail(B2D_EINVAL, "null handle");
  if (bs < 1 || mww < 1) return fail(B2D_EINVAL, "mem_word_width and blk_sz must be >= 1");
  const size_t N = h->d.n_taps;
  const int ft = h->d.ftype;
  const size_t used = (ft == B2D_FOLD_EVEN || ft == B2D_FOLD_EVEN_ANTI) ? N / 2 : ((ft == B2D_FOLD_ODD || ft == B2D_FOLD_ODD_ANTI) ? (N - 1) / 2 + 1 : N);
  if (used % bs) return fail(B2D_EUNSUPPORTED, "tap loop length %zu is not a multiple of blk_sz %u (the reference reads its delay line out of range)", used, bs);
  const size_t need = used ? (used / bs - 1) * (size_t)mww + bo + bs : 0;
  const bool have_local = !h->comm || h->comm->rank == h->root;
  const bool bad = have_local && (!ram || n_ram < need);   // seen by the root only: travels with the broadcast's status word
  std::vector<unsigned char> taps(N * (size_t)h->c_bytes, 0);
  if (have_local && !bad)
    for (size_t t = 0; t < used; t++)
      memcpy(&taps[t * h->c_bytes], (const char *)ram + ((t / bs) * (size_t)mww + bo + t % bs) * h->c_bytes, h->c_bytes);
  const int st = b2d_fir_load(h, (have_local && !bad) ? taps.data() : nullptr, N, channel);
  if (bad) return fail(B2D_EINVAL, "coefficient RAM needs %zu words, got %zu", need, n_ram);
  return st;
}

extern "C" int b2d_fir_delay_line_out(b2d_fir *h, void *out_raw) {
  if (!h || !out_raw) return fail(B2D_EINVAL, "null argument");
  if (!h->d_dl) return fail(B2D_ESTATE, "delay-line output exists for B2D_FIR_REG_SHARE handles only");
  int st = use_device(h->device);
  if (st) return st;
  CU(cudaDeviceSynchronize());
  std::vector<int64_t> v(h->d.n_channels);
  CU(cudaMemcpy(v.data(), h->d_dl, v.size() * sizeof(int64_t), cudaMemcpyDeviceToHost));
  for (size_t c = 0; c < v.size(); c++) {
    if (h->out_bytes == 2) ((int16_t *)out_raw)[c] = (int16_t)v[c];
    else if (h->out_bytes == 4) ((int32_t *)out_raw)[c] = (int32_t)v[c];
    else ((int64_t *)out_raw)[c] = v[c];
  }
  return B2D_OK;
}

extern "C" int b2d_fir_run_window(b2d_fir *h, const void *window, void *out_raw) {
  TraceRange trace__("b2d_fir_run_window");
  if (!h || !window || !out_raw) return fail(B2D_EINVAL, "null argument");
  if (!all_loaded(h)) return fail(B2D_ESTATE, "run() before the coefficients of every channel were loaded");
  int st = use_device(h->device);
  if (st) return st;
  const uint32_t C = h->d.n_channels;
  const size_t N = h->d.n_taps, T = N - 1, ib = h->in_bytes, ob = h->out_bytes;
  const size_t tail_b = (std::max<size_t>(T * C * ib, 16) + 15) & ~(size_t)15, in_b = (C * ib + 15) & ~(size_t)15, out_b = (C * ob + 15) & ~(size_t)15;
  if (!h->d_win) CU(cudaMalloc(&h->d_win, 2 * tail_b + in_b + out_b));
  // reg order (newest first) -> planar tail, oldest first, + the newest sample as the one-sample input
  std::vector<unsigned char> host(tail_b + in_b, 0);
  for (uint32_t c = 0; c < C; c++) {
    const unsigned char *w = (const unsigned char *)window + (size_t)c * N * ib;
    for (size_t j = 0; j < T; j++) memcpy(&host[(c * T + j) * ib], w + (T - j) * ib, ib);   // tail[j] = reg[T - j]
    memcpy(&host[tail_b + c * ib], w, ib);
  }
  char *base = (char *)h->d_win;
  CU(cudaMemcpy(base, host.data(), host.size(), cudaMemcpyHostToDevice));
  FirLaunch p;
  p.fin = h->fin; p.fcoeff = h->fc; p.facc = h->fa; p.fout = h->fo;
  p.n_taps = (int)N; p.ftype = h->d.ftype; p.C = C; p.interleaved = 1;   // one sample per channel: [1][C]
  p.ascending = h->d.kind == B2D_FIR_REG_SHARE;
  p.in = base + tail_b; p.out = base + tail_b + in_b; p.n = 1;
  p.tail = base; p.tail_next = base + tail_b + in_b + out_b;
  p.coeff64 = h->d_coeff64; p.coeff_pk = h->d_coeff_pk; p.pk_words = h->pk_words; p.coeff32 = h->d_coeff32;
  CU(fir_dispatch(h, p, nullptr));
  CU(cudaMemcpy(out_raw, p.out, C * ob, cudaMemcpyDeviceToHost));
  return B2D_OK;
}

// FIR checkpoint: the input history; a TRANSPOSED filter adds the pending partial sums of its last coefficient change
// (zeros when none is pending) and their count in the header.
static size_t fir_pend_bytes(const b2d_fir *h) {
  return h->d.ftype == B2D_TRANSPOSED ? (size_t)h->T * h->d.n_channels * sizeof(int64_t) : 0;
}
extern "C" int b2d_fir_state_bytes(b2d_fir *h, size_t *bytes) {
  if (!h || !bytes) return fail(B2D_EINVAL, "null argument");
  *bytes = sizeof(StateHdr) + (size_t)h->T * h->d.n_channels * h->in_bytes + fir_pend_bytes(h);
  return B2D_OK;
}
extern "C" int b2d_fir_get_state(b2d_fir *h, void *blob, size_t bytes) {
  size_t need = 0;
  if (!h || !blob) return fail(B2D_EINVAL, "null argument");
  b2d_fir_state_bytes(h, &need);
  if (bytes < need) return fail(B2D_EINVAL, "state blob needs %zu bytes", need);
  int st = use_device(h->device);
  if (st) return st;
  CU(cudaDeviceSynchronize());
  StateHdr hd{kFirMagic, 1, (uint64_t)h->pend_rem, (uint32_t)h->T, h->d.n_channels, (uint32_t)h->in_bytes, 0};
  memcpy(blob, &hd, sizeof(hd));
  const size_t tail_b = (size_t)h->T * h->d.n_channels * h->in_bytes, pend_b = fir_pend_bytes(h);
  if (tail_b) CU(cudaMemcpy((char *)blob + sizeof(hd), h->d_tail[h->cur], tail_b, cudaMemcpyDeviceToHost));
  if (pend_b) {
    if (h->pend_rem) CU(cudaMemcpy((char *)blob + sizeof(hd) + tail_b, h->d_pend[h->pcur], pend_b, cudaMemcpyDeviceToHost));
    else memset((char *)blob + sizeof(hd) + tail_b, 0, pend_b);
  }
  return B2D_OK;
}
extern "C" int b2d_fir_set_state(b2d_fir *h, const void *blob, size_t bytes) {
  size_t need = 0;
  if (!h || !blob) return fail(B2D_EINVAL, "null argument");
  b2d_fir_state_bytes(h, &need);
  StateHdr hd;
  if (bytes < need) return fail(B2D_EINVAL, "state blob needs %zu bytes", need);
  memcpy(&hd, blob, sizeof(hd));
  if (hd.magic != kFirMagic || hd.hist != (uint32_t)h->T || hd.channels != h->d.n_channels || hd.bytes != (uint32_t)h->in_bytes)
    return fail(B2D_EINVAL, "state blob does not belong to this filter configuration");
  int st = use_device(h->device);
  if (st) return st;
  CU(cudaDeviceSynchronize());
  const size_t tail_b = (size_t)h->T * h->d.n_channels * h->in_bytes, pend_b = fir_pend_bytes(h);
  if (hd.n_seen > (uint64_t)h->T) return fail(B2D_EINVAL, "state blob: pending count out of range");
  if (tail_b) CU(cudaMemcpy(h->d_tail[h->cur], (const char *)blob + sizeof(hd), tail_b, cudaMemcpyHostToDevice));
  if (pend_b && hd.n_seen) {
    for (int i = 0; i < 2; i++)
      if (!h->d_pend[i]) CU(cudaMalloc(&h->d_pend[i], pend_b));
    CU(cudaMemcpy(h->d_pend[h->pcur], (const char *)blob + sizeof(hd) + tail_b, pend_b, cudaMemcpyHostToDevice));
  }
  h->pend_rem = pend_b ? (size_t)hd.n_seen : 0;
  h->ran = true;                  // a restored history stands for consumed samples
  return B2D_OK;
}
