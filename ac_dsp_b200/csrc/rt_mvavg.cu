// rt_mvavg.cu -- host runtime of ac_mv_avg (reference include/ac_dsp/ac_mv_avg.h:140-204).
#include "rt_common.h"

using namespace b2d;

struct b2d_mvavg {
  b2d_mvavg_desc d;
  Fmt fin, fc, fa, fo;
  int device = 0, in_bytes = 2, out_bytes = 2;
  int64_t *d_coeff = nullptr;
  void *d_in = nullptr, *d_out = nullptr;
  size_t cap_in = 0, cap_out = 0;
};

extern "C" int b2d_mvavg_destroy(b2d_mvavg *h) {
  if (!h) return B2D_OK;
  use_device(h->device);
  cudaDeviceSynchronize();
  if (h->d_coeff) cudaFree(h->d_coeff);
  if (h->d_in) cudaFree(h->d_in);
  if (h->d_out) cudaFree(h->d_out);
  delete h;
  return B2D_OK;
}

// Class instantiation + constructor: the weights are a const array handed to the constructor (ac_mv_avg.h:149).
extern "C" int b2d_mvavg_create(b2d_mvavg **out, const b2d_mvavg_desc *desc, const void *coeff_raw) {
  if (!out || !desc || !coeff_raw) return fail(B2D_EINVAL, "null argument");
  *out = nullptr;
  int st;
  if ((st = check_fmt(desc->in, 32, "IN_TYPE"))) return st;
  if ((st = check_fmt(desc->coeff, 32, "COEFF_TYPE"))) return st;
  if ((st = check_fmt(desc->acc, 64, "ACC_TYPE"))) return st;
  if ((st = check_fmt(desc->out, 64, "OUT_TYPE"))) return st;
  if (desc->taps < 1 || !(desc->taps & 1) || desc->taps > 65535) return fail(B2D_EINVAL, "TAPS = %u: the window span must be odd (1..65535)", desc->taps);
  if (desc->max_sample < desc->taps) return fail(B2D_EINVAL, "MAX_SAMPLE = %u below the window span", desc->max_sample);
  if (desc->win_type != B2D_WIN && desc->win_type != B2D_CLIP && desc->win_type != B2D_MIRROR) return fail(B2D_EINVAL, "bad window type");
  const Fmt fin = to_fmt(desc->in), fc = to_fmt(desc->coeff), fa = to_fmt(desc->acc), fo = to_fmt(desc->out);
  {  // bit budget of the 128-bit evaluation: ACC_TYPE cast of the sample, ACC x COEFF product, sum, output conversion
    const int Fp = fa.F() + fc.F(), Wp = fa.W + fc.W + 2;
    if (Wp > 125 || fa.W + std::max(0, Fp - fa.F()) > 125 || fin.W + std::max(0, fa.F() - fin.F()) > 125 || fa.W + std::max(0, fo.F() - fa.F()) > 125)
      return fail(B2D_EUNSUPPORTED, "format combination exceeds the 128-bit intermediate budget");
  }
  int dev = desc->device;
  if (dev < 0) CU(cudaGetDevice(&dev));
  if ((st = use_device(dev))) return st;
  b2d_mvavg *h = new (std::nothrow) b2d_mvavg();
  if (!h) return fail(B2D_ENOMEM, "handle");
  h->d = *desc; h->fin = fin; h->fc = fc; h->fa = fa; h->fo = fo; h->device = dev;
  h->in_bytes = container_bytes(fin.W); h->out_bytes = container_bytes(fo.W);
  std::vector<int64_t> v(desc->taps);
  widen_coeffs(coeff_raw, desc->taps, container_bytes(fc.W), fc, v.data());
  cudaError_t e = cudaMalloc(&h->d_coeff, v.size() * sizeof(int64_t));
  if (e == cudaSuccess) e = cudaMemcpy(h->d_coeff, v.data(), v.size() * sizeof(int64_t), cudaMemcpyHostToDevice);
  if (e != cudaSuccess) { cudaGetLastError(); b2d_mvavg_destroy(h); return fail(B2D_ECUDA, "b2d_mvavg_create: %s", cudaGetErrorString(e)); }
  *out = h;
  return B2D_OK;
}

extern "C" const char *b2d_mvavg_path(b2d_mvavg *h) { return h ? "mvavg_generic" : ""; }

static size_t mvavg_per_burst(const b2d_mvavg *h, size_t n_sample) {
  return h->d.win_type == B2D_WIN ? n_sample - h->d.taps + 1 : n_sample;
}
extern "C" size_t b2d_mvavg_max_out(b2d_mvavg *h, size_t n) { return h ? n : 0; }

static int mvavg_check(const b2d_mvavg *h, size_t n_in, size_t n_sample) {
  // the manual's limitation (section 2.4.2): the window span must not exceed the number of samples of a burst; the
  // reference loop handles at most MAX_SAMPLE samples per burst and reads whole bursts off its input channel
  if (n_sample < h->d.taps || n_sample > h->d.max_sample)
    return fail(B2D_EINVAL, "n_sample = %zu outside TAPS .. MAX_SAMPLE (%u .. %u)", n_sample, h->d.taps, h->d.max_sample);
  if (n_in % n_sample) return fail(B2D_EINVAL, "%zu samples are not a whole number of bursts of %zu", n_in, n_sample);
  return B2D_OK;
}

extern "C" int b2d_mvavg_run_dev(b2d_mvavg *h, const void *d_in, size_t n_in, size_t n_sample, void *d_out, size_t *n_out, void *cuda_stream) {
  TraceRange trace__("b2d_mvavg_run_dev");
  if (!h || (n_in && (!d_in || !d_out))) return fail(B2D_EINVAL, "null argument");
  int st = mvavg_check(h, n_in, n_sample);
  if (st) return st;
  if ((st = use_device(h->device))) return st;
  MvLaunch p;
  p.fin = h->fin; p.fcoeff = h->fc; p.facc = h->fa; p.fout = h->fo; p.taps = (int)h->d.taps; p.win = h->d.win_type;
  p.in = d_in; p.out = d_out; p.coeff64 = h->d_coeff; p.n_sample = n_sample; p.per = mvavg_per_burst(h, n_sample);
  p.n_out = (n_in / n_sample) * p.per;
  CU(launch_mvavg(p, (cudaStream_t)cuda_stream));
  if (n_out) *n_out = p.n_out;
  return B2D_OK;
}

extern "C" int b2d_mvavg_run(b2d_mvavg *h, const void *in, size_t n_in, size_t n_sample, void *out, size_t *n_out) {
  TraceRange trace__("b2d_mvavg_run");
  if (!h || (n_in && (!in || !out))) return fail(B2D_EINVAL, "null argument");
  int st = mvavg_check(h, n_in, n_sample);
  if (st) return st;
  if ((st = use_device(h->device))) return st;
  const size_t in_b = n_in * h->in_bytes, out_b = n_in * h->out_bytes;
  if (in_b > h->cap_in) { if (h->d_in) cudaFree(h->d_in); h->d_in = nullptr; h->cap_in = 0; CU(cudaMalloc(&h->d_in, in_b)); h->cap_in = in_b; }
  if (out_b > h->cap_out) { if (h->d_out) cudaFree(h->d_out); h->d_out = nullptr; h->cap_out = 0; CU(cudaMalloc(&h->d_out, out_b)); h->cap_out = out_b; }
  if (in_b) CU(cudaMemcpy(h->d_in, in, in_b, cudaMemcpyHostToDevice));
  size_t no = 0;
  if ((st = b2d_mvavg_run_dev(h, h->d_in, n_in, n_sample, h->d_out, &no, nullptr))) return st;
  CU(cudaDeviceSynchronize());
  if (no) CU(cudaMemcpy(out, h->d_out, no * h->out_bytes, cudaMemcpyDeviceToHost));
  if (n_out) *n_out = no;
  return B2D_OK;
}
