// rt_intgdump.cu -- host runtime of ac_intg_dump.
#include "rt_common.h"

using namespace b2d;

// -------------------------------------------------------------------------------------------- ac_intg_dump
struct b2d_intgdump {
  b2d_intgdump_desc d;
  Fmt fin, fa, fo;
  int device = 0, in_bytes = 2, out_bytes = 8;
  int64_t *d_carry[2] = {nullptr, nullptr};
  int cur = 0;
  unsigned long long *d_table = nullptr;
  size_t table_cap = 0;
  void *d_in = nullptr, *d_out = nullptr;
  size_t cap_in = 0, cap_out = 0;
  const char *path = "none";
};

extern "C" const char *b2d_intgdump_path(b2d_intgdump *h) { return h ? h->path : "none"; }

extern "C" int b2d_intgdump_destroy(b2d_intgdump *h) {
  if (!h) return B2D_OK;
  use_device(h->device);
  cudaDeviceSynchronize();
  for (int i = 0; i < 2; i++) if (h->d_carry[i]) cudaFree(h->d_carry[i]);
  if (h->d_table) cudaFree(h->d_table);
  if (h->d_in) cudaFree(h->d_in);
  if (h->d_out) cudaFree(h->d_out);
  delete h;
  return B2D_OK;
}

extern "C" int b2d_intgdump_create(b2d_intgdump **out, const b2d_intgdump_desc *desc) {
  if (!out || !desc) return fail(B2D_EINVAL, "null argument");
  *out = nullptr;
  int st;
  if ((st = check_fmt(desc->in, 32, "IN_TYPE"))) return st;
  if ((st = check_fmt(desc->acc, 64, "ACC_TYPE"))) return st;
  if ((st = check_fmt(desc->out, 64, "OUT_TYPE"))) return st;
  if (desc->ns < 1 || desc->chn < 1 || desc->chn > 65536) return fail(B2D_EINVAL, "NS = %u, CHN = %u invalid", desc->ns, desc->chn);
  const Fmt fin = to_fmt(desc->in), fa = to_fmt(desc->acc), fo = to_fmt(desc->out);
  if (std::abs(fin.F() - fa.F()) > 60 || fa.W + std::max(0, fo.F() - fa.F()) > 125) return fail(B2D_EUNSUPPORTED, "format combination exceeds the intermediate budget");
  int dev = desc->device;
  if (dev < 0) CU(cudaGetDevice(&dev));
  if ((st = use_device(dev))) return st;
  b2d_intgdump *h = new (std::nothrow) b2d_intgdump();
  if (!h) return fail(B2D_ENOMEM, "handle");
  h->d = *desc; h->fin = fin; h->fa = fa; h->fo = fo; h->device = dev;
  h->in_bytes = container_bytes(fin.W); h->out_bytes = container_bytes(fo.W);
  for (int i = 0; i < 2; i++) {
    cudaError_t e = cudaMalloc(&h->d_carry[i], desc->chn * sizeof(int64_t));
    if (e == cudaSuccess) e = cudaMemset(h->d_carry[i], 0, desc->chn * sizeof(int64_t));
    if (e != cudaSuccess) { cudaGetLastError(); b2d_intgdump_destroy(h); return fail(B2D_ECUDA, "b2d_intgdump_create: %s", cudaGetErrorString(e)); }
  }
  *out = h;
  return B2D_OK;
}

extern "C" int b2d_intgdump_reset(b2d_intgdump *h) {
  if (!h) return fail(B2D_EINVAL, "null handle");
  int st = use_device(h->device);
  if (st) return st;
  CU(cudaDeviceSynchronize());
  for (int i = 0; i < 2; i++) CU(cudaMemset(h->d_carry[i], 0, h->d.chn * sizeof(int64_t)));
  return B2D_OK;
}

// the running sums temp[CHN] (ac_intg_dump.h:78)
extern "C" int b2d_intgdump_state_bytes(b2d_intgdump *h, size_t *bytes) {
  if (!h || !bytes) return fail(B2D_EINVAL, "null argument");
  *bytes = sizeof(StateHdr) + (size_t)h->d.chn * sizeof(int64_t);
  return B2D_OK;
}
extern "C" int b2d_intgdump_get_state(b2d_intgdump *h, void *blob, size_t bytes) {
  if (!h) return fail(B2D_EINVAL, "null handle");
  int st = use_device(h->device);
  if (st) return st;
  const StatePart parts[1] = {{h->d_carry[h->cur], (size_t)h->d.chn * sizeof(int64_t)}};
  return state_get(StateHdr{kDumpMagic, 1, 0, 0, h->d.chn, 8, 0}, parts, 1, blob, bytes);
}
extern "C" int b2d_intgdump_set_state(b2d_intgdump *h, const void *blob, size_t bytes) {
  if (!h) return fail(B2D_EINVAL, "null handle");
  int st = use_device(h->device);
  if (st) return st;
  const StatePart parts[1] = {{h->d_carry[h->cur], (size_t)h->d.chn * sizeof(int64_t)}};
  StateHdr got;
  return state_set(StateHdr{kDumpMagic, 1, 0, 0, h->d.chn, 8, 0}, parts, 1, blob, bytes, &got);
}

// token sequence -> segments of the per-channel sample axis (control flow of ac_intg_dump.h:133-147)
static int intgdump_plan(const b2d_intgdump *h, const uint32_t *n_sample, size_t n_frames, size_t n_in, std::vector<unsigned long long> &bounds,
                         size_t *nseg_out, int *has_tail, bool *regular) {
  const unsigned long long NS = h->d.ns, CHN = h->d.chn;
  bounds.clear();
  unsigned long long pos = 0;
  // equal dumping frames (the common case) need no boundary table: one vectorisable pass over the tokens
  uint32_t diff = 0;
  const uint32_t n0 = n_frames ? n_sample[0] : 0;
  for (size_t f = 0; f < n_frames; f++) diff |= n_sample[f] ^ n0;
  *regular = n_frames > 0 && diff == 0 && n0 >= 1 && n0 <= NS;
  if (*regular) {
    pos = (unsigned long long)n0 * n_frames;
    if (pos * CHN != n_in) return fail(B2D_EINVAL, "the frames consume %llu samples, got %zu", pos * CHN, n_in);
    *nseg_out = n_frames;
    *has_tail = 0;
    return B2D_OK;
  }
  bounds.push_back(0);
  for (size_t f = 0; f < n_frames; f++) {
    const unsigned long long n = n_sample[f];
    if (n >= 1 && n <= NS) { pos += n; bounds.push_back(pos); }
    else pos += NS;
  }
  if (pos * CHN != n_in) return fail(B2D_EINVAL, "the frames consume %llu samples, got %zu", pos * CHN, n_in);
  *nseg_out = bounds.size() - 1;
  *has_tail = pos > bounds.back() ? 1 : 0;
  if (*has_tail) bounds.push_back(pos);
  return B2D_OK;
}

extern "C" int b2d_intgdump_run_dev(b2d_intgdump *h, const void *d_in, size_t n_in, const uint32_t *n_sample, size_t n_frames, void *d_out,
                                    size_t *n_out, void *cuda_stream) {
  TraceRange trace__("b2d_intgdump_run_dev");
  if (!h || (n_frames && !n_sample) || (n_in && !d_in)) return fail(B2D_EINVAL, "null argument");
  int st = use_device(h->device);
  if (st) return st;
  std::vector<unsigned long long> bounds;
  size_t nseg_out = 0;
  int has_tail = 0;
  bool regular = false;
  if ((st = intgdump_plan(h, n_sample, n_frames, n_in, bounds, &nseg_out, &has_tail, &regular))) return st;
  if (nseg_out && !d_out) return fail(B2D_EINVAL, "null output");
  if (n_out) *n_out = nseg_out * h->d.chn;
  if (nseg_out == 0 && !has_tail) return B2D_OK;
  cudaStream_t stream = (cudaStream_t)cuda_stream;
  IdLaunch p;
  p.fin = h->fin; p.facc = h->fa; p.fout = h->fo; p.chn = (int)h->d.chn;
  { const char *f = getenv("B2D_FORCE_GENERIC"); p.force_thread = (f && *f == '1') ? 1 : 0; }
  p.in = d_in; p.out = d_out; p.carry = h->d_carry[h->cur]; p.carry_next = h->d_carry[h->cur ^ 1];
  p.nseg_out = nseg_out; p.has_tail = has_tail; p.table = nullptr; p.n_reg = 0; p.tail_end = 0;
  if (regular) p.n_reg = n_sample[0];
  else {
    if (bounds.size() > h->table_cap) {
      if (h->d_table) cudaFree(h->d_table);
      h->d_table = nullptr; h->table_cap = 0;
      CU(cudaMalloc(&h->d_table, bounds.size() * sizeof(unsigned long long)));
      h->table_cap = bounds.size();
    }
    CU(cudaMemcpyAsync(h->d_table, bounds.data(), bounds.size() * sizeof(unsigned long long), cudaMemcpyHostToDevice, stream));
    CU(cudaStreamSynchronize(stream));          // `bounds` is a local
    p.table = h->d_table;
  }
  // a call without a tail segment leaves temp[] cleared (the last dump zeroed it)
  if (!has_tail) CU(cudaMemsetAsync(h->d_carry[h->cur ^ 1], 0, h->d.chn * sizeof(int64_t), stream));
  h->path = intgdump_path(p);
  CU(launch_intgdump(p, stream));
  h->cur ^= 1;
  return B2D_OK;
}

extern "C" int b2d_intgdump_run(b2d_intgdump *h, const void *in, size_t n_in, const uint32_t *n_sample, size_t n_frames, void *out, size_t *n_out) {
  TraceRange trace__("b2d_intgdump_run");
  if (!h || (n_frames && !n_sample) || (n_in && !in)) return fail(B2D_EINVAL, "null argument");
  int st = use_device(h->device);
  if (st) return st;
  size_t n_dump = 0;
  for (size_t f = 0; f < n_frames; f++) n_dump += (n_sample[f] >= 1 && n_sample[f] <= h->d.ns) ? 1 : 0;
  if (n_dump && !out) return fail(B2D_EINVAL, "null output");
  const size_t in_b = n_in * h->in_bytes, out_b = n_dump * h->d.chn * h->out_bytes;
  if (in_b > h->cap_in) { if (h->d_in) cudaFree(h->d_in); h->d_in = nullptr; h->cap_in = 0; CU(cudaMalloc(&h->d_in, in_b)); h->cap_in = in_b; }
  if (out_b > h->cap_out) { if (h->d_out) cudaFree(h->d_out); h->d_out = nullptr; h->cap_out = 0; CU(cudaMalloc(&h->d_out, out_b)); h->cap_out = out_b; }
  if (in_b) CU(cudaMemcpy(h->d_in, in, in_b, cudaMemcpyHostToDevice));
  size_t no = 0;
  if ((st = b2d_intgdump_run_dev(h, h->d_in, n_in, n_sample, n_frames, h->d_out, &no, nullptr))) return st;
  CU(cudaDeviceSynchronize());
  if (no) CU(cudaMemcpy(out, h->d_out, no * h->out_bytes, cudaMemcpyDeviceToHost));
  if (n_out) *n_out = no;
  return B2D_OK;
}
