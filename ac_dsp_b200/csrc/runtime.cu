// runtime.cu -- the part of the host runtime behind include/b200dsp.h that every handle family shares: error text,
// page-locked host buffers, the device slots and streams of the pipelined host-buffer path (the loop itself is the
// template in rt_common.h), the NCCL communicator and checkpoint blobs.  The handle families live in rt_fir.cu, rt_cic.cu,
// rt_poly.cu, rt_intgdump.cu and rt_mvavg.cu.  No CPU compute path exists: every run() ends in a CUDA kernel launch.
#include "rt_common.h"

using namespace b2d;

// ------------------------------------------------------------------------------------------ errors
static thread_local char g_err[512] = "";
int b2d::fail(int status, const char *fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return status;
}
extern "C" const char *b2d_version(void) { return "b200dsp 0.1 (sm_100a)"; }
extern "C" const char *b2d_last_error(void) { return g_err; }
extern "C" const char *b2d_strerror(int s) {
  switch (s) {
    case B2D_OK: return "ok";
    case B2D_EUNSUPPORTED: return "configuration not supported by the CUDA engine";
    case B2D_EINVAL: return "invalid argument";
    case B2D_ECUDA: return "CUDA error";
    case B2D_ENCCL: return "NCCL error";
    case B2D_ENOMEM: return "out of memory";
    case B2D_ESTATE: return "call sequence error";
    default: return "unknown status";
  }
}
extern "C" int b2d_container_bytes(int32_t W) { return (W < 1 || W > 64) ? 0 : container_bytes(W); }
extern "C" int b2d_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
  return n;
}
// Page-locked host buffers (cudaHostAlloc, portable across devices).  Round 1 prepared a variant that took the pages from
// the GPU's own NUMA node (mmap + mbind + cudaHostRegister); the 8-GPU A/B of round 2 showed no difference on this class of
// box (one NUMA node, every GPU behind the same host bridge: profiles/r02_bench_numa1_n8.json vs r02_bench_default_n8.json,
// 5313 vs 5307 M IQ samples/s), so it was removed.
extern "C" int b2d_host_alloc(void **p, size_t bytes) {
  if (!p) return fail(B2D_EINVAL, "null pointer");
  CU(cudaHostAlloc(p, bytes ? bytes : 1, cudaHostAllocPortable));
  return B2D_OK;
}
extern "C" int b2d_host_free(void *p) {
  if (!p) return B2D_OK;
  CU(cudaFreeHost(p));
  return B2D_OK;
}

int b2d::check_fmt(const b2d_fmt &f, int maxW, const char *what) {
  if (f.W < 1 || f.W > 64) return fail(B2D_EINVAL, "%s: width %d outside 1..64", what, f.W);
  if (f.W > maxW) return fail(B2D_EUNSUPPORTED, "%s: width %d > %d", what, f.W, maxW);
  if (f.I < -64 || f.I > 128) return fail(B2D_EUNSUPPORTED, "%s: integer width %d outside -64..128", what, f.I);
  if (f.Q < B2D_TRN || f.Q > B2D_RND_CONV_ODD) return fail(B2D_EINVAL, "%s: bad quantisation mode %d", what, f.Q);
  if (f.O < B2D_WRAP || f.O > B2D_SAT_SYM) return fail(B2D_EINVAL, "%s: bad overflow mode %d", what, f.O);
  return B2D_OK;
}
int b2d::use_device(int dev) {
  int cur = -1;
  CU(cudaGetDevice(&cur));
  if (cur != dev) CU(cudaSetDevice(dev));
  return B2D_OK;
}

// ------------------------------------------------------------------------------------ host pipeline
// run() on HOST buffers: the stream is cut into chunks; chunk i+1 is copied in while chunk i computes
// and chunk i-1 is copied out (three streams, three device slots).  Each chunk is an ordinary run_dev()
// call, so the result is the reference's own "several run() calls" behaviour by construction.
int Pipe::init() {
  if (ready) return B2D_OK;
  CU(cudaStreamCreateWithFlags(&s_in, cudaStreamNonBlocking));
  CU(cudaStreamCreateWithFlags(&s_k, cudaStreamNonBlocking));
  CU(cudaStreamCreateWithFlags(&s_out, cudaStreamNonBlocking));
  for (int i = 0; i < S; i++) {
    CU(cudaEventCreateWithFlags(&e_in[i], cudaEventDisableTiming));
    CU(cudaEventCreateWithFlags(&e_k[i], cudaEventDisableTiming));
    CU(cudaEventCreateWithFlags(&e_out[i], cudaEventDisableTiming));
  }
  ready = true;
  return B2D_OK;
}
static int grow_slots(void **slots, size_t *cap, size_t bytes) {
  if (bytes <= *cap) return B2D_OK;
  for (int i = 0; i < Pipe::S; i++) { if (slots[i]) cudaFree(slots[i]); slots[i] = nullptr; }
  *cap = 0;
  for (int i = 0; i < Pipe::S; i++)
    if (cudaMalloc(&slots[i], bytes) != cudaSuccess) { cudaGetLastError(); return fail(B2D_ENOMEM, "cudaMalloc(%zu)", bytes); }
  *cap = bytes;
  return B2D_OK;
}
int Pipe::ensure(size_t in_bytes, size_t out_bytes, size_t pk_bytes) {
  int st;
  if ((st = grow_slots(d_in, &cap_in, in_bytes))) return st;
  if ((st = grow_slots(d_out, &cap_out, out_bytes))) return st;
  return grow_slots(d_pk, &cap_pk, pk_bytes);
}
void Pipe::destroy() {
  for (int i = 0; i < S; i++) {
    if (d_in[i]) cudaFree(d_in[i]);
    if (d_out[i]) cudaFree(d_out[i]);
    if (d_pk[i]) cudaFree(d_pk[i]);
    if (e_in[i]) cudaEventDestroy(e_in[i]);
    if (e_k[i]) cudaEventDestroy(e_k[i]);
    if (e_out[i]) cudaEventDestroy(e_out[i]);
  }
  if (s_in) cudaStreamDestroy(s_in);
  if (s_k) cudaStreamDestroy(s_k);
  if (s_out) cudaStreamDestroy(s_out);
}

// copy `len` samples per channel starting at time `off` between a full buffer (n_full per channel) and a
// compact chunk buffer (len per channel)
cudaError_t b2d::copy_chunk(void *dst, const void *src, bool to_device, int bytes, uint32_t C, int interleaved,
                              size_t n_full, size_t off, size_t len, cudaStream_t st) {
  if (len == 0) return cudaSuccess;
  const cudaMemcpyKind kind = to_device ? cudaMemcpyHostToDevice : cudaMemcpyDeviceToHost;
  if (interleaved || C == 1) {
    const size_t o = off * C * bytes, sz = len * C * bytes;
    return to_device ? cudaMemcpyAsync(dst, (const char *)src + o, sz, kind, st) : cudaMemcpyAsync((char *)dst + o, src, sz, kind, st);
  }
  if (to_device) return cudaMemcpy2DAsync(dst, len * bytes, (const char *)src + off * bytes, n_full * bytes, len * bytes, C, kind, st);
  return cudaMemcpy2DAsync((char *)dst + off * bytes, n_full * bytes, src, len * bytes, len * bytes, C, kind, st);
}

// ----------------------------------------------------------------------------------------------- comm
static int nccl_fail(const NcclApi *api, int rc, const char *what) {
  return fail(B2D_ENCCL, "%s: %s", what, api->GetErrorString(rc));
}

extern "C" int b2d_shard_count(uint32_t n_channels, int32_t rank, int32_t world, uint32_t *n_local) {
  if (!n_local || world < 1 || rank < 0 || rank >= world) return fail(B2D_EINVAL, "bad rank/world");
  *n_local = n_channels / world + ((uint32_t)rank < n_channels % world ? 1u : 0u);  // channel c -> rank c % world
  return B2D_OK;
}
extern "C" int b2d_comm_unique_id(void *id128) {
  if (!id128) return fail(B2D_EINVAL, "null id");
  const char *why = "";
  const NcclApi *api = nccl_api(&why);
  if (!api) return fail(B2D_ENCCL, "%s", why);
  ncclUniqueId id;
  int rc = api->GetUniqueId(&id);
  if (rc != ncclSuccess) return nccl_fail(api, rc, "ncclGetUniqueId");
  memcpy(id128, &id, sizeof(id));
  return B2D_OK;
}
extern "C" int b2d_comm_create(b2d_comm **c, const void *id128, int32_t rank, int32_t world, int32_t device) {
  if (!c || !id128 || world < 1 || rank < 0 || rank >= world) return fail(B2D_EINVAL, "bad arguments");
  const char *why = "";
  const NcclApi *api = nccl_api(&why);
  if (!api) return fail(B2D_ENCCL, "%s", why);
  if (device < 0) CU(cudaGetDevice(&device));
  int st = use_device(device);
  if (st) return st;
  b2d_comm *k = new (std::nothrow) b2d_comm();
  if (!k) return fail(B2D_ENOMEM, "comm");
  k->rank = rank; k->world = world; k->device = device;
  ncclUniqueId id;
  memcpy(&id, id128, sizeof(id));
  int rc = api->CommInitRank(&k->comm, world, id, rank);
  if (rc != ncclSuccess) { delete k; return nccl_fail(api, rc, "ncclCommInitRank"); }
  cudaError_t e = cudaStreamCreateWithFlags(&k->stream, cudaStreamNonBlocking);
  if (e != cudaSuccess) { api->CommDestroy(k->comm); delete k; return fail(B2D_ECUDA, "cudaStreamCreate: %s", cudaGetErrorString(e)); }
  *c = k;
  return B2D_OK;
}
extern "C" int b2d_comm_destroy(b2d_comm *c) {
  if (!c) return B2D_OK;
  const NcclApi *api = nccl_api(nullptr);
  use_device(c->device);
  if (c->stream) cudaStreamSynchronize(c->stream);
  if (api && c->comm) api->CommDestroy(c->comm);
  if (c->d_buf) cudaFree(c->d_buf);
  if (c->stream) cudaStreamDestroy(c->stream);
  delete c;
  return B2D_OK;
}
static int comm_buf(b2d_comm *c, size_t bytes) {
  if (bytes <= c->cap) return B2D_OK;
  if (c->d_buf) cudaFree(c->d_buf);
  c->d_buf = nullptr; c->cap = 0;
  CU(cudaMalloc(&c->d_buf, bytes));
  c->cap = bytes;
  return B2D_OK;
}
extern "C" int b2d_comm_barrier(b2d_comm *c) {
  if (!c) return fail(B2D_EINVAL, "null comm");
  const NcclApi *api = nccl_api(nullptr);
  int st = use_device(c->device);
  if (st) return st;
  if ((st = comm_buf(c, 8))) return st;
  CU(cudaMemsetAsync(c->d_buf, 0, 8, c->stream));
  int rc = api->AllReduce(c->d_buf, c->d_buf, 1, ncclInt32, ncclSum, c->comm, c->stream);
  if (rc != ncclSuccess) return nccl_fail(api, rc, "ncclAllReduce");
  CU(cudaStreamSynchronize(c->stream));
  return B2D_OK;
}
// values[0..n) of rank `root` -> every rank (int64 payload), synchronous.
int b2d::comm_bcast_i64(b2d_comm *c, int64_t *values, size_t n, int root) {
  const NcclApi *api = nccl_api(nullptr);
  int st = use_device(c->device);
  if (st) return st;
  if ((st = comm_buf(c, n * 8))) return st;
  if (c->rank == root) CU(cudaMemcpyAsync(c->d_buf, values, n * 8, cudaMemcpyHostToDevice, c->stream));
  int rc = api->Broadcast(c->d_buf, c->d_buf, n, ncclInt64, root, c->comm, c->stream);
  if (rc != ncclSuccess) return nccl_fail(api, rc, "ncclBroadcast");
  CU(cudaMemcpyAsync(values, c->d_buf, n * 8, cudaMemcpyDeviceToHost, c->stream));
  CU(cudaStreamSynchronize(c->stream));
  return B2D_OK;
}

// ---------------------------------------------------------------------------------------- checkpoints
size_t b2d::state_total(const StatePart *parts, int np) {
  size_t t = sizeof(StateHdr);
  for (int i = 0; i < np; i++) t += parts[i].bytes;
  return t;
}
int b2d::state_get(const StateHdr &hd, const StatePart *parts, int np, void *blob, size_t bytes) {
  if (!blob) return fail(B2D_EINVAL, "null argument");
  if (bytes < state_total(parts, np)) return fail(B2D_EINVAL, "state blob needs %zu bytes", state_total(parts, np));
  CU(cudaDeviceSynchronize());
  memcpy(blob, &hd, sizeof(hd));
  size_t o = sizeof(hd);
  for (int i = 0; i < np; i++) { if (parts[i].bytes) CU(cudaMemcpy((char *)blob + o, parts[i].dev, parts[i].bytes, cudaMemcpyDeviceToHost)); o += parts[i].bytes; }
  return B2D_OK;
}
int b2d::state_set(const StateHdr &want, const StatePart *parts, int np, const void *blob, size_t bytes, StateHdr *got) {
  if (!blob) return fail(B2D_EINVAL, "null argument");
  if (bytes < state_total(parts, np)) return fail(B2D_EINVAL, "state blob needs %zu bytes", state_total(parts, np));
  memcpy(got, blob, sizeof(*got));
  if (got->magic != want.magic || got->hist != want.hist || got->channels != want.channels || got->bytes != want.bytes)
    return fail(B2D_EINVAL, "state blob does not belong to this filter configuration");
  CU(cudaDeviceSynchronize());
  size_t o = sizeof(*got);
  for (int i = 0; i < np; i++) { if (parts[i].bytes) CU(cudaMemcpy(parts[i].dev, (const char *)blob + o, parts[i].bytes, cudaMemcpyHostToDevice)); o += parts[i].bytes; }
  return B2D_OK;
}

// ------------------------------------------------------------------------------------ shared helpers
void b2d::widen_coeffs(const void *raw, size_t n, int c_bytes, const Fmt &fc, int64_t *out) {
  for (size_t i = 0; i < n; i++) {
    int64_t r;
    if (c_bytes == 2) r = fc.S ? (int64_t)((const int16_t *)raw)[i] : (int64_t)((const uint16_t *)raw)[i];
    else if (c_bytes == 4) r = fc.S ? (int64_t)((const int32_t *)raw)[i] : (int64_t)((const uint32_t *)raw)[i];
    else r = ((const int64_t *)raw)[i];
    out[i] = wrap_bits(r, fc.W, fc.S);
  }
}

int b2d::hist_wait(cudaEvent_t &ev, cudaStream_t st) {
  if (ev) CU(cudaStreamWaitEvent(st, ev, 0));
  return B2D_OK;
}
int b2d::hist_mark(cudaEvent_t &ev, cudaStream_t st) {
  if (!ev) CU(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
  CU(cudaEventRecord(ev, st));
  return B2D_OK;
}

// ------------------------------------------------------------------------------------------ wire format
int b2d::check_wire(int32_t wire) {
  if (wire != B2D_WIRE_CONTAINER && wire != B2D_WIRE_PACKED) return fail(B2D_EINVAL, "bad wire format %d", wire);
  return B2D_OK;
}
extern "C" int b2d_wire_bytes(int32_t W, int32_t wire) {
  if (W < 1 || W > 64 || (wire != B2D_WIRE_CONTAINER && wire != B2D_WIRE_PACKED)) return 0;
  return wire_bytes_of(W, wire);
}
// Packed values (ceil(W/8) little-endian bytes each) -> sign- / zero-extended containers, on the host.
extern "C" int b2d_unpack_wire(const void *packed, size_t count, int32_t W, int32_t S, void *out) {
  if (W < 1 || W > 64) return fail(B2D_EINVAL, "width %d outside 1..64", W);
  if (count && (!packed || !out)) return fail(B2D_EINVAL, "null argument");
  const int pb = (W + 7) / 8, cb = container_bytes(W);
  const unsigned char *p = (const unsigned char *)packed;
  for (size_t i = 0; i < count; i++, p += pb) {
    uint64_t u = 0;
    for (int b = 0; b < pb; b++) u |= (uint64_t)p[b] << (8 * b);
    const int64_t v = wrap_bits((int64_t)u, W, S ? 1 : 0);
    if (cb == 2) ((int16_t *)out)[i] = (int16_t)v;
    else if (cb == 4) ((int32_t *)out)[i] = (int32_t)v;
    else ((int64_t *)out)[i] = v;
  }
  return B2D_OK;
}
