// tools/ubench_pcie.cu -- what the host link of this box can carry: bare pinned cudaMemcpyAsync, no kernels.
//
// The e2e figure of bench.py (C-ABI with HOST buffers) is bounded by the H2D / D2H copies, not by a kernel
// ("what's weak" #3).  This measures the ceiling those copies can reach, on 1 .. all visible GPUs
// at once (one host thread per GPU, all started together), in the directions and mixes the engine uses:
//   d2h        device -> pinned host only
//   h2d        pinned host -> device only
//   duplex     both at once, equal sizes
//   fir256     both at once, 4 B in : 16 B out (ac_fixed<40,8> in int64 containers)
//   fir256p    both at once, 4 B in : 10 B out (the packed 5-byte wire format)
// Variants of the host allocation: cudaHostAlloc default, and write-combined for the H2D source.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/bin/ubench_pcie tools/ubench_pcie.cu -lpthread
// Output: one JSON line per (pattern, n_gpus): aggregate and slowest-GPU GB/s per direction (events on each GPU,
// wall clock across all of them).
#include <atomic>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <thread>
#include <vector>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e__ = (x); if (e__ != cudaSuccess) { fprintf(stderr, "CUDA error %s at line %d\n", cudaGetErrorString(e__), __LINE__); exit(1); } } while (0)

struct Pattern { const char *name; double in_frac, out_frac; };   // fractions of the base chunk per direction
static const Pattern kPatterns[] = {{"d2h", 0, 1}, {"h2d", 1, 0}, {"duplex", 1, 1}, {"fir256", 0.25, 1}, {"fir256p", 0.4, 1}};

struct Barrier {
  std::atomic<int> count{0}, gen{0};
  int n;
  explicit Barrier(int n_) : n(n_) {}
  void wait() {
    const int g = gen.load();
    if (count.fetch_add(1) + 1 == n) { count.store(0); gen.fetch_add(1); }
    else while (gen.load() == g) std::this_thread::yield();
  }
};

struct Result { double in_gbs = 0, out_gbs = 0, ms = 0; };

static void worker(int dev, size_t chunk, int reps, const Pattern &pt, bool wc, Barrier *bar, Result *res) {
  CK(cudaSetDevice(dev));
  const size_t in_b = (size_t)(chunk * pt.in_frac), out_b = (size_t)(chunk * pt.out_frac);
  void *h_in = nullptr, *h_out = nullptr, *d_in = nullptr, *d_out = nullptr;
  cudaStream_t s_in, s_out;
  CK(cudaStreamCreateWithFlags(&s_in, cudaStreamNonBlocking));
  CK(cudaStreamCreateWithFlags(&s_out, cudaStreamNonBlocking));
  if (in_b) { CK(cudaHostAlloc(&h_in, in_b, wc ? cudaHostAllocWriteCombined : cudaHostAllocDefault)); memset(h_in, 1, in_b); CK(cudaMalloc(&d_in, in_b)); }
  if (out_b) { CK(cudaHostAlloc(&h_out, out_b, cudaHostAllocDefault)); memset(h_out, 0, out_b); CK(cudaMalloc(&d_out, out_b)); CK(cudaMemset(d_out, 2, out_b)); }
  cudaEvent_t e0, e1, f0, f1;
  CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1)); CK(cudaEventCreate(&f0)); CK(cudaEventCreate(&f1));
  for (int w = 0; w < 2; w++) {   // warm-up
    if (in_b) CK(cudaMemcpyAsync(d_in, h_in, in_b, cudaMemcpyHostToDevice, s_in));
    if (out_b) CK(cudaMemcpyAsync(h_out, d_out, out_b, cudaMemcpyDeviceToHost, s_out));
  }
  CK(cudaDeviceSynchronize());
  bar->wait();
  const auto t0 = std::chrono::steady_clock::now();
  CK(cudaEventRecord(e0, s_in)); CK(cudaEventRecord(f0, s_out));
  for (int r = 0; r < reps; r++) {
    if (in_b) CK(cudaMemcpyAsync(d_in, h_in, in_b, cudaMemcpyHostToDevice, s_in));
    if (out_b) CK(cudaMemcpyAsync(h_out, d_out, out_b, cudaMemcpyDeviceToHost, s_out));
  }
  CK(cudaEventRecord(e1, s_in)); CK(cudaEventRecord(f1, s_out));
  CK(cudaDeviceSynchronize());
  const auto t1 = std::chrono::steady_clock::now();
  bar->wait();
  float ms_in = 0, ms_out = 0;
  CK(cudaEventElapsedTime(&ms_in, e0, e1)); CK(cudaEventElapsedTime(&ms_out, f0, f1));
  res->ms = std::chrono::duration<double, std::milli>(t1 - t0).count();
  res->in_gbs = in_b ? (double)in_b * reps / (ms_in * 1e-3) / 1e9 : 0;
  res->out_gbs = out_b ? (double)out_b * reps / (ms_out * 1e-3) / 1e9 : 0;
  if (h_in) cudaFreeHost(h_in);
  if (h_out) cudaFreeHost(h_out);
  if (d_in) cudaFree(d_in);
  if (d_out) cudaFree(d_out);
  cudaStreamDestroy(s_in); cudaStreamDestroy(s_out);
}

// The engine's host pipeline in miniature (rt_common.h: run_host_pipeline): three slots, H2D / kernel / D2H on three
// streams chained by events, 4 B in : 16 B out per unit.  `busy_us` > 0 runs a kernel that keeps every SM busy for about
// that long per chunk (the FIR kernel's place); 0 launches nothing.  Answers: does the D2H leg reach the bare-copy rate
// when small H2D copies and a compute kernel run beside it?
__global__ void spin_kernel(long long cycles, int *sink) {
  const long long t0 = clock64();
  int v = 0;
  while (clock64() - t0 < cycles) v++;
  if (v == -1) *sink = v;
}

static void pipeline_probe(size_t chunk, int nchunks, int busy_us, int same_stream_copies) {
  CK(cudaSetDevice(0));
  const size_t out_b = chunk, in_b = chunk / 4;
  void *h_in, *h_out, *d_in[3], *d_out[3];
  CK(cudaHostAlloc(&h_in, in_b * nchunks, cudaHostAllocPortable));
  CK(cudaHostAlloc(&h_out, out_b * nchunks, cudaHostAllocPortable));
  memset(h_in, 1, in_b * nchunks); memset(h_out, 0, out_b * nchunks);
  for (int i = 0; i < 3; i++) { CK(cudaMalloc(&d_in[i], in_b)); CK(cudaMalloc(&d_out[i], out_b)); }
  cudaStream_t s_in, s_k, s_out;
  CK(cudaStreamCreateWithFlags(&s_in, cudaStreamNonBlocking));
  CK(cudaStreamCreateWithFlags(&s_k, cudaStreamNonBlocking));
  CK(cudaStreamCreateWithFlags(&s_out, cudaStreamNonBlocking));
  if (same_stream_copies) s_out = s_in;
  cudaEvent_t e_in[3], e_k[3], e_out[3], t0, t1;
  for (int i = 0; i < 3; i++) { cudaEventCreateWithFlags(&e_in[i], cudaEventDisableTiming); cudaEventCreateWithFlags(&e_k[i], cudaEventDisableTiming); cudaEventCreateWithFlags(&e_out[i], cudaEventDisableTiming); }
  CK(cudaEventCreate(&t0)); CK(cudaEventCreate(&t1));
  int *sink; CK(cudaMalloc(&sink, 4));
  cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
  const long long cycles = (long long)busy_us * (p.clockRate / 1000);
  for (int rep = 0; rep < 2; rep++) {
    CK(cudaDeviceSynchronize());
    const auto w0 = std::chrono::steady_clock::now();
    for (int i = 0; i < nchunks; i++) {
      const int s = i % 3;
      if (i >= 3) CK(cudaStreamWaitEvent(s_in, e_k[s], 0));
      CK(cudaMemcpyAsync(d_in[s], (char *)h_in + (size_t)i * in_b, in_b, cudaMemcpyHostToDevice, s_in));
      CK(cudaEventRecord(e_in[s], s_in));
      CK(cudaStreamWaitEvent(s_k, e_in[s], 0));
      if (i >= 3) CK(cudaStreamWaitEvent(s_k, e_out[s], 0));
      if (busy_us > 0) spin_kernel<<<p.multiProcessorCount * 4, 128, 0, s_k>>>(cycles, sink);
      CK(cudaEventRecord(e_k[s], s_k));
      CK(cudaStreamWaitEvent(s_out, e_k[s], 0));
      CK(cudaMemcpyAsync((char *)h_out + (size_t)i * out_b, d_out[s], out_b, cudaMemcpyDeviceToHost, s_out));
      CK(cudaEventRecord(e_out[s], s_out));
    }
    CK(cudaDeviceSynchronize());
    const double ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - w0).count();
    if (rep == 1)
      printf("{\"pattern\": \"pipeline\", \"chunk_mib\": %.1f, \"chunks\": %d, \"kernel_busy_us\": %d, \"copies_share_a_stream\": %d, \"d2h_gbs\": %.1f, \"h2d_gbs\": %.1f, \"wall_ms\": %.2f}\n",
             chunk / 1048576.0, nchunks, busy_us, same_stream_copies, (double)out_b * nchunks / (ms * 1e-3) / 1e9, (double)in_b * nchunks / (ms * 1e-3) / 1e9, ms);
  }
  fflush(stdout);
  cudaFreeHost(h_in); cudaFreeHost(h_out);
  for (int i = 0; i < 3; i++) { cudaFree(d_in[i]); cudaFree(d_out[i]); }
}

int main(int argc, char **argv) {
  if (argc > 1 && !strcmp(argv[1], "pipeline")) {
    for (int busy : {0, 50, 100, 300})
      for (size_t mib : {20, 80}) pipeline_probe(mib << 20, (int)(2048 / mib), busy, 0);
    pipeline_probe((size_t)20 << 20, 100, 100, 1);
    return 0;
  }
  int ndev = 0;
  CK(cudaGetDeviceCount(&ndev));
  const size_t chunk = (size_t)(argc > 1 ? atoi(argv[1]) : 256) << 20;   // MiB per D2H copy
  const int reps = argc > 2 ? atoi(argv[2]) : 12;
  cudaDeviceProp p;
  CK(cudaGetDeviceProperties(&p, 0));
  printf("{\"device\": \"%s\", \"visible_gpus\": %d, \"chunk_mib\": %zu, \"reps\": %d, \"host_threads\": %u}\n", p.name, ndev, chunk >> 20, reps,
         std::thread::hardware_concurrency());
  for (int wc = 0; wc < 2; wc++)
    for (const Pattern &pt : kPatterns) {
      if (wc && pt.in_frac == 0) continue;
      for (int g = 1; g <= ndev; g *= 2) {
        Barrier bar(g);
        std::vector<Result> res(g);
        std::vector<std::thread> th;
        for (int d = 0; d < g; d++) th.emplace_back(worker, d, chunk, reps, std::cref(pt), wc != 0, &bar, &res[d]);
        for (auto &t : th) t.join();
        double in_sum = 0, out_sum = 0, in_min = 1e30, out_min = 1e30, wall = 0;
        for (const Result &r : res) { in_sum += r.in_gbs; out_sum += r.out_gbs; in_min = r.in_gbs < in_min ? r.in_gbs : in_min; out_min = r.out_gbs < out_min ? r.out_gbs : out_min; wall = r.ms > wall ? r.ms : wall; }
        const double bytes = ((double)(size_t)(chunk * pt.in_frac) + (double)(size_t)(chunk * pt.out_frac)) * reps * g;
        printf("{\"pattern\": \"%s\", \"h2d_source\": \"%s\", \"n_gpus\": %d, \"h2d_gbs_sum\": %.1f, \"d2h_gbs_sum\": %.1f, \"h2d_gbs_min\": %.1f, \"d2h_gbs_min\": %.1f, "
               "\"wall_ms\": %.2f, \"aggregate_gbs_wall\": %.1f}\n",
               pt.name, wc ? "write-combined" : "default", g, in_sum, out_sum, pt.in_frac ? in_min : 0.0, pt.out_frac ? out_min : 0.0, wall, bytes / (wall * 1e-3) / 1e9);
        fflush(stdout);
      }
    }
  return 0;
}
