// tools/ubench_pipes.cu -- issue-rate microbenchmark for the integer / FP pipes of sm_100a.
//
// The FIR tap-MAC kernels are bounded by CUDA-core multiply-accumulate issue, not by HBM
// (SURVEY.md 8d), so the kernel design (IMAD vs IMAD.WIDE vs DP4A vs exact-integer DFMA, and
// which mixes dual-issue across pipes) is chosen from these measurements.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o build/ubench_pipes tools/ubench_pipes.cu
// Output: one JSON line per experiment: lane-ops per clock per SM (from clock64) and Tops/s (events).
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); return 1; } } while (0)

enum Op { IMAD_LO, IMAD_WIDE, IMAD_WIDE_U, IMAD_HI, DP4A, DP2A, FFMA, FFMA2, DFMA, IADD, IADD64, LOP3,
          MIX_IMAD_IADD, MIX_WIDE_DFMA, MIX_IMAD_FFMA, MIX_IMAD_DFMA, MIX_WIDE_IADD, MIX_WIDE_FFMA, MIX_WIDE_DFMA_21,
          MIX_WIDE_LDS, MIX_DFMA_IADD, MIX_WIDE_DFMA_IADD, MIX_IMAD_LOP3, MIX_IMAD_SHF, MIX_IMAD_PRMT, MIX_IMAD2_DFMA, MIX_IMAD_DP4A, MIX_IMAD_LDS, SHF, PRMT,
          MIX_DP2A_FFMA_21, MIX_DP2A_FFMA_41, MIX_DP2A_FFMA_81, MIX_DP2A_IADD_41, MIX_DP2A_PRMT_81, NOPS };
static const char *kNames[] = {"imad.lo.s32", "imad.wide.s32", "imad.wide.u32", "imad.hi.s32", "dp4a.s32", "dp2a.lo.s32", "ffma.f32",
                               "ffma2.f32x2(2 fma/lane)", "dfma.f64", "iadd.s32", "iadd.s64", "lop3",
                               "mix imad.lo+iadd 1:1", "mix imad.wide+dfma 1:1", "mix imad.lo+ffma 1:1", "mix imad.lo+dfma 1:1",
                               "mix imad.wide+iadd 1:1", "mix imad.wide+ffma 1:1", "mix imad.wide+dfma 2:1",
                               "mix imad.wide+lds32 4:1", "mix dfma+iadd 1:1", "mix imad.wide+dfma+iadd 1:1:1",
                               "mix imad.lo+lop3 1:1", "mix imad.lo+shf 1:1", "mix imad.lo+prmt 1:1", "mix imad.lo+dfma 2:1", "mix imad.lo+dp4a 1:1", "mix imad.lo+lds32 8:1", "shf.r", "prmt",
                               "mix dp2a+ffma 2:1", "mix dp2a+ffma 4:1", "mix dp2a+ffma 8:1", "mix dp2a+iadd 4:1", "mix dp2a+prmt 8:1"};

constexpr int K = 8;        // independent accumulators per thread
constexpr int UNROLL = 32;  // "taps" per loop trip; every (x[(k+u)&7], h[u]) product in a trip is distinct

// Shaped like the FIR inner loop: acc[k] += x[(k+u) & 7] * h[u].  Multiplicands are independent of the
// accumulators (as in the real kernel) and all products within a trip are distinct, so ptxas can neither
// hoist nor CSE them; x[] is perturbed once per trip (8 adds per 256 MACs).
template <int OP>
__global__ void __launch_bounds__(512, 1) bench(int iters, int32_t a0, int32_t b0, long long *sink, long long *cycles) {
  __shared__ int32_t sm[1024];
  for (int i = threadIdx.x; i < 1024; i += blockDim.x) sm[i] = i * a0;
  __syncthreads();
  constexpr bool kInt = (OP != FFMA && OP != FFMA2 && OP != DFMA && OP != MIX_DFMA_IADD);
  constexpr bool kF32 = (OP == FFMA || OP == MIX_IMAD_FFMA || OP == MIX_WIDE_FFMA);
  // r02: would the FMA-lite pipe take exact-integer FFMAs next to a full DP2A stream?  
  constexpr int kSide = OP == MIX_DP2A_FFMA_21 ? 2 : ((OP == MIX_DP2A_FFMA_41 || OP == MIX_DP2A_IADD_41) ? 4 : ((OP == MIX_DP2A_FFMA_81 || OP == MIX_DP2A_PRMT_81) ? 8 : 0));
  constexpr bool kSideF = (OP == MIX_DP2A_FFMA_21 || OP == MIX_DP2A_FFMA_41 || OP == MIX_DP2A_FFMA_81);
  constexpr bool kF64 = (OP == MIX_IMAD2_DFMA || OP == DFMA || OP == MIX_WIDE_DFMA || OP == MIX_IMAD_DFMA || OP == MIX_WIDE_DFMA_21 || OP == MIX_DFMA_IADD || OP == MIX_WIDE_DFMA_IADD);
  int32_t b = b0 - threadIdx.x;
  int32_t x[8], h[UNROLL], r[K], q[K];
  float xf[8], hf[UNROLL], f[K];
  double xd[8], hd[UNROLL], d[K];
  unsigned long long x2[8], h2[UNROLL], f2[K];
  long long w[K];
#pragma unroll
  for (int k = 0; k < K; k++) { r[k] = k + threadIdx.x; q[k] = 7 * k - threadIdx.x; w[k] = k * 3 + threadIdx.x; f[k] = (float)k; d[k] = (double)k; f2[k] = k; }
#pragma unroll
  for (int i = 0; i < 8; i++) { x[i] = a0 * (i + 1) + threadIdx.x; xf[i] = 1.0f + i * 1e-3f * a0; xd[i] = 1.0 + i * 1e-6 * a0; x2[i] = 0x3f8000003f800000ULL + i * a0; }
#pragma unroll
  for (int u = 0; u < UNROLL; u++) { h[u] = b0 * (u + 3) - threadIdx.x; hf[u] = 1e-3f * (u + b0); hd[u] = 1e-6 * (u + b0); h2[u] = 0x3a8000003a800000ULL + u * b0; }
  int lidx = threadIdx.x & 1023;
  long long t0 = clock64();
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int u = 0; u < UNROLL; u++) {
#pragma unroll
      for (int k = 0; k < K; k++) {
        const int xi = (k + u) & 7;
        if (OP == IMAD_LO || OP == MIX_IMAD_IADD || OP == MIX_IMAD_FFMA || OP == MIX_IMAD_DFMA || OP == MIX_IMAD_LOP3 || OP == MIX_IMAD_SHF || OP == MIX_IMAD_PRMT || OP == MIX_IMAD2_DFMA || OP == MIX_IMAD_DP4A || OP == MIX_IMAD_LDS)
          asm volatile("mad.lo.s32 %0, %1, %2, %0;" : "+r"(r[k]) : "r"(x[xi]), "r"(h[u]));
        if (OP == IMAD_WIDE || OP == MIX_WIDE_DFMA || OP == MIX_WIDE_IADD || OP == MIX_WIDE_FFMA || OP == MIX_WIDE_DFMA_21 || OP == MIX_WIDE_LDS || OP == MIX_WIDE_DFMA_IADD)
          asm volatile("mad.wide.s32 %0, %1, %2, %0;" : "+l"(w[k]) : "r"(x[xi]), "r"(h[u]));
        if (OP == MIX_WIDE_DFMA_21)
          asm volatile("mad.wide.s32 %0, %1, %2, %0;" : "+l"(w[k]) : "r"(x[xi]), "r"(h[(u + 1) % UNROLL]));
        if (OP == MIX_IMAD2_DFMA)
          asm volatile("mad.lo.s32 %0, %1, %2, %0;" : "+r"(q[k]) : "r"(x[xi]), "r"(h[(u + 1) % UNROLL]));
        if (OP == MIX_IMAD_LOP3)
          asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(q[k]) : "r"(x[xi]), "r"(h[u]));
        if (OP == MIX_IMAD_SHF || OP == SHF)
          asm volatile("shf.r.clamp.b32 %0, %0, %1, %2;" : "+r"(q[k]) : "r"(x[xi]), "r"(h[u]));
        if (OP == MIX_IMAD_PRMT || OP == PRMT)
          asm volatile("prmt.b32 %0, %0, %1, %2;" : "+r"(q[k]) : "r"(x[xi]), "r"(h[u]));
        if (OP == MIX_IMAD_DP4A)
          asm volatile("dp4a.s32.s32 %0, %1, %2, %0;" : "+r"(q[k]) : "r"(x[xi]), "r"(h[u]));
        if (OP == MIX_IMAD_LDS && k == 0) {
          int32_t v;
          asm volatile("ld.shared.s32 %0, [%1];" : "=r"(v) : "r"((unsigned)__cvta_generic_to_shared(&sm[(lidx + u * 32) & 1023])));
          q[k] ^= v;
        }
        if (OP == IMAD_WIDE_U)
          asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(w[k]) : "r"(x[xi]), "r"(h[u]));
        if (OP == IMAD_HI)
          asm volatile("mad.hi.s32 %0, %1, %2, %0;" : "+r"(r[k]) : "r"(x[xi]), "r"(h[u]));
        if (OP == DP4A)
          asm volatile("dp4a.s32.s32 %0, %1, %2, %0;" : "+r"(r[k]) : "r"(x[xi]), "r"(h[u]));
        if (OP == DP2A || kSide)
          asm volatile("dp2a.lo.s32.s32 %0, %1, %2, %0;" : "+r"(r[k]) : "r"(x[xi]), "r"(h[u]));
        if (kSide && kSideF && (k % kSide) == 0)
          asm volatile("fma.rn.f32 %0, %1, %2, %0;" : "+f"(f[k]) : "f"(xf[xi]), "f"(hf[u]));
        if (OP == MIX_DP2A_IADD_41 && (k % kSide) == 0)
          asm volatile("add.s32 %0, %0, %1;" : "+r"(q[k]) : "r"(x[xi]));
        if (OP == MIX_DP2A_PRMT_81 && (k % kSide) == 0)
          asm volatile("prmt.b32 %0, %0, %1, %2;" : "+r"(q[k]) : "r"(x[xi]), "r"(h[u]));
        if (kF32)
          asm volatile("fma.rn.f32 %0, %1, %2, %0;" : "+f"(f[k]) : "f"(xf[xi]), "f"(hf[u]));
        if (OP == FFMA2)
          asm volatile("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(f2[k]) : "l"(x2[xi]), "l"(h2[u]));
        if (kF64)
          asm volatile("fma.rn.f64 %0, %1, %2, %0;" : "+d"(d[k]) : "d"(xd[xi]), "d"(hd[u]));
        if (OP == IADD || OP == MIX_IMAD_IADD || OP == MIX_WIDE_IADD || OP == MIX_DFMA_IADD || OP == MIX_WIDE_DFMA_IADD)
          asm volatile("add.s32 %0, %0, %1;" : "+r"(q[k]) : "r"(x[xi]));
        if (OP == IADD64)
          asm volatile("add.s64 %0, %0, %1;" : "+l"(w[k]) : "l"((long long)x[xi]));
        if (OP == LOP3)
          asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(r[k]) : "r"(x[xi]), "r"(h[u]));
        if (OP == MIX_WIDE_LDS && (k & 3) == 0) {
          int32_t v;
          asm volatile("ld.shared.s32 %0, [%1];" : "=r"(v) : "r"((unsigned)__cvta_generic_to_shared(&sm[(lidx + u * 32 + k) & 1023])));
          r[k] ^= v;
        }
      }
    }
#pragma unroll
    for (int i = 0; i < 8; i++) {
      if (kInt) x[i] += b;
      if (kF32 || kSideF) xf[i] += 1e-7f;
      if (kF64) xd[i] += 1e-9;
      if (OP == FFMA2) x2[i] += 2;
    }
  }
  long long t1 = clock64();
  long long acc = 0;
#pragma unroll
  for (int k = 0; k < K; k++) acc += r[k] + q[k] + w[k] + (long long)f[k] + (long long)d[k] + (long long)f2[k];
  if (acc == 0x123456789LL) sink[0] = acc;
  if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

static int ops_per_body(int op) {  // lane-ops issued per (u,k) body
  switch (op) {
    case MIX_IMAD_LOP3: case MIX_IMAD_SHF: case MIX_IMAD_PRMT: case MIX_IMAD_DP4A:
    case MIX_IMAD_IADD: case MIX_WIDE_DFMA: case MIX_IMAD_FFMA: case MIX_IMAD_DFMA: case MIX_WIDE_IADD: case MIX_WIDE_FFMA: case MIX_DFMA_IADD: return 2;
    case MIX_WIDE_DFMA_21: case MIX_WIDE_DFMA_IADD: case MIX_IMAD2_DFMA: return 3;
    default: return 1;
  }
}

template <int OP>
int run_one(int sms, int blocks_per_sm, long long *d_sink, long long *d_cyc) {
  int iters = 500;
  int grid = sms * blocks_per_sm;
  bench<OP><<<grid, 512>>>(10, 3, 5, d_sink, d_cyc);
  CK(cudaDeviceSynchronize());
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  cudaEventRecord(e0);
  bench<OP><<<grid, 512>>>(iters, 3, 5, d_sink, d_cyc);
  cudaEventRecord(e1);
  CK(cudaEventSynchronize(e1));
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  long long *h = (long long *)malloc(sizeof(long long) * grid);
  cudaMemcpy(h, d_cyc, sizeof(long long) * grid, cudaMemcpyDeviceToHost);
  double avg = 0; long long mx = 0;
  for (int i = 0; i < grid; i++) { avg += h[i]; if (h[i] > mx) mx = h[i]; }
  avg /= grid;
  free(h);
  double lane_ops_per_sm = (double)iters * UNROLL * K * ops_per_body(OP) * 512.0 * blocks_per_sm;
  if (OP == MIX_WIDE_LDS || OP == MIX_IMAD_LDS) lane_ops_per_sm = (double)iters * UNROLL * K * 512.0 * blocks_per_sm;  // count the IMADs only
  double side = 0;   // r02 mixes: report the DP2A rate; the side stream rides along at 1/ratio of it
  if (OP == MIX_DP2A_FFMA_21) side = 2; else if (OP == MIX_DP2A_FFMA_41 || OP == MIX_DP2A_IADD_41) side = 4; else if (OP == MIX_DP2A_FFMA_81 || OP == MIX_DP2A_PRMT_81) side = 8;
  double total = lane_ops_per_sm * sms;
  if (side > 0)
    printf("{\"op\": \"%s\", \"dp2a_lane_ops_per_clk_per_sm\": %.2f, \"side_lane_ops_per_clk_per_sm\": %.2f, \"note\": \"dp2a alone = 63.7\"}\n",
           kNames[OP], lane_ops_per_sm / (double)mx, lane_ops_per_sm / (double)mx / side);
  printf("{\"op\": \"%s\", \"blocks_per_sm\": %d, \"lane_ops_per_clk_per_sm\": %.2f, \"tera_ops_per_s\": %.3f, \"ms\": %.3f, \"eff_clock_mhz\": %.0f}\n",
         kNames[OP], blocks_per_sm, lane_ops_per_sm / (double)mx, total / (ms * 1e-3) / 1e12, ms, (double)mx / (ms * 1e3));
  return 0;
}

template <int OP>
struct Runner {
  static int go(int sms, long long *s, long long *c) {
    if (run_one<OP>(sms, 1, s, c)) return 1;
    return Runner<OP + 1>::go(sms, s, c);
  }
};
template <>
struct Runner<NOPS> { static int go(int, long long *, long long *) { return 0; } };

int main() {
  cudaDeviceProp p;
  CK(cudaGetDeviceProperties(&p, 0));
  int sms = p.multiProcessorCount;
  printf("{\"device\": \"%s\", \"sms\": %d, \"clock_khz\": %d}\n", p.name, sms, p.clockRate);
  long long *d_sink, *d_cyc;
  CK(cudaMalloc(&d_sink, 64));
  CK(cudaMalloc(&d_cyc, sizeof(long long) * sms * 8));
  return Runner<0>::go(sms, d_sink, d_cyc);
}
