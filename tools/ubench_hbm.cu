// tools/ubench_hbm.cu -- what HBM rate does a streaming kernel reach on this B200 for the read:write mixes of the
// CIC kernels?  (The roofline denominator in MEASURED_PEAKS.json is a 1:1 copy.)  Prints one JSON line per mix.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/bin/ubench_hbm tools/ubench_hbm.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)

// each thread reads RD uint4 and writes WR uint4 per iteration (grid-stride over "groups")
template <int RD, int WR>
__global__ void __launch_bounds__(256) mix_kernel(const uint4 *__restrict__ in, uint4 *__restrict__ out, size_t groups) {
  for (size_t g = (size_t)blockIdx.x * blockDim.x + threadIdx.x; g < groups; g += (size_t)gridDim.x * blockDim.x) {
    uint4 acc = make_uint4(0, 0, 0, 0);
#pragma unroll
    for (int r = 0; r < RD; r++) {
      const uint4 v = in[g + (size_t)r * groups];
      acc.x += v.x; acc.y ^= v.y; acc.z += v.z; acc.w ^= v.w;
    }
    if (WR == 0) { if (acc.x == 0x12345678u && acc.y == 0x9abcdef0u) out[0] = acc; }
#pragma unroll
    for (int w = 0; w < WR; w++) { acc.x += w; out[g + (size_t)w * groups] = acc; }
  }
}

template <int RD, int WR>
static void run(const char *name, uint4 *a, uint4 *b, size_t bytes_total, int ctas_per_sm) {
  const size_t groups = bytes_total / 16 / (RD + WR);
  const int grid = 148 * ctas_per_sm;
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  float best = 1e30f;
  for (int it = 0; it < 8; it++) {
    CK(cudaEventRecord(e0));
    mix_kernel<RD, WR><<<grid, 256>>>(a, b, groups);
    CK(cudaEventRecord(e1));
    CK(cudaEventSynchronize(e1));
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
    if (it >= 2 && ms < best) best = ms;
  }
  const double gb = (double)groups * 16 * (RD + WR) / 1e9;
  printf("{\"mix\": \"%s\", \"read_parts\": %d, \"write_parts\": %d, \"ctas_per_sm\": %d, \"GB\": %.3f, \"ms\": %.4f, \"GBps\": %.1f}\n",
         name, RD, WR, ctas_per_sm, gb, best, gb / best * 1e3);
}

int main() {
  const size_t bytes = (size_t)4 << 30;
  uint4 *a, *b;
  CK(cudaMalloc(&a, bytes)); CK(cudaMalloc(&b, bytes));
  CK(cudaMemset(a, 1, bytes)); CK(cudaMemset(b, 2, bytes));
  for (int c : {4, 8, 16}) {
    run<1, 0>("read only", a, b, bytes, c);
    run<0, 1>("write only", a, b, bytes, c);
    run<1, 1>("copy 1:1", a, b, bytes, c);
    run<4, 1>("cic_dec 4:1 (R=8 int16 IQ in, int32 out)", a, b, bytes, c);
    run<1, 8>("cic_intr 1:8 (R=4 int16 in, int32 out)", a, b, bytes, c);
    run<1, 4>("fir 1:4 (int16 in, int64 out)", a, b, bytes, c);
  }
  float ms; cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  CK(cudaEventRecord(e0)); CK(cudaMemcpyAsync(b, a, bytes, cudaMemcpyDeviceToDevice)); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
  CK(cudaEventRecord(e0)); CK(cudaMemcpyAsync(b, a, bytes, cudaMemcpyDeviceToDevice)); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
  CK(cudaEventElapsedTime(&ms, e0, e1));
  printf("{\"mix\": \"cudaMemcpy D2D\", \"GB\": %.3f, \"ms\": %.4f, \"GBps\": %.1f}\n", 2.0 * bytes / 1e9, ms, 2.0 * bytes / 1e9 / ms * 1e3);
  return 0;
}
