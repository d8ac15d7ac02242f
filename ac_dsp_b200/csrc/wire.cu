// wire.cu -- the packed host-link format (B2D_WIRE_PACKED): output values leave the device as ceil(W/8) little-endian
// bytes each instead of their 2 / 4 / 8-byte containers.  An ac_fixed<40,8> result is 5 bytes on the wire instead of 8,
// which is what the host link -- the bottleneck of every run() on host buffers -- carries per value.
// Runs on the compute stream of the host pipeline, between the filter kernel and the D2H copy of the slot.
#include "kernels.h"

namespace b2d {

// A thread packs 4 consecutive values: 4 * PB bytes = PB 32-bit words, word-aligned at 4 * PB * t.
template <int CB, int PB>
__global__ void __launch_bounds__(256) pack_wire_kernel(const void *src, uint32_t *dst, size_t count) {
  const size_t groups = count / 4;
  for (size_t g = (size_t)blockIdx.x * blockDim.x + threadIdx.x; g < groups; g += (size_t)gridDim.x * blockDim.x) {
    unsigned long long v[4];
    if (CB == 8) {
      const ulonglong2 a = ((const ulonglong2 *)src)[2 * g], b = ((const ulonglong2 *)src)[2 * g + 1];
      v[0] = a.x; v[1] = a.y; v[2] = b.x; v[3] = b.y;
    } else if (CB == 4) {
      const uint4 a = ((const uint4 *)src)[g];
      v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w;
    } else {
      const uint2 a = ((const uint2 *)src)[g];
      v[0] = a.x & 0xFFFFu; v[1] = a.x >> 16; v[2] = a.y & 0xFFFFu; v[3] = a.y >> 16;
    }
    // 4 * PB bytes, little endian, value k at byte offset k * PB
    uint32_t w[PB];
#pragma unroll
    for (int i = 0; i < PB; i++) w[i] = 0;
#pragma unroll
    for (int k = 0; k < 4; k++) {
      const unsigned long long m = v[k] & (PB == 8 ? ~0ULL : ((1ULL << (8 * (PB & 7))) - 1));
#pragma unroll
      for (int b = 0; b < PB; b++) {
        const int byte = k * PB + b;
        w[byte / 4] |= (uint32_t)((m >> (8 * b)) & 0xFFu) << (8 * (byte % 4));
      }
    }
#pragma unroll
    for (int i = 0; i < PB; i++) dst[g * PB + i] = w[i];
  }
  // the last count % 4 values, byte-wise
  if (blockIdx.x == 0 && threadIdx.x < (count & 3)) {
    const size_t e = groups * 4 + threadIdx.x;
    unsigned long long v;
    if (CB == 8) v = ((const unsigned long long *)src)[e];
    else if (CB == 4) v = ((const uint32_t *)src)[e];
    else v = ((const uint16_t *)src)[e];
    unsigned char *d = (unsigned char *)dst + e * PB;
    for (int b = 0; b < PB; b++) d[b] = (unsigned char)(v >> (8 * b));
  }
}

template <int CB, int PB>
static cudaError_t pack_launch(const void *src, void *dst, size_t count, cudaStream_t st) {
  size_t blocks = (count / 4 + 255) / 256;
  if (blocks < 1) blocks = 1;
  if (blocks > 148 * 16) blocks = 148 * 16;
  pack_wire_kernel<CB, PB><<<(unsigned)blocks, 256, 0, st>>>(src, (uint32_t *)dst, count);
  return cudaGetLastError();
}

cudaError_t launch_pack_wire(const void *src, int container_bytes_, void *dst, int wire_bytes, size_t count, cudaStream_t st) {
  if (count == 0) return cudaSuccess;
  switch (container_bytes_ * 10 + wire_bytes) {
    case 21: return pack_launch<2, 1>(src, dst, count, st);
    case 43: return pack_launch<4, 3>(src, dst, count, st);
    case 85: return pack_launch<8, 5>(src, dst, count, st);
    case 86: return pack_launch<8, 6>(src, dst, count, st);
    case 87: return pack_launch<8, 7>(src, dst, count, st);
    default: return cudaErrorInvalidValue;   // wire_bytes == container: the pipeline does not call this
  }
}

}  // namespace b2d
