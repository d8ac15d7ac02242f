// cic_fast.cu -- specialised CIC kernels (placeholder until the generic path is validated on the GPU).
#include "kernels.h"

namespace b2d {
bool cic_fast_supported(const CicLaunch &) { return false; }
cudaError_t launch_cic_fast(const CicLaunch &, cudaStream_t) { return cudaErrorNotSupported; }
}  // namespace b2d
