// nccl_dl.h -- NCCL bound at run time (dlopen) so that single-GPU users need no NCCL at all and a
// process that already carries an NCCL (e.g. torch's bundled one) shares it instead of loading a second.
// Only the handful of entry points the coefficient broadcast needs; types follow NCCL's public ABI.
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>

namespace b2d {

typedef struct ncclComm *ncclComm_t;
typedef struct { char internal[128]; } ncclUniqueId;
enum { ncclSuccess = 0 };
enum { ncclInt8 = 0, ncclInt32 = 2, ncclInt64 = 4 };
enum { ncclSum = 0 };

struct NcclApi {
  int (*GetUniqueId)(ncclUniqueId *);
  int (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int);
  int (*CommDestroy)(ncclComm_t);
  int (*Broadcast)(const void *, void *, size_t, int, int, ncclComm_t, cudaStream_t);
  int (*AllReduce)(const void *, void *, size_t, int, int, ncclComm_t, cudaStream_t);
  const char *(*GetErrorString)(int);
  int (*GetVersion)(int *);
};

// nullptr if no libnccl could be loaded (error text in *why).
const NcclApi *nccl_api(const char **why);

}  // namespace b2d
