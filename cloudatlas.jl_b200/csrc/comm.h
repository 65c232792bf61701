// The library's NCCL communicator (comm.cu): NCCL resolved at run time, one communicator per local device.
#pragma once
#include <nccl.h>

#include <vector>

#include "common.h"

namespace ca {

struct NcclApi {
  void* handle = nullptr;
  ncclResult_t (*GetVersion)(int*) = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommInitAll)(ncclComm_t*, int, const int*) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*Broadcast)(const void*, void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*Send)(const void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*Recv)(void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*GroupStart)() = nullptr;
  ncclResult_t (*GroupEnd)() = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;
};

// the run-time resolved NCCL entry points (dlopen libnccl.so.2); fails with CA_ERR_UNSUPPORTED if NCCL is not there
const NcclApi& nccl();

}  // namespace ca

#define CA_NCCL(expr)                                                                                            \
  do {                                                                                                           \
    ncclResult_t r__ = (expr);                                                                                   \
    if (r__ != ncclSuccess)                                                                                      \
      ::ca::fail(CA_ERR_CUDA, "%s failed at %s:%d: %s", #expr, __FILE__, __LINE__, ::ca::nccl().GetErrorString(r__)); \
  } while (0)

struct ca_comm {
  int nranks = 0;
  std::vector<int> devices;  // local devices, one communicator each
  std::vector<int> ranks;    // their ranks
  std::vector<ncclComm_t> comms;
  std::vector<cudaStream_t> streams;
  ~ca_comm();
};

// the communicator installed by ca_init (may be null)
ca_comm* default_comm();
