// nccl_dl.h -- NCCL resolved at run time (dlopen "libnccl.so.2"), so libparakeet_b200.so has no
// link-time NCCL dependency and uses the SAME NCCL the host process already loaded (torch's bundled
// copy under torch.distributed, the system library under a C++ host).  Only the five entry points
// the single exchange step needs: unique id, comm init / destroy, all-gather.
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>

namespace pk {

struct NcclApi {
    // signatures of nccl.h (2.x): ncclResult_t is an int enum, ncclComm_t an opaque pointer,
    // ncclUniqueId a 128-byte struct passed BY VALUE to ncclCommInitRank, ncclInt32 == 2.
    struct UniqueId { char internal[128]; };
    int (*GetUniqueId)(UniqueId *) = nullptr;
    int (*CommInitRank)(void **comm, int nranks, UniqueId id, int rank) = nullptr;
    int (*CommDestroy)(void *comm) = nullptr;
    int (*AllGather)(const void *send, void *recv, size_t count, int dtype, void *comm, cudaStream_t st) = nullptr;
    const char *(*GetErrorString)(int) = nullptr;
    bool ok = false;
    const char *why = "";
};

// Loads the library once per process; returns an object with ok == false (and `why`) if NCCL is absent.
const NcclApi &nccl_api();

}  // namespace pk
