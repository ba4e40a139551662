// nccl_dl.cpp -- see nccl_dl.h
#include "nccl_dl.h"

#include <dlfcn.h>

namespace pk {

const NcclApi &nccl_api() {
    static NcclApi api = [] {
        NcclApi a;
        // RTLD_NOLOAD first: a process that already runs NCCL (torch.distributed) must not get a second copy
        void *h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD | RTLD_GLOBAL);
        if (!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
        if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
        if (!h) {
            a.why = "libnccl.so.2 not found (dlopen)";
            return a;
        }
        a.GetUniqueId = reinterpret_cast<decltype(a.GetUniqueId)>(dlsym(h, "ncclGetUniqueId"));
        a.CommInitRank = reinterpret_cast<decltype(a.CommInitRank)>(dlsym(h, "ncclCommInitRank"));
        a.CommDestroy = reinterpret_cast<decltype(a.CommDestroy)>(dlsym(h, "ncclCommDestroy"));
        a.AllGather = reinterpret_cast<decltype(a.AllGather)>(dlsym(h, "ncclAllGather"));
        a.GetErrorString = reinterpret_cast<decltype(a.GetErrorString)>(dlsym(h, "ncclGetErrorString"));
        a.ok = a.GetUniqueId && a.CommInitRank && a.CommDestroy && a.AllGather && a.GetErrorString;
        if (!a.ok) a.why = "libnccl.so.2 lacks an expected symbol";
        return a;
    }();
    return api;
}

}  // namespace pk
