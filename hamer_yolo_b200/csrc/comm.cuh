// NCCL binding for hm_allgather: the one collective of the path (north_star: "an NCCL all-gather over NVLink only to
// collect per-hand MANO outputs"; the reference itself is single-process, hamer/infer.py:1268-1274).
// libnccl is resolved at run time: if the process already holds a libnccl.so.2 (PyTorch's bundled copy) that instance is
// used, otherwise the system library is opened.  No link-time dependency, so the library still loads on a box without NCCL
// (the comm entry points then fail with a message).
#pragma once
#include <dlfcn.h>
#include <nccl.h>

#include <string>

namespace hm {

struct NcclApi {
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*CommCount)(const ncclComm_t, int*) = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
    ncclResult_t (*GetVersion)(int*) = nullptr;
    std::string error;
    bool ok = false;
};

inline NcclApi& nccl_api() {
    static NcclApi api = [] {
        NcclApi a;
        void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD);          // already in the process (torch's copy)?
        if (!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
        if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
        if (!h) { a.error = std::string("libnccl.so.2 not found: ") + dlerror(); return a; }
        auto sym = [&](const char* n) { return dlsym(h, n); };
        a.GetUniqueId = reinterpret_cast<decltype(a.GetUniqueId)>(sym("ncclGetUniqueId"));
        a.CommInitRank = reinterpret_cast<decltype(a.CommInitRank)>(sym("ncclCommInitRank"));
        a.CommDestroy = reinterpret_cast<decltype(a.CommDestroy)>(sym("ncclCommDestroy"));
        a.AllGather = reinterpret_cast<decltype(a.AllGather)>(sym("ncclAllGather"));
        a.CommCount = reinterpret_cast<decltype(a.CommCount)>(sym("ncclCommCount"));
        a.GetErrorString = reinterpret_cast<decltype(a.GetErrorString)>(sym("ncclGetErrorString"));
        a.GetVersion = reinterpret_cast<decltype(a.GetVersion)>(sym("ncclGetVersion"));
        a.ok = a.GetUniqueId && a.CommInitRank && a.CommDestroy && a.AllGather && a.CommCount && a.GetErrorString;
        if (!a.ok) a.error = "libnccl.so.2 lacks a required symbol";
        return a;
    }();
    return api;
}

}  // namespace hm
