// The one exchange step of the path: the sum of the per-GPU count matrices
// (combineReadCounts, /root/reference/tagdigger_fun.py:1088-1095, across GPUs), as ONE
// ncclAllReduce(int32, sum) issued on the context's own stream right behind the last count kernel.
//
// NCCL is bound at run time (dlopen "libnccl.so.2"): the library has no link-time dependency on
// it, a process that already carries an NCCL (PyTorch's) shares that copy, and a single-GPU
// caller never touches it.  Only the five entry points below are used; their prototypes follow
// nccl.h (2.x ABI).
#pragma once
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <stdint.h>

#include <string>

namespace tdg {

struct NcclApi {
    typedef struct { char internal[128]; } UniqueId;
    typedef void *Comm;
    enum { kInt32 = 2, kSum = 0 };     // ncclInt32, ncclSum

    int (*GetUniqueId)(UniqueId *) = nullptr;
    int (*CommInitRank)(Comm *, int, UniqueId, int) = nullptr;
    int (*CommDestroy)(Comm) = nullptr;
    int (*AllReduce)(const void *, void *, size_t, int, int, Comm, cudaStream_t) = nullptr;
    const char *(*GetErrorString)(int) = nullptr;
    void *handle = nullptr;
    std::string why;

    bool load()
    {
        if (handle) return true;
        const char *names[] = {"libnccl.so.2", "libnccl.so"};
        for (const char *n : names) {
            handle = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
            if (handle) break;
        }
        if (!handle) { why = std::string("cannot load NCCL: ") + dlerror(); return false; }
        GetUniqueId = (decltype(GetUniqueId))dlsym(handle, "ncclGetUniqueId");
        CommInitRank = (decltype(CommInitRank))dlsym(handle, "ncclCommInitRank");
        CommDestroy = (decltype(CommDestroy))dlsym(handle, "ncclCommDestroy");
        AllReduce = (decltype(AllReduce))dlsym(handle, "ncclAllReduce");
        GetErrorString = (decltype(GetErrorString))dlsym(handle, "ncclGetErrorString");
        if (!GetUniqueId || !CommInitRank || !CommDestroy || !AllReduce || !GetErrorString) {
            why = "libnccl lacks an expected entry point";
            return false;
        }
        return true;
    }
    std::string message(int rc) const { return GetErrorString ? GetErrorString(rc) : "NCCL error"; }
};

inline NcclApi &nccl()
{
    static NcclApi api;
    return api;
}

}  // namespace tdg
