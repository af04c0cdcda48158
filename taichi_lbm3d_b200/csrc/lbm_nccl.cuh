// NCCL, bound at run time from the library torch has already loaded (shared by the single-phase
// and two-phase C-ABI layers; each translation unit gets its own copy of the binding).
#pragma once
#include <cuda_runtime.h>
#include <dlfcn.h>

#include <cstdlib>
#include <string>

namespace {

// (prototypes from nccl.h 2.28: ncclUniqueId is 128 bytes, ncclFloat = 7, ncclSuccess = 0)
struct NcclId { char internal[128]; };
struct NcclApi {
    void *h = nullptr;
    int (*GetUniqueId)(NcclId *) = nullptr;
    int (*CommInitRank)(void **, int, NcclId, int) = nullptr;
    int (*CommDestroy)(void *) = nullptr;
    int (*Send)(const void *, size_t, int, int, void *, cudaStream_t) = nullptr;
    int (*Recv)(void *, size_t, int, int, void *, cudaStream_t) = nullptr;
    int (*GroupStart)() = nullptr;
    int (*GroupEnd)() = nullptr;
    const char *(*GetErrorString)(int) = nullptr;
    bool ok = false;
};
NcclApi g_nccl;

bool load_nccl(std::string &err) {
    if (g_nccl.ok) return true;
    const char *names[] = {getenv("LBM3D_NCCL_LIB"), "libnccl.so.2", "libnccl.so"};
    for (const char *n : names) {
        if (!n) continue;
        g_nccl.h = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
        if (g_nccl.h) break;
    }
    if (!g_nccl.h) { err = std::string("cannot load NCCL: ") + dlerror(); return false; }
#define BIND(field, sym)                                                                       \
    *(void **)(&g_nccl.field) = dlsym(g_nccl.h, sym);                                          \
    if (!g_nccl.field) { err = std::string("NCCL symbol missing: ") + sym; return false; }
    BIND(GetUniqueId, "ncclGetUniqueId")
    BIND(CommInitRank, "ncclCommInitRank")
    BIND(CommDestroy, "ncclCommDestroy")
    BIND(Send, "ncclSend")
    BIND(Recv, "ncclRecv")
    BIND(GroupStart, "ncclGroupStart")
    BIND(GroupEnd, "ncclGroupEnd")
    BIND(GetErrorString, "ncclGetErrorString")
#undef BIND
    g_nccl.ok = true;
    return true;
}

}  // namespace
