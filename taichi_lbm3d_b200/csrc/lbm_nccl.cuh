// NCCL, bound at run time from the library torch has already loaded.  One binding and ONE
// communicator per process, shared by the single-phase and two-phase C-ABI layers and by every
// context (C++17 inline variables): ncclCommInitRank costs seconds at 8 ranks, and several
// communicators with kernels in flight on different streams can dead-lock each other.
#pragma once
#include <cuda_runtime.h>
#include <dlfcn.h>

#include <cstdlib>
#include <string>

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
inline NcclApi g_nccl;

// the process-wide communicator: created by the first lbm_comm_init / lbm2p_comm_init of a
// (world, rank, device) and reused by every later context with the same triple; never destroyed
// before process exit (contexts do not own it)
struct NcclShared {
    void *comm = nullptr;
    int world = 0, rank = -1, device = -1;
};
inline NcclShared g_nccl_shared;

inline bool load_nccl(std::string &err) {
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

// returns 0 and the shared communicator (creating it from `id` when there is none for this triple)
inline int shared_comm(int world, int rank, const NcclId &id, void **comm) {
    int dev = 0;
    cudaGetDevice(&dev);
    NcclShared &s = g_nccl_shared;
    if (s.comm && s.world == world && s.rank == rank && s.device == dev) { *comm = s.comm; return 0; }
    if (s.comm) { g_nccl.CommDestroy(s.comm); s.comm = nullptr; }
    void *c = nullptr;
    const int r = g_nccl.CommInitRank(&c, world, id, rank);
    if (r != 0) return r;
    s.comm = c; s.world = world; s.rank = rank; s.device = dev;
    *comm = c;
    return 0;
}
inline bool have_shared_comm(int world, int rank) {
    int dev = 0;
    cudaGetDevice(&dev);
    const NcclShared &s = g_nccl_shared;
    return s.comm && s.world == world && s.rank == rank && s.device == dev;
}
