// Process-wide cache of released device buffers.
//
// Setting a solver up is dominated by the driver, not by the kernels: cudaMalloc / cudaFree of the
// two population buffers of a 256^3 lattice (1.3 GB each) take ~8 of the ~11 ms of lbm_init, and
// on a busy driver occasionally 100-300 ms (measured on B200: bench.py's end-to-end job, whose
// 20 steps take 7.8 ms).  A job that creates solvers repeatedly -- a parameter sweep, a restart, the
// reference's own pattern of init_geo(); init_simulation() on a new sample -- re-uses the buffers of
// the solver it destroyed before: released blocks are kept (at most LBM3D_POOL_MB MiB in total,
// default 8192, 0 = off, and at most 512 blocks; oldest evicted first) and handed out again to an
// allocation of exactly the same size on the same device -- a second solver of the same shape makes
// no cudaMalloc / cudaFree call at all.  Contents are NOT cleared: callers initialise what
// they allocate, as with cudaMalloc.  When the driver is out of memory the cache is emptied and the
// allocation retried; lbm_pool_trim() empties it on request.
#pragma once
#include <cuda_runtime.h>

#include <cstdlib>
#include <mutex>
#include <unordered_map>
#include <vector>

namespace lbm_pool {

struct Block {
    void *p;
    size_t bytes;
    int dev;
};
struct State {
    std::mutex mu;
    std::vector<Block> cached;                       // oldest first
    std::unordered_map<void *, Block> live;          // blocks handed out by alloc()
    size_t cached_bytes = 0;
    size_t cap = 0;
    bool cap_read = false;
};
inline State g_state;
constexpr size_t kMinBytes = 1;
constexpr size_t kMaxBlocks = 512;

inline size_t capacity(State &s) {
    if (!s.cap_read) {
        const char *e = getenv("LBM3D_POOL_MB");
        long long mb = e ? atoll(e) : 8192;
        s.cap = mb > 0 ? (size_t)mb << 20 : 0;
        s.cap_read = true;
    }
    return s.cap;
}

// frees every cached block (of one device, or of all with dev < 0); returns the bytes given back
inline size_t trim(int dev = -1) {
    State &s = g_state;
    std::lock_guard<std::mutex> lk(s.mu);
    int cur = 0;
    cudaGetDevice(&cur);
    size_t freed = 0;
    std::vector<Block> keep;
    for (const Block &b : s.cached) {
        if (dev >= 0 && b.dev != dev) { keep.push_back(b); continue; }
        cudaSetDevice(b.dev);
        cudaFree(b.p);
        freed += b.bytes;
    }
    s.cached.swap(keep);
    s.cached_bytes -= freed;
    cudaSetDevice(cur);
    return freed;
}

// like cudaMalloc on the current device
inline cudaError_t alloc(void **p, size_t bytes) {
    State &s = g_state;
    int dev = 0;
    cudaGetDevice(&dev);
    if (bytes >= kMinBytes && capacity(s) > 0) {
        std::lock_guard<std::mutex> lk(s.mu);
        for (size_t i = s.cached.size(); i-- > 0;) {              // most recently released first
            if (s.cached[i].bytes == bytes && s.cached[i].dev == dev) {
                Block b = s.cached[i];
                s.cached.erase(s.cached.begin() + (long)i);
                s.cached_bytes -= bytes;
                s.live[b.p] = b;
                *p = b.p;
                return cudaSuccess;
            }
        }
    }
    cudaError_t e = cudaMalloc(p, bytes);
    if (e == cudaErrorMemoryAllocation && trim(dev) > 0) {
        cudaGetLastError();
        e = cudaMalloc(p, bytes);
    }
    if (e == cudaSuccess && bytes >= kMinBytes && capacity(s) > 0) {
        std::lock_guard<std::mutex> lk(s.mu);
        s.live[*p] = Block{*p, bytes, dev};
    }
    return e;
}

// like cudaFree, including its wait for the device: work that still uses the block may be in flight
inline void release(void *p) {
    if (!p) return;
    State &s = g_state;
    cudaDeviceSynchronize();
    std::vector<Block> evict;
    bool keep = false;
    {
        std::lock_guard<std::mutex> lk(s.mu);
        auto it = s.live.find(p);
        if (it != s.live.end()) {
            const Block b = it->second;
            s.live.erase(it);
            if (b.bytes <= capacity(s)) {
                while ((s.cached_bytes + b.bytes > s.cap || s.cached.size() >= kMaxBlocks) && !s.cached.empty()) {
                    evict.push_back(s.cached.front());
                    s.cached_bytes -= s.cached.front().bytes;
                    s.cached.erase(s.cached.begin());
                }
                s.cached.push_back(b);
                s.cached_bytes += b.bytes;
                keep = true;
            }
        }
    }
    if (!evict.empty()) {
        int cur = 0;
        cudaGetDevice(&cur);
        for (const Block &b : evict) {
            cudaSetDevice(b.dev);
            cudaFree(b.p);
        }
        cudaSetDevice(cur);
    }
    if (!keep) cudaFree(p);
}

}  // namespace lbm_pool
