// Device-side pieces of the sparse (compact fluid list) kernels shared by the single-phase and
// two-phase translation units: TMA bulk copy + mbarrier helpers, the shared-memory image of a
// 256-node slice of the compressed pull table, and its decode.  Included inside each unit (the
// functions are __device__ __forceinline__, no ODR issue).
#pragma once
#include "lbm_kernels.cuh"

#define SPARSE_BLOCK 256

// Gathers of the sparse kernel touch partial 128-byte lines (pores are a few nodes wide), but
// every line is consumed completely within one step by neighbouring warps.  The L2 prefetch-size
// hint makes a miss bring the whole line (or line pair) from HBM in one burst instead of one
// 32-byte sector per requesting warp.
#ifndef LBM_SPARSE_L2HINT
#define LBM_SPARSE_L2HINT 0
#endif
__device__ __forceinline__ float ldg_gather(const float *p) {
    float v;
#if LBM_SPARSE_L2HINT == 256
    asm("ld.global.nc.L2::256B.f32 %0, [%1];" : "=f"(v) : "l"(p));
#elif LBM_SPARSE_L2HINT == 128
    asm("ld.global.nc.L2::128B.f32 %0, [%1];" : "=f"(v) : "l"(p));
#else
    v = __ldg(p);
#endif
    return v;
}

// ---- TMA bulk copy + mbarrier (sm_90+/sm_100a PTX) ----------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}

// Sparse step.  A block owns the 256 consecutive stored nodes [B*256, B*256+256) of the
// compact list (B counted from the start of the arrays so that every bulk copy is 16-byte
// aligned); threads outside [first, first+count) idle.
//
// Phase 1 (COMP): one elected thread pulls the block's slice of the pull table (8 arrays of
// 8-bit row-rank offsets, the link words, the block's rank bases: 3.1 KB) into shared memory
// with TMA bulk copies on one mbarrier.  Doing this through registers instead lets ptxas sink
// each index load to its first use and chains up to 9 DRAM round trips; through shared memory
// it is exactly one, costs no registers, and 10 instructions per block.
// Phase 2: 19 gathers, then the same BC / macro / collide / store code as the dense kernel.
// A direction whose pull source is solid reads the node's own opposite population instead
// (half-way bounce-back); the node INDEX is selected, so either way it is one load (and, in
// place, one store) per direction and warp.
struct SparseTable {
    alignas(128) uint8_t rb8[8][SPARSE_BLOCK];     // 8-bit offsets of rb - i from the block's base; exception slots in row 0
    alignas(16) uint32_t fl[SPARSE_BLOCK];
    alignas(16) int32_t blk[16];
};
constexpr uint32_t kTableBytes = 8u * SPARSE_BLOCK + SPARSE_BLOCK * 4u + 64u;

// issue the bulk copies of table block `blk` into `tab`, completing on `bar`
__device__ __forceinline__ void table_fetch(const StepArgs &a, uint32_t blk, SparseTable &tab, uint64_t *bar) {
    const uint32_t base = blk * SPARSE_BLOCK;
    mbar_expect_tx(bar, kTableBytes);
#pragma unroll
    for (int k = 0; k < 8; ++k) bulk_g2s(&tab.rb8[k][0], a.rb8[k] + base, SPARSE_BLOCK, bar);
    bulk_g2s(&tab.fl[0], a.flags + base, SPARSE_BLOCK * 4u, bar);
    bulk_g2s(&tab.blk[0], a.blk + (size_t)blk * 16, 64u, bar);
}

// every thread of the block: wait for the slice (`parity`: how often the barrier has completed before, & 1)
__device__ __forceinline__ void table_wait(const StepArgs &a, uint32_t blk, SparseTable &tab, uint64_t *bar,
                                           uint32_t parity = 0u) {
    (void)a; (void)blk; (void)tab;
    mbar_wait(bar, parity);
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// neighbour-row ranks of stored node i (thread threadIdx.x of the block that fetched `tab`)
__device__ __forceinline__ void table_ranks(const SparseTable &tab, uint32_t i, int32_t (&rb)[8]) {
#pragma unroll
    for (int k = 0; k < 8; ++k) rb[k] = tab.blk[k] + (int32_t)i + (int32_t)tab.rb8[k][threadIdx.x];
}
__device__ __forceinline__ uint32_t table_exc_slot(const SparseTable &tab) {
    return (uint32_t)tab.blk[8] + tab.rb8[0][threadIdx.x];
}
