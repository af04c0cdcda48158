// Host-side construction of the sparse storage: compacted fluid-node list (ascending linear
// index, z fastest) and its compressed 18-neighbour pull table (lbm_kernels.cuh: 8 neighbour-row
// ranks as 8-bit offsets of rank - index from per-block bases + link word, exception table).  Shared by the
// single-phase and two-phase C-ABI layers; everything is in an anonymous namespace.
// Reference: Single_phase/LBM_3D_SinglePhase_Solver.py:36-44 (the pointer SNode tree this
// replaces), :247-257 (periodic_index), :259-268 (the push whose pull form the table encodes).
#pragma once
#include <cub/cub.cuh>
#include <thrust/iterator/transform_iterator.h>

#include <climits>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "lbm_geometry.cuh"

namespace {

// true pull sources of a fluid node: j[s-1] = compact index of i - e_s, or -1 (bounce)
__device__ __forceinline__ void true_sources(const GeoParams &g, const int8_t *solid, const uint32_t *rank,
                                             int x, int y, int z, int32_t (&j)[18]) {
    for (int s = 1; s < 19; ++s) {
        size_t src;
        j[s - 1] = (pull_source(g, x, y, z, s, src) && solid[src] == 0) ? (int32_t)rank[src] : -1;
    }
}

// Pass 1 of the sparse tables: linear index, link word (bounce bits + BC bits), and either
// the full 18-entry pull table or the 8 neighbour-row ranks of the compressed one (32-bit,
// temporary; nodes whose sources do not follow the rank rule are flagged).
__global__ void k_build_sparse(const GeoParams g, const int8_t *__restrict__ solid,
                               const uint32_t *__restrict__ rank, size_t stride,
                               uint32_t *__restrict__ lin, uint32_t *__restrict__ flags,
                               int32_t *__restrict__ nbr, int32_t *__restrict__ rb) {
    const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t N = (size_t)g.nx * g.ny * g.nz;
    if (idx >= N || solid[idx] != 0) return;
    const int z = (int)(idx % g.nz);
    const size_t t = idx / g.nz;
    const int y = (int)(t % g.ny), x = (int)(t / g.ny);
    const uint32_t r = rank[idx];
    lin[r] = (uint32_t)idx;
    int32_t j[18];
    true_sources(g, solid, rank, x, y, z, j);
    uint32_t fl = bc_word(g, solid, x, y, z);
    for (int s = 1; s < 19; ++s)
        if (j[s - 1] < 0) fl |= 1u << s;
    // two-phase: face bits (psi / velocity BCs, clamped psi stencil) and the wetting flag; the
    // single-phase word leaves bits 24..30 clear (wraps live in the table)
    if (g.two_phase) fl |= at_face_bits(g, x, y, z) | near_solid_bit(g, solid, x, y, z);
    else if (g.vel_in_place) fl |= at_face_bits(g, x, y, z);
    if (nbr != nullptr)
        for (int s = 1; s < 19; ++s) nbr[(size_t)(s - 1) * stride + r] = j[s - 1];
    if (rb != nullptr) {
        const int center[8] = {1, 2, 3, 4, 7, 8, 9, 10};     // directions (ex,ey,0) of the 8 rows
        int32_t rbv[8];
        for (int k = 0; k < 8; ++k) {
            size_t src;
            rbv[k] = pull_source(g, x, y, z, center[k], src) ? (int32_t)rank[src] : 0;
        }
        bool ok = true;
        const bool ghost = g.halo_x && (x == 0 || x == g.nx - 1);   // never updated
#define X(s, ex, ey, ez, o)                                                                    \
    if (s > 0 && j[s > 0 ? s - 1 : 0] >= 0 && comp_source<ex, ey, ez>(r, fl, rbv) != j[s > 0 ? s - 1 : 0]) ok = false;
        D3Q19_DIRS(X)
#undef X
        if (!ok && !ghost) fl |= FL_EXCEPTION;
        for (int k = 0; k < 8; ++k) rb[(size_t)k * stride + r] = rbv[k];
    }
    flags[r] = fl;
}

// Pass 2, one thread block per 256-node table block: 8-bit offsets of rank - index from the block's
// minimum.  A block in which one of them varies by more than 255 (a periodic x / y wrap falls inside
// it, or the neighbouring rows fill very unevenly) turns all its nodes into exceptions.  Exception nodes get consecutive slots from
// blk[B][8]; their index inside the block goes to rb16[0].  Nodes outside [own_first, own_end)
// (ghost planes of a slab) are never updated and do not take part.
__global__ void __launch_bounds__(256) k_pack_table(uint32_t own_first, uint32_t own_end, size_t stride,
                                                    const int32_t *__restrict__ rb32, uint32_t *__restrict__ flags,
                                                    uint8_t *__restrict__ rb8, int32_t *__restrict__ blk,
                                                    uint32_t *__restrict__ exc_count, uint32_t *__restrict__ wide_count) {
    __shared__ int s_dmin[8], s_dmax[8];
    __shared__ uint32_t s_warp[8], s_base;
    __shared__ int s_wide;
    const uint32_t B = blockIdx.x, t = threadIdx.x, i = B * 256u + t;
    const bool valid = i >= own_first && i < own_end;
    uint32_t fl = valid ? flags[i] : 0u;
    if (t < 8) { s_dmin[t] = INT32_MAX; s_dmax[t] = INT32_MIN; }
    __syncthreads();
    int32_t v[8];
    for (int k = 0; k < 8; ++k) {
        v[k] = valid ? rb32[(size_t)k * stride + i] : 0;
        const bool use = valid && !(fl & FL_EXCEPTION);
        // rb - i: how far the neighbour row's rank runs ahead of the node's own index
        int dlo = use ? v[k] - (int32_t)i : INT32_MAX, dhi = use ? v[k] - (int32_t)i : INT32_MIN;
        for (int o = 16; o > 0; o >>= 1) {
            dlo = min(dlo, __shfl_xor_sync(0xffffffffu, dlo, o));
            dhi = max(dhi, __shfl_xor_sync(0xffffffffu, dhi, o));
        }
        if ((t & 31) == 0) { atomicMin(&s_dmin[k], dlo); atomicMax(&s_dmax[k], dhi); }
    }
    __syncthreads();
    if (t == 0) {
        int wide = 0;
        for (int k = 0; k < 8; ++k)
            if (s_dmax[k] >= s_dmin[k] && (int64_t)s_dmax[k] - (int64_t)s_dmin[k] > 255) wide = 1;
        s_wide = wide;
        if (wide) atomicAdd(wide_count, 1u);
    }
    __syncthreads();
    if (s_wide && valid) fl |= FL_EXCEPTION;
    const bool exc = valid && (fl & FL_EXCEPTION);
    const uint32_t bal = __ballot_sync(0xffffffffu, exc);
    if ((t & 31) == 0) s_warp[t >> 5] = __popc(bal);
    __syncthreads();
    uint32_t before = __popc(bal & ((1u << (t & 31)) - 1u));
    uint32_t total = 0;
    for (int w = 0; w < 8; ++w) {
        if (w < (int)(t >> 5)) before += s_warp[w];
        total += s_warp[w];
    }
    if (t == 0) s_base = total ? atomicAdd(exc_count, total) : 0u;
    __syncthreads();
    if (t < 16) {
        int32_t b = 0;
        if (t < 8) {
            if (s_dmax[t] >= s_dmin[t] && !s_wide) b = s_dmin[t];
        } else if (t == 8) b = (int32_t)s_base;
        blk[(size_t)B * 16 + t] = b;
    }
    if (i < stride) {
        for (int k = 0; k < 8; ++k) {
            uint8_t w8 = 0;
            if (exc) { if (k == 0) w8 = (uint8_t)before; }           // slot inside the block, < 256
            else if (valid) w8 = (uint8_t)(v[k] - (int32_t)i - s_dmin[k]);
            rb8[(size_t)k * stride + i] = w8;
        }
        if (valid) flags[i] = fl;
    }
}

// Pass 3: explicit sources of the exception nodes
__global__ void k_fill_exceptions(const GeoParams g, const int8_t *__restrict__ solid,
                                  const uint32_t *__restrict__ rank, uint32_t nf, size_t stride,
                                  const uint32_t *__restrict__ lin, const uint32_t *__restrict__ flags,
                                  const uint8_t *__restrict__ rb8, const int32_t *__restrict__ blk,
                                  int32_t *__restrict__ exc, size_t exc_stride) {
    const uint32_t r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= nf || !(flags[r] & FL_EXCEPTION)) return;
    const size_t idx = lin[r];
    const int z = (int)(idx % g.nz);
    const size_t t = idx / g.nz;
    const int y = (int)(t % g.ny), x = (int)(t / g.ny);
    int32_t j[18];
    true_sources(g, solid, rank, x, y, z, j);
    const uint32_t slot = (uint32_t)blk[(size_t)(r / 256u) * 16 + 8] + rb8[r];
    for (int s = 0; s < 18; ++s) exc[(size_t)s * exc_stride + slot] = j[s];
}

struct SparseTables {
    size_t nf = 0, stride = 0, n_exc = 0, exc_stride = 0;
    uint32_t own_first = 0, own_count = 0;                 // node range a step updates
    uint32_t plane_first[4] = {0, 0, 0, 0}, plane_count[4] = {0, 0, 0, 0};   // halo planes (x-slab)
    uint32_t *d_rank = nullptr;    // [N+1] exclusive fluid count (freed once the tables exist)
    uint32_t *d_lin = nullptr;     // [stride] linear index of each stored node
    uint32_t *d_flags = nullptr;   // [stride] link word
    int32_t *d_nbr = nullptr;      // full table [18][stride] (only when !compressed)
    uint8_t *d_rb8 = nullptr;      // [8][stride]  8-bit offsets of rank - index; exception slots in row 0
    size_t n_wide = 0;             // table blocks turned into exceptions (a lead varies by more than 255)
    int32_t *d_blk = nullptr;      // [stride/256][16]
    int32_t *d_exc = nullptr;      // [18][exc_stride]
    std::vector<uint32_t> plane_rank;   // rank at the start of every x plane, [nx+1] (host)
    int launches = 0;
};

inline void free_sparse_tables(SparseTables &t) {
    cudaFree(t.d_rank); cudaFree(t.d_lin); cudaFree(t.d_flags); cudaFree(t.d_nbr);
    cudaFree(t.d_rb8); cudaFree(t.d_blk); cudaFree(t.d_exc);
    t = SparseTables();
}

// On failure the caller frees `t`.  `msg` is set for errors that are not CUDA errors.
inline cudaError_t build_sparse_tables(const GeoParams &g, const int8_t *d_solid, bool compressed,
                                       SparseTables &t, std::string &msg) {
#define SB_CU(call)                                                                            \
    do {                                                                                       \
        cudaError_t _e = (call);                                                               \
        if (_e != cudaSuccess) return _e;                                                      \
    } while (0)
    const int nx = g.nx;
    const size_t plane = (size_t)g.ny * g.nz, N = plane * nx;
    SB_CU(cudaMalloc(&t.d_rank, (N + 1) * sizeof(uint32_t)));
    auto it = thrust::make_transform_iterator(d_solid, IsFluid());
    size_t tmp_bytes = 0;
    void *tmp = nullptr;
    SB_CU(cub::DeviceScan::ExclusiveSum(nullptr, tmp_bytes, it, t.d_rank, N));
    SB_CU(cudaMalloc(&tmp, tmp_bytes));
    cudaError_t e = cub::DeviceScan::ExclusiveSum(tmp, tmp_bytes, it, t.d_rank, N);
    cudaError_t e2 = cudaDeviceSynchronize();
    cudaFree(tmp);
    SB_CU(e);
    SB_CU(e2);
    t.launches += 2;
    uint32_t last_rank = 0;
    int8_t last_solid = 1;
    SB_CU(cudaMemcpy(&last_rank, t.d_rank + (N - 1), sizeof(uint32_t), cudaMemcpyDeviceToHost));
    SB_CU(cudaMemcpy(&last_solid, d_solid + (N - 1), 1, cudaMemcpyDeviceToHost));
    t.nf = (size_t)last_rank + (last_solid == 0 ? 1 : 0);
    const uint32_t nf32 = (uint32_t)t.nf;
    SB_CU(cudaMemcpy(t.d_rank + N, &nf32, sizeof(uint32_t), cudaMemcpyHostToDevice));
    // planes padded to the 256-node block: the sparse kernel bulk-copies whole table slices
    t.stride = (t.nf + 255) / 256 * 256;
    if (t.stride == 0) t.stride = 256;
    // the step kernel reaches the opposite population of a node as index i +- stride (int32)
    if (t.stride >= ((size_t)1 << 30)) {
        char b[200];
        snprintf(b, sizeof b, "sparse storage holds at most 2^30 fluid nodes per context (got %zu): split the domain into x-slabs", t.nf);
        msg = b;
        return cudaSuccess;
    }
    SB_CU(cudaMalloc(&t.d_lin, t.stride * sizeof(uint32_t)));
    SB_CU(cudaMalloc(&t.d_flags, t.stride * sizeof(uint32_t)));
    SB_CU(cudaMemset(t.d_flags, 0, t.stride * sizeof(uint32_t)));
    SB_CU(cudaMemset(t.d_lin, 0, t.stride * sizeof(uint32_t)));
    uint32_t *d_cnt = nullptr;
    SB_CU(cudaMalloc(&d_cnt, sizeof(uint32_t)));
    e = cudaMemset(d_cnt, 0, sizeof(uint32_t));
    if (e != cudaSuccess) { cudaFree(d_cnt); return e; }
    t.plane_rank.assign((size_t)nx + 1, 0);
    e = cudaMemcpy2D(t.plane_rank.data(), sizeof(uint32_t), t.d_rank, plane * sizeof(uint32_t),
                     sizeof(uint32_t), (size_t)nx + 1, cudaMemcpyDeviceToHost);
    if (e != cudaSuccess) { cudaFree(d_cnt); return e; }
    if (g.halo_x) {
        const uint32_t r[4] = {t.plane_rank[1], t.plane_rank[2], t.plane_rank[nx - 2], t.plane_rank[nx - 1]};
        t.own_first = r[0];
        t.own_count = r[3] - r[0];
        t.plane_first[0] = 0; t.plane_count[0] = r[0];
        t.plane_first[1] = r[0]; t.plane_count[1] = r[1] - r[0];
        t.plane_first[2] = r[2]; t.plane_count[2] = r[3] - r[2];
        t.plane_first[3] = r[3]; t.plane_count[3] = nf32 - r[3];
    } else {
        t.own_first = 0;
        t.own_count = nf32;
    }
    int32_t *d_rb32 = nullptr;     // 32-bit neighbour-row ranks, only while the table is built
    if (compressed) {
        e = cudaMalloc(&d_rb32, t.stride * 8 * sizeof(int32_t));
        if (e == cudaSuccess) e = cudaMemset(d_rb32, 0, t.stride * 8 * sizeof(int32_t));
    } else {
        e = cudaMalloc(&t.d_nbr, t.stride * 18 * sizeof(int32_t));
        if (e == cudaSuccess) e = cudaMemset(t.d_nbr, 0xff, t.stride * 18 * sizeof(int32_t));
    }
    if (e == cudaSuccess) {
        k_build_sparse<<<nblocks(N, 256), 256>>>(g, d_solid, t.d_rank, t.stride, t.d_lin, t.d_flags, t.d_nbr, d_rb32);
        e = cudaGetLastError();
        t.launches++;
    }
    if (e == cudaSuccess && compressed) {
        const unsigned nblk = (unsigned)(t.stride / 256);
        e = cudaMalloc(&t.d_rb8, t.stride * 8);
        if (e == cudaSuccess) e = cudaMalloc(&t.d_blk, (size_t)nblk * 16 * sizeof(int32_t));
        uint32_t *d_narrow = nullptr;
        if (e == cudaSuccess) e = cudaMalloc(&d_narrow, sizeof(uint32_t));
        if (e == cudaSuccess) e = cudaMemset(d_narrow, 0, sizeof(uint32_t));
        if (e == cudaSuccess) {
            k_pack_table<<<nblk, 256>>>(t.own_first, t.own_first + t.own_count, t.stride, d_rb32, t.d_flags,
                                        t.d_rb8, t.d_blk, d_cnt, d_narrow);
            e = cudaGetLastError();
            t.launches++;
        }
        if (e == cudaSuccess) e = cudaDeviceSynchronize();
        uint32_t ne = 0, nn = 0;
        if (e == cudaSuccess) e = cudaMemcpy(&ne, d_cnt, sizeof ne, cudaMemcpyDeviceToHost);
        if (e == cudaSuccess) e = cudaMemcpy(&nn, d_narrow, sizeof nn, cudaMemcpyDeviceToHost);
        cudaFree(d_narrow);
        t.n_wide = nn;
        if (e == cudaSuccess) {
            t.n_exc = ne;
            t.exc_stride = (ne + 31) / 32 * 32 + 32;
            e = cudaMalloc(&t.d_exc, t.exc_stride * 18 * sizeof(int32_t));
        }
        if (e == cudaSuccess) e = cudaMemset(t.d_exc, 0xff, t.exc_stride * 18 * sizeof(int32_t));
        if (e == cudaSuccess && ne) {
            k_fill_exceptions<<<nblocks(t.nf, 256), 256>>>(g, d_solid, t.d_rank, nf32, t.stride, t.d_lin, t.d_flags,
                                                           t.d_rb8, t.d_blk, t.d_exc, t.exc_stride);
            e = cudaGetLastError();
            t.launches++;
        }
    }
    if (e == cudaSuccess) e = cudaDeviceSynchronize();
    cudaFree(d_rb32);
    cudaFree(d_cnt);
    cudaFree(t.d_rank);            // 4 B per LATTICE node, only needed while the tables are built
    t.d_rank = nullptr;
    return e;
#undef SB_CU
}

}  // namespace
