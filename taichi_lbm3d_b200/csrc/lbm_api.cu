// C-ABI layer of the single-phase solver (include/lbm3d.h): context, geometry
// preprocessing (link flags / fluid-node compaction + neighbour table), the state
// machine between the reference's user-visible state (F, rho, v) and the fused
// pipeline state (post-collision f*), getters/setters, halo staging.
//
// Reference: Single_phase/LBM_3D_SinglePhase_Solver.py (line numbers below).
#include <cub/cub.cuh>
#include <thrust/iterator/transform_iterator.h>
#include <cuda_runtime.h>

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/lbm3d.h"
#include "lbm_kernels.cuh"
#include "lbm_geometry.cuh"
#include "lbm_nccl.cuh"
#include "lbm_devpool.cuh"
#include "lbm_sparse_build.cuh"

namespace {

thread_local std::string g_create_error;

struct Face {
    int type = 0;
    float rho = 1.0f;
    float vel[3] = {0.f, 0.f, 0.f};
};

}  // namespace

struct lbm_ctx {
    lbm_config cfg{};
    std::string err;
    int block = 256;
    // parameters
    float S[19]{};
    float force[3] = {0.f, 0.f, 0.f};
    int guo_unscaled = 0;
    int vel_bc_script = 0;
    Face face[6];
    float invM[361]{};
    bool have_geometry = false;
    bool inited = false;
    // sizes
    size_t N = 0;          // nx*ny*nz
    size_t nf = 0;         // stored nodes (dense: N; sparse: fluid nodes incl. ghost planes)
    size_t stride = 0;     // population plane stride (elements)
    uint32_t own_first = 0, own_count = 0;   // sparse: node range updated by a step
    uint32_t row_first = 0, row_count = 0;   // dense: z-row range updated by a step
    int layout = 0;        // dense: 0 = SoA planes [19][N]; 1 = row-blocked [row][19][nzp]
    uint32_t nzp = 0;      // nz rounded up to 32 (row-blocked layout)
    uint32_t prow = 0;     // elements between consecutive z-rows inside a population plane
    size_t fsize = 0;      // elements of one population buffer (without guard bands)
    int spec = 1;          // speculative pull (dense, few solid nodes)
    uint32_t plane_first[4] = {0, 0, 0, 0}, plane_count[4] = {0, 0, 0, 0};  // halo planes
    int xface0 = -1, xface1 = -1;            // local x index of the global x0 / x1 faces
    // device buffers
    int8_t *d_solid = nullptr;
    float *d_ns = nullptr;         // grey-scale lattice: [N] solid fraction per node (kept across lbm_init, like d_solid)
    uint32_t *d_flags = nullptr;   // dense: [N] link words; sparse: [nf] BC words
    int32_t *d_nbr = nullptr;      // sparse, full table: [18][stride]
    uint8_t *d_rb8 = nullptr;      // sparse, compressed table: [8][stride] 8-bit offsets of rank - index; exception slots
    size_t n_wide = 0;             // table blocks turned into exceptions
    int32_t *d_blk = nullptr;      // sparse, compressed table: [stride/256][16] rank bases + exception-slot base
    int32_t *d_exc = nullptr;      // sparse, compressed table: [18][exc_stride] explicit sources
    size_t n_exc = 0, exc_stride = 0;
    int compressed = 1;
    uint32_t prefetch_dist = 0;    // sparse: table prefetch distance in nodes (multiple of the block)
    uint32_t *d_lin = nullptr;     // sparse: [nf]
    uint32_t *d_rank = nullptr;    // sparse: [N+1] exclusive fluid count
    std::vector<uint32_t> plane_rank;   // sparse: rank at the start of every x plane, [nx+1] (host)
    uint8_t *d_cls = nullptr;      // dense: [N] node class (NODE_BULK / NODE_SOLID / NODE_SPECIAL)
    float *d_fbase[2] = {nullptr, nullptr};   // allocations; d_f = d_fbase + pad
    size_t pad = 0;                // guard elements before/after the populations (speculative pull)
    float *d_f[2] = {nullptr, nullptr};
    float *d_rho = nullptr, *d_v = nullptr, *d_F = nullptr;
    float *d_vbc = nullptr;
    uint32_t vbc_off[6]{};
    float *d_scalar = nullptr;
    float *d_ff = nullptr;         // per-node force, [3][ff_stride] in stored order (null: uniform force)
    float *d_ffm = nullptr;        // the array the PENDING step was collided with (valid while ffm_pending)
    bool ffm_pending = false;
    size_t ff_stride = 0;
    // state machine
    bool dense_aa = false;       // cfg.sparse == 3: dense lattice stepped in place
    bool aa = false;             // in-place (AA-pattern) stepping on one buffer
    int parity = 0;              // aa: 0 = natural layout, 1 = arrival layout (see lbm_kernels.cuh)
    int cur = 0;
    bool pipe_valid = false;     // d_f[cur] holds f* of the current step
    bool macro_valid = true;     // d_rho / d_v current
    bool F_valid = true;         // d_F current (or implicit w when d_F == nullptr)
    cudaStream_t stream = nullptr;
    int64_t launches = 0;
    // multi-GPU slab runner (lbm_comm_init / lbm_run_slab)
    void *comm = nullptr;            // ncclComm_t
    int comm_world = 1, comm_rank = 0;
    cudaStream_t comm_stream = nullptr;
    cudaEvent_t ev_boundary = nullptr, ev_comm = nullptr, ev_interior = nullptr;
    float *d_send[2] = {nullptr, nullptr}, *d_recv[2] = {nullptr, nullptr};
    // direct peer-memory halo (lbm_p2p_*): this rank's flag words, the neighbours' mapped buffers
    int *d_p2p = nullptr;                    // [0],[1] launches completed by the left / right neighbour, [2] blocks done, [3] timeout
    void *peer_map[2][3] = {{nullptr, nullptr, nullptr}, {nullptr, nullptr, nullptr}};   // opened IPC handles (side; f0, f1, flags)
    bool peer_shared = false;                // left and right neighbour are the same rank (world of two)
    float *peer_f[2][2] = {{nullptr, nullptr}, {nullptr, nullptr}};   // neighbour's population buffers (pad applied)
    int *peer_flags[2] = {nullptr, nullptr};
    long long peer_pstride[2] = {0, 0}, peer_delta[2] = {0, 0};
    bool p2p_ready = false, p2p_on = false;
    int p2p_launches = 0;
};

struct P2pBlob {                             // what lbm_p2p_export writes (<= LBM_P2P_BLOB_BYTES)
    cudaIpcMemHandle_t f[2], flags;
    int32_t nx, ny, nz, layout;
    uint32_t prow;
    uint64_t pad, pstride;
};
static_assert(sizeof(P2pBlob) <= LBM_P2P_BLOB_BYTES, "blob too large");

#define CTX_CHECK(ctx)                                                                         \
    if ((ctx) == nullptr) return LBM_ERR_INVALID;
#define FAIL(ctx, code, ...)                                                                   \
    do {                                                                                       \
        char _b[512];                                                                          \
        snprintf(_b, sizeof _b, __VA_ARGS__);                                                  \
        (ctx)->err = _b;                                                                       \
        return (code);                                                                         \
    } while (0)
#define CU(ctx, call)                                                                          \
    do {                                                                                       \
        cudaError_t _e = (call);                                                               \
        if (_e != cudaSuccess) {                                                               \
            char _b[512];                                                                      \
            snprintf(_b, sizeof _b, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(_e),    \
                     __FILE__, __LINE__);                                                      \
            (ctx)->err = _b;                                                                   \
            return _e == cudaErrorMemoryAllocation ? LBM_ERR_NOMEM : LBM_ERR_CUDA;             \
        }                                                                                      \
    } while (0)

namespace {

// lbm_get_neighbor_table: expand whichever table the step kernel uses into [18][nf]
__global__ void k_decode_table(StepArgs a, uint32_t nf, int32_t *__restrict__ out) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nf) return;
    if (!a.compressed) {
        for (int s = 0; s < 18; ++s) out[(size_t)s * nf + i] = a.nbr[s][i];
        return;
    }
    const uint32_t fl = a.flags[i];
    const int32_t *blk = a.blk + (size_t)(i / 256u) * 16;
    int32_t rb[8];
    for (int k = 0; k < 8; ++k) rb[k] = blk[k] + (int32_t)i + (int32_t)a.rb8[k][i];
    const uint32_t slot = (uint32_t)blk[8] + a.rb8[0][i];
#define X(s, ex, ey, ez, o)                                                                    \
    if (s > 0) {                                                                               \
        int32_t j = -1;                                                                        \
        if (!((fl >> s) & 1u))                                                                 \
            j = (fl & FL_EXCEPTION) ? a.exc[s > 0 ? s - 1 : 0][slot] : comp_source<ex, ey, ez>(i, fl, rb); \
        out[(size_t)(s > 0 ? s - 1 : 0) * nf + i] = j;                                         \
    }
    D3Q19_DIRS(X)
#undef X
}

// peer-memory halo: wait until both neighbours have completed `need` boundary launches, i.e. until
// everything they store into this rank's ghost planes has landed (end of lbm_run_slab)
__global__ void k_p2p_drain(int *p2p, int need) {
    const long long t0 = clock64();
    for (;;) {
        int a, b;
        asm volatile("ld.acquire.sys.global.s32 %0, [%1];" : "=r"(a) : "l"(p2p) : "memory");
        asm volatile("ld.acquire.sys.global.s32 %0, [%1];" : "=r"(b) : "l"(p2p + 1) : "memory");
        if (a >= need && b >= need) break;
        if (clock64() - t0 > 4000000000ll) { atomicExch(p2p + 3, 1); break; }
        __nanosleep(128);
    }
}

// lbm_get_nodes: rows of a [N][width] array at selected nodes
__global__ void k_take_rows(const float *__restrict__ src, const int64_t *__restrict__ index, int64_t n, int width,
                            float *__restrict__ dst) {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n * width) return;
    const int64_t i = t / width;
    dst[t] = src[index[i] * width + (t - i * width)];
}

// user force array [N][3] (reference layout) -> three planes in stored order
__global__ void k_force_planes(const float *__restrict__ src, const uint32_t *__restrict__ lin, size_t n,
                               size_t stride, float *__restrict__ dst) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const size_t node = lin ? lin[i] : i;
    for (int k = 0; k < 3; ++k) dst[(size_t)k * stride + i] = src[node * 3 + k];
}

void default_relaxation(double niu, int textbook, float S[19]) {
    // init_simulation :126-131, same double arithmetic as the Python source
    const double tau_f = textbook ? 3.0 * niu + 0.5 : niu / 3.0 + 0.5;
    const double s_v = 1.0 / tau_f;
    const double s_other = 8.0 * (2.0 - s_v) / (8.0 - s_v);
    const double S64[19] = {0, s_v, s_v, 0, s_other, 0, s_other, 0, s_other, s_v, s_v, s_v, s_v,
                            s_v, s_v, s_v, s_other, s_other, s_other};
    for (int i = 0; i < 19; ++i) S[i] = (float)S64[i];
}

void free_device(lbm_ctx *c) {
    // the buffers go back to the process-wide cache (lbm_devpool.cuh)
    lbm_pool::release(c->d_flags); lbm_pool::release(c->d_fbase[0]); lbm_pool::release(c->d_fbase[1]);
    lbm_pool::release(c->d_rho); lbm_pool::release(c->d_v); lbm_pool::release(c->d_F);
    lbm_pool::release(c->d_solid); lbm_pool::release(c->d_nbr); lbm_pool::release(c->d_lin);
    lbm_pool::release(c->d_rank);
    lbm_pool::release(c->d_rb8); lbm_pool::release(c->d_blk); lbm_pool::release(c->d_exc);
    c->d_rb8 = nullptr; c->d_blk = nullptr; c->d_exc = nullptr;
    lbm_pool::release(c->d_cls); c->d_cls = nullptr; c->d_fbase[0] = c->d_fbase[1] = nullptr;
    lbm_pool::release(c->d_vbc); lbm_pool::release(c->d_scalar);
    lbm_pool::release(c->d_ff); c->d_ff = nullptr;
    lbm_pool::release(c->d_ffm); c->d_ffm = nullptr; c->ffm_pending = false;
    c->d_solid = nullptr; c->d_flags = nullptr; c->d_nbr = nullptr; c->d_lin = nullptr;
    c->d_rank = nullptr; c->d_f[0] = c->d_f[1] = nullptr; c->d_rho = c->d_v = c->d_F = nullptr;
    c->d_vbc = nullptr; c->d_scalar = nullptr;
}

// cudaMalloc through the cache (every release in this file goes through lbm_pool::release, which
// passes what it does not know on to cudaFree).  x-slab contexts too: a block that was exported to
// the neighbours (CUDA IPC) comes back only after lbm_p2p_disconnect on every rank has unmapped it
// (SlabSolver.close is collective), and exporting it again for the next context yields the same
// handle, which the neighbours may open again.
static cudaError_t big_alloc(const lbm_ctx *c, void **p, size_t bytes) {
    static const bool slabs_too = getenv("LBM3D_POOL_SLABS") ? atoi(getenv("LBM3D_POOL_SLABS")) != 0 : true;
    return (c->cfg.halo_x && !slabs_too) ? cudaMalloc(p, bytes) : lbm_pool::alloc(p, bytes);
}

void fill_args(const lbm_ctx *c, StepArgs &a) {
    memset(&a, 0, sizeof a);
    a.stride = c->stride;
    a.first = c->own_first;
    a.count = c->own_count;
    a.row_first = c->row_first;
    a.row_count = c->row_count;
    a.row_split = 0xFFFFFFFFu; a.row_skip = 0;
    a.nb1 = 0xFFFFFFFFu; a.first2 = 0; a.count2 = 0;
    a.prow = c->prow;
    a.spec = c->spec;
    a.nx = c->cfg.nx; a.ny = c->cfg.ny; a.nz = c->cfg.nz;
    a.halo_x = c->cfg.halo_x ? 1 : 0;
    a.flags = c->d_flags;
    a.cls = c->d_cls;
    for (int s = 0; s < 18; ++s) {
        a.nbr[s] = c->d_nbr ? c->d_nbr + (size_t)s * c->stride : nullptr;
        a.exc[s] = c->d_exc ? c->d_exc + (size_t)s * c->exc_stride : nullptr;
    }
    for (int k = 0; k < 8; ++k) a.rb8[k] = c->d_rb8 ? c->d_rb8 + (size_t)k * c->stride : nullptr;
    a.blk = c->d_blk;
    a.compressed = c->cfg.sparse ? c->compressed : 0;
    a.prefetch_dist = c->prefetch_dist;
    a.aa = AA_OFF;
    a.lin = c->d_lin;
    a.rho = c->d_rho; a.v = c->d_v; a.F = nullptr;
    a.vbc = c->d_vbc;
    for (int i = 0; i < 6; ++i) a.vbc_off[i] = c->vbc_off[i];
    a.force = (fabsf(c->force[0]) > 0.f || fabsf(c->force[1]) > 0.f || fabsf(c->force[2]) > 0.f) ? 1 : 0;
    if (c->d_ff) {
        a.force = 2;
        for (int k = 0; k < 3; ++k) a.ff[k] = c->d_ff + (size_t)k * c->ff_stride;
        if (c->ffm_pending) {
            a.force = 3;
            for (int k = 0; k < 3; ++k) a.ffm[k] = c->d_ffm + (size_t)k * c->ff_stride;
        }
    }
    a.has_bc = 0;
    a.ns = c->d_ns;
    for (int i = 0; i < 19; ++i) a.P.S[i] = c->S[i];
    for (int i = 0; i < 3; ++i) a.P.force[i] = c->force[i];
    {
        // Guo term = A/ga + B/gb with A from (e-v).f and B from (e.v)(e.f): (ga, gb) = (3, 9) in the
        // class (:236), (1, 1) in the script variants; closed-form coefficients per moment group
        const double ga = c->guo_unscaled ? 1.0 : 1.0 / 3.0, gb = c->guo_unscaled ? 1.0 : 1.0 / 9.0;
        a.P.guo_unscaled = c->guo_unscaled;
        a.P.gc[0] = (float)(-ga + gb / 3.0);
        a.P.gc[1] = (float)(2.0 * gb / 9.0);
        a.P.gc[2] = (float)(ga / 3.0);
        a.P.gc[3] = (float)(2.0 * gb / 9.0);
        a.P.gc[4] = (float)(gb / 9.0);
    }
    a.P.vel_bc_script = c->vel_bc_script;
    for (int i = 0; i < 6; ++i) {
        a.P.bc_type[i] = c->face[i].type;
        a.P.bc_rho[i] = c->face[i].rho;
        for (int k = 0; k < 3; ++k) a.P.bc_vel[i][k] = c->face[i].vel[k];
        if (c->face[i].type != 0) a.has_bc = 1;
    }
}

// plane pointers of the input / output population buffers
void set_buffers(const lbm_ctx *c, StepArgs &a, const float *fin, float *fout) {
    static const int e[19][3] = {{0, 0, 0}, {1, 0, 0}, {-1, 0, 0}, {0, 1, 0}, {0, -1, 0}, {0, 0, 1},
        {0, 0, -1}, {1, 1, 0}, {-1, -1, 0}, {1, -1, 0}, {-1, 1, 0}, {1, 0, 1}, {-1, 0, -1},
        {1, 0, -1}, {-1, 0, 1}, {0, 1, 1}, {0, -1, -1}, {0, 1, -1}, {0, -1, 1}};
    const long long sy = c->cfg.sparse ? 0 : (long long)c->prow, sx = (long long)c->cfg.ny * sy;
    const size_t pstride = (!c->cfg.sparse && c->layout == 1) ? (size_t)c->nzp : c->stride;
    for (int s = 0; s < 19; ++s) {
        a.pown[s] = fin ? fin + (size_t)s * pstride : nullptr;
        a.ppull[s] = fin ? a.pown[s] - (e[s][0] * sx + e[s][1] * sy + e[s][2]) : nullptr;
        a.pout[s] = fout ? fout + (size_t)s * pstride : nullptr;
    }
}

int launch(lbm_ctx *c, int mode, const StepArgs &a, cudaStream_t st) {
    cudaError_t e;
    if (c->cfg.strict)
        e = c->cfg.sparse ? lbm_strict::launch_sparse(mode, a, c->block, st)
                          : lbm_strict::launch_dense(mode, a, c->block, st);
    else
        e = c->cfg.sparse ? lbm_fast::launch_sparse(mode, a, c->block, st)
                          : lbm_fast::launch_dense(mode, a, c->block, st);
    if (e != cudaSuccess) FAIL(c, LBM_ERR_CUDA, "kernel launch failed: %s", cudaGetErrorString(e));
    if (c->cfg.sparse ? a.count : a.row_count) c->launches++;
    return LBM_OK;
}

// user-visible F array, allocated on first need; solid nodes hold w (:164-169)
int ensure_F(lbm_ctx *c) {
    if (c->d_F != nullptr) return LBM_OK;
    CU(c, big_alloc(c, (void **)&c->d_F, c->N * 19 * sizeof(float)));
    k_fill_weights<<<nblocks(c->N * 19, 256), 256, 0, c->stream>>>(c->d_F, c->N);
    CU(c, cudaGetLastError());
    c->launches++;
    // pristine state: F = w everywhere is already the truth; otherwise it must be extracted
    return LBM_OK;
}

// pipeline -> user-visible state (the streaming3 pass of the last step)
int sync_fields(lbm_ctx *c, bool need_F) {
    if (!c->inited) FAIL(c, LBM_ERR_STATE, "lbm_init has not been called");
    CU(c, cudaSetDevice(c->cfg.device));
    if (c->p2p_on) {                 // did a boundary kernel give up waiting for a neighbour?
        int timed_out = 0;
        CU(c, cudaStreamSynchronize(c->stream));
        CU(c, cudaMemcpy(&timed_out, c->d_p2p + 3, sizeof(int), cudaMemcpyDeviceToHost));
        if (timed_out) FAIL(c, LBM_ERR_STATE, "peer-memory halo: a neighbour rank did not arrive within the time-out; the state is invalid");
    }
    if (need_F) {
        const bool fresh = c->d_F == nullptr;
        int r = ensure_F(c);
        if (r) return r;
        if (fresh && c->pipe_valid) c->F_valid = false;
    }
    const bool want = (!c->macro_valid) || (need_F && !c->F_valid);
    if (want && c->pipe_valid) {
        StepArgs a;
        fill_args(c, a);
        set_buffers(c, a, c->d_f[c->cur], nullptr);
        a.F = need_F ? c->d_F : nullptr;
        if (c->aa && c->parity) a.aa = AA_EVEN;      // arrival layout: the streamed state is local
        if (a.force == 3) {                          // the pending step ran with the previous array
            a.force = 2;
            for (int k = 0; k < 3; ++k) a.ff[k] = a.ffm[k];
        }
        int r = launch(c, MODE_EXTRACT, a, c->stream);
        if (r) return r;
        c->macro_valid = true;
        if (need_F) c->F_valid = true;
    }
    return LBM_OK;
}

// close the pending step into the user-visible state (F, rho, v) and restart the pipeline from
// there: used when a parameter of the collision changes in a way a fused launch cannot bridge
int flush_pipeline(lbm_ctx *c) {
    if (!c->pipe_valid) return LBM_OK;
    int r = sync_fields(c, true);
    if (r) return r;
    c->pipe_valid = false;
    c->ffm_pending = false;
    return LBM_OK;
}

int copy_out(lbm_ctx *c, void *dst, const void *src, size_t bytes) {
    CU(c, cudaStreamSynchronize(c->stream));
    CU(c, cudaMemcpy(dst, src, bytes, cudaMemcpyDefault));
    return LBM_OK;
}

#define NC(ctx, call)                                                                          \
    do {                                                                                       \
        int _r = (call);                                                                       \
        if (_r != 0) FAIL(ctx, LBM_ERR_CUDA, "%s failed: %s", #call, g_nccl.GetErrorString(_r)); \
    } while (0)

int halo_pack_impl(lbm_ctx *c, int side, int which, float *dst, cudaStream_t st) {
    const int plane = side == 0 ? 1 : 2;
    const uint32_t cnt = c->plane_count[plane];
    if (cnt == 0) return LBM_OK;
    static const HaloDirs kR = {{1, 7, 9, 11, 13}}, kL = {{2, 8, 10, 12, 14}};
    StepArgs a;
    fill_args(c, a);
    set_buffers(c, a, c->d_f[which ? c->cur ^ 1 : c->cur], nullptr);
    if (c->cfg.sparse) a.nz = 0;
    k_halo_pack<<<nblocks(cnt, 256), 256, 0, st>>>(a, c->plane_first[plane], c->plane_first[plane], cnt,
                                                     side == 0 ? kL : kR, dst);
    CU(c, cudaGetLastError());
    c->launches++;
    return LBM_OK;
}

int halo_unpack_impl(lbm_ctx *c, int side, int which, const float *src, cudaStream_t st) {
    const int plane = side == 0 ? 0 : 3;
    const uint32_t cnt = c->plane_count[plane];
    if (cnt == 0) return LBM_OK;
    static const HaloDirs kR = {{1, 7, 9, 11, 13}}, kL = {{2, 8, 10, 12, 14}};
    StepArgs a;
    fill_args(c, a);
    set_buffers(c, a, nullptr, c->d_f[which ? c->cur ^ 1 : c->cur]);
    if (c->cfg.sparse) a.nz = 0;
    // the left ghost receives what the left neighbour sent to its right (e_x = +1) and v.v.
    k_halo_unpack<<<nblocks(cnt, 256), 256, 0, st>>>(a, c->plane_first[plane], c->plane_first[plane], cnt,
                                                       side == 0 ? kR : kL, src);
    CU(c, cudaGetLastError());
    c->launches++;
    return LBM_OK;
}

// one halo exchange of buffer `which` on stream st (see include/lbm3d.h)
int exchange_impl(lbm_ctx *c, int which, cudaStream_t st) {
    int r = halo_pack_impl(c, 0, which, c->d_send[0], st);
    if (r) return r;
    r = halo_pack_impl(c, 1, which, c->d_send[1], st);
    if (r) return r;
    const float *from_left = c->d_send[1], *from_right = c->d_send[0];   // ring of one slab
    if (c->comm_world > 1) {
        const int left = (c->comm_rank + c->comm_world - 1) % c->comm_world;
        const int right = (c->comm_rank + 1) % c->comm_world;
        NC(c, g_nccl.GroupStart());
        // posting order matters when left == right (two ranks): pairs match in order
        NC(c, g_nccl.Send(c->d_send[1], (size_t)5 * c->plane_count[2], 7, right, c->comm, st));
        NC(c, g_nccl.Send(c->d_send[0], (size_t)5 * c->plane_count[1], 7, left, c->comm, st));
        NC(c, g_nccl.Recv(c->d_recv[0], (size_t)5 * c->plane_count[0], 7, left, c->comm, st));
        NC(c, g_nccl.Recv(c->d_recv[1], (size_t)5 * c->plane_count[3], 7, right, c->comm, st));
        NC(c, g_nccl.GroupEnd());
        from_left = c->d_recv[0];
        from_right = c->d_recv[1];
    }
    r = halo_unpack_impl(c, 0, which, from_left, st);
    if (r) return r;
    return halo_unpack_impl(c, 1, which, from_right, st);
}

// LBM3D_TIMELINE=<file>: CUDA-event timestamps of the last steps of lbm_run_slab on both streams
// (there is no nsys in this image): per step, when the interior kernel, the boundary kernel and the
// pack / ncclSend+Recv / unpack of the exchange started and ended, in microseconds from the first
// recorded event.  Timed events serialise nothing, but they are only created when asked for.
struct Timeline {
    struct Mark { const char *what; int step; cudaEvent_t ev; };
    std::vector<Mark> marks;
    bool on = false;
    void mark(const char *what, int step, cudaStream_t st) {
        if (!on) return;
        cudaEvent_t ev;
        if (cudaEventCreate(&ev) != cudaSuccess) return;
        cudaEventRecord(ev, st);
        marks.push_back({what, step, ev});
    }
    void write(const char *path, int rank) {
        if (!on || marks.empty()) return;
        cudaDeviceSynchronize();
        std::string name = std::string(path) + (rank >= 0 ? ".rank" + std::to_string(rank) : "");
        FILE *fh = fopen(name.c_str(), "w");
        if (fh) fprintf(fh, "step,event,us_since_first\n");
        for (auto &m : marks) {
            float ms = 0.f;
            cudaEventElapsedTime(&ms, marks[0].ev, m.ev);
            if (fh) fprintf(fh, "%d,%s,%.1f\n", m.step, m.what, ms * 1e3f);
        }
        if (fh) fclose(fh);
        for (auto &m : marks) cudaEventDestroy(m.ev);
        marks.clear();
    }
};
Timeline g_timeline;

// the same exchange with one pack launch and one unpack launch for both sides (native slab loop)
int exchange_merged(lbm_ctx *c, int which, cudaStream_t st, int step = -1) {
    static const HaloDirs kR = {{1, 7, 9, 11, 13}}, kL = {{2, 8, 10, 12, 14}};
    float *buf = c->d_f[which ? c->cur ^ 1 : c->cur];
    StepArgs a;
    fill_args(c, a);
    if (c->cfg.sparse) a.nz = 0;
    const uint32_t *pf = c->plane_first, *pc = c->plane_count;
    const uint32_t most = pc[1] > pc[2] ? pc[1] : pc[2], most_g = pc[0] > pc[3] ? pc[0] : pc[3];
    if (most) {
        set_buffers(c, a, buf, nullptr);
        HaloSide s0{pf[1], pf[1], pc[1], kL, c->d_send[0]}, s1{pf[2], pf[2], pc[2], kR, c->d_send[1]};
        k_halo_pack2<<<dim3(nblocks(most, 256), 2), 256, 0, st>>>(a, s0, s1);
        CU(c, cudaGetLastError());
        c->launches++;
    }
    g_timeline.mark("pack_end", step, st);
    float *from_left = c->d_send[1], *from_right = c->d_send[0];   // ring of one slab
    if (c->comm_world > 1) {
        const int left = (c->comm_rank + c->comm_world - 1) % c->comm_world;
        const int right = (c->comm_rank + 1) % c->comm_world;
        NC(c, g_nccl.GroupStart());
        // posting order matters when left == right (two ranks): pairs match in order
        NC(c, g_nccl.Send(c->d_send[1], (size_t)5 * pc[2], 7, right, c->comm, st));
        NC(c, g_nccl.Send(c->d_send[0], (size_t)5 * pc[1], 7, left, c->comm, st));
        NC(c, g_nccl.Recv(c->d_recv[0], (size_t)5 * pc[0], 7, left, c->comm, st));
        NC(c, g_nccl.Recv(c->d_recv[1], (size_t)5 * pc[3], 7, right, c->comm, st));
        NC(c, g_nccl.GroupEnd());
        from_left = c->d_recv[0];
        from_right = c->d_recv[1];
    }
    g_timeline.mark("sendrecv_end", step, st);
    if (most_g) {
        set_buffers(c, a, nullptr, buf);
        // the left ghost receives what the left neighbour sent to its right (e_x = +1) and v.v.
        HaloSide g0{pf[0], pf[0], pc[0], kR, from_left}, g1{pf[3], pf[3], pc[3], kL, from_right};
        k_halo_unpack2<<<dim3(nblocks(most_g, 256), 2), 256, 0, st>>>(a, g0, g1);
        CU(c, cudaGetLastError());
        c->launches++;
    }
    g_timeline.mark("unpack_end", step, st);
    return LBM_OK;
}

// the boundary-plane kernel doubles as the halo exchange (k_dense_peer): dense slabs whose neighbours
// are mapped, uniform force (the per-node force variants have no peer form); the same on every rank
bool p2p_active(const lbm_ctx *c) { return c->p2p_on && c->d_ff == nullptr && !c->cfg.sparse; }

// first and last owned plane of a slab in one launch (buffer cur -> cur ^ 1)
int launch_boundary_planes(lbm_ctx *c, cudaStream_t st) {
    const int own = c->cfg.nx - 2;
    StepArgs a;
    fill_args(c, a);
    set_buffers(c, a, c->d_f[c->cur], c->d_f[c->cur ^ 1]);
    if (!c->cfg.sparse) {
        const uint32_t ny = (uint32_t)c->cfg.ny;
        a.row_first = ny;
        a.row_count = 2 * ny;
        a.row_split = ny;
        a.row_skip = ny * (uint32_t)(own - 2);
        if (p2p_active(c)) {
            static const int kL[5] = {2, 8, 10, 12, 14}, kR[5] = {1, 7, 9, 11, 13};
            const int out = c->cur ^ 1;             // every rank flips in lockstep
            for (int q = 0; q < 5; ++q) {
                a.peer_out[0][q] = c->peer_f[0][out] + (size_t)kL[q] * c->peer_pstride[0];
                a.peer_out[1][q] = c->peer_f[1][out] + (size_t)kR[q] * c->peer_pstride[1];
            }
            a.peer_delta[0] = c->peer_delta[0];
            a.peer_delta[1] = c->peer_delta[1];
            a.p2p = c->d_p2p;
            a.peer_flag[0] = c->peer_flags[0];
            a.peer_flag[1] = c->peer_flags[1];
            a.p2p_launch = c->p2p_launches++;
        }
    } else {
        a.first = c->plane_rank[1];
        a.count = c->plane_rank[2] - a.first;
        a.first2 = c->plane_rank[own];
        a.count2 = c->plane_rank[own + 1] - a.first2;
        if (a.count2 == 0) a.first2 = a.first;
        a.nb1 = a.count ? (a.first + a.count + 255u) / 256u - a.first / 256u : 0u;
    }
    return launch(c, MODE_STEP, a, st);
}

}  // namespace

extern "C" {

int lbm_abi_version(void) { return LBM3D_ABI_VERSION; }

const char *lbm_last_error(const lbm_ctx *ctx) {
    return ctx ? ctx->err.c_str() : g_create_error.c_str();
}

int lbm_create(const lbm_config *cfg, lbm_ctx **out) {
    if (!cfg || !out) { g_create_error = "null argument"; return LBM_ERR_INVALID; }
    *out = nullptr;
    if (cfg->nx < 1 || cfg->ny < 1 || cfg->nz < 1) { g_create_error = "extents must be >= 1"; return LBM_ERR_INVALID; }
    if (cfg->halo_x && cfg->nx < 3) { g_create_error = "halo_x needs nx >= 3 (two ghost planes)"; return LBM_ERR_INVALID; }
    const size_t N = (size_t)cfg->nx * cfg->ny * cfg->nz;
    if (N >= ((size_t)1 << 32)) { g_create_error = "lattice per context limited to 2^32 nodes"; return LBM_ERR_INVALID; }
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0) {
        g_create_error = std::string("no CUDA device available: ") + cudaGetErrorString(e) +
                         " (this library has no CPU fallback)";
        return LBM_ERR_CUDA;
    }
    if (cfg->device < 0 || cfg->device >= ndev) { g_create_error = "bad device ordinal"; return LBM_ERR_INVALID; }
    e = cudaSetDevice(cfg->device);
    if (e != cudaSuccess) { g_create_error = cudaGetErrorString(e); return LBM_ERR_CUDA; }
    lbm_ctx *c = new lbm_ctx();
    c->cfg = *cfg;
    if (cfg->sparse == 3) {          // dense storage, one population buffer
        c->dense_aa = true;
        c->cfg.sparse = 0;
    }
    c->N = N;
    if (const char *b = getenv("LBM3D_BLOCK")) {
        int v = atoi(b);
        if (v >= 32 && v <= 256 && v % 32 == 0) c->block = v;
    }
    default_relaxation(0.16667, 0, c->S);        // :18 niu default
    for (int i = 0; i < 19; ++i)
        for (int j = 0; j < 19; ++j) c->invM[i * 19 + j] = (float)kInvM[i][j];
    e = big_alloc(c, (void **)&c->d_solid, N);
    if (e == cudaSuccess) e = cudaMemset(c->d_solid, 0, N);
    if (e != cudaSuccess) {
        g_create_error = std::string("cudaMalloc(solid): ") + cudaGetErrorString(e);
        delete c;
        return LBM_ERR_NOMEM;
    }
    *out = c;
    return LBM_OK;
}

int lbm_p2p_disconnect(lbm_ctx *ctx);

int lbm_destroy(lbm_ctx *ctx) {
    CTX_CHECK(ctx);
    cudaSetDevice(ctx->cfg.device);
    cudaDeviceSynchronize();
    if (ctx->comm_stream) cudaStreamDestroy(ctx->comm_stream);
    if (ctx->ev_boundary) cudaEventDestroy(ctx->ev_boundary);
    if (ctx->ev_comm) cudaEventDestroy(ctx->ev_comm);
    if (ctx->ev_interior) cudaEventDestroy(ctx->ev_interior);
    for (int i = 0; i < 2; ++i) { lbm_pool::release(ctx->d_send[i]); lbm_pool::release(ctx->d_recv[i]); }
    lbm_p2p_disconnect(ctx);
    lbm_pool::release(ctx->d_p2p);
    lbm_pool::release(ctx->d_ns);
    free_device(ctx);
    delete ctx;
    return LBM_OK;
}

long long lbm_pool_trim(void) { return (long long)lbm_pool::trim(-1); }

int lbm_set_geometry(lbm_ctx *ctx, const int8_t *solid) {
    CTX_CHECK(ctx);
    if (!solid) FAIL(ctx, LBM_ERR_INVALID, "null geometry");
    if (ctx->inited) FAIL(ctx, LBM_ERR_STATE, "geometry must be set before lbm_init");
    CU(ctx, cudaSetDevice(ctx->cfg.device));
    CU(ctx, cudaMemcpy(ctx->d_solid, solid, ctx->N, cudaMemcpyDefault));
    k_binarize<<<nblocks(ctx->N, 256), 256>>>(ctx->d_solid, ctx->N);
    CU(ctx, cudaGetLastError());
    ctx->launches++;
    ctx->have_geometry = true;
    return LBM_OK;
}

int lbm_set_bc(lbm_ctx *ctx, int face, int type, float rho, const float vel[3]) {
    CTX_CHECK(ctx);
    if (face < 0 || face > 5 || type < 0 || type > 2) FAIL(ctx, LBM_ERR_INVALID, "bad face/type");
    if (ctx->inited) FAIL(ctx, LBM_ERR_STATE, "boundary conditions are fixed at lbm_init (the reference bakes them into its kernels at first launch)");
    ctx->face[face].type = type;
    if (type == 1) ctx->face[face].rho = rho;
    if (type == 2 && vel) for (int k = 0; k < 3; ++k) ctx->face[face].vel[k] = vel[k];
    return LBM_OK;
}

int lbm_set_force(lbm_ctx *ctx, const float force[3]) {
    CTX_CHECK(ctx);
    if (!force) FAIL(ctx, LBM_ERR_INVALID, "null force");
    const bool changed = force[0] != ctx->force[0] || force[1] != ctx->force[1] || force[2] != ctx->force[2];
    if (ctx->inited && ctx->pipe_valid && changed) {
        // the pending step was collided with the old force and must be closed with it (:385-388)
        int r = flush_pipeline(ctx);
        if (r) return r;
    }
    for (int k = 0; k < 3; ++k) ctx->force[k] = force[k];
    return LBM_OK;
}

int lbm_set_guo_form(lbm_ctx *ctx, int unscaled) {
    CTX_CHECK(ctx);
    if (ctx->inited) FAIL(ctx, LBM_ERR_STATE, "the form of the force term is fixed at lbm_init");
    ctx->guo_unscaled = unscaled ? 1 : 0;
    return LBM_OK;
}

int lbm_set_vel_bc_form(lbm_ctx *ctx, int script_form) {
    CTX_CHECK(ctx);
    if (ctx->inited) FAIL(ctx, LBM_ERR_STATE, "the form of the velocity faces is fixed at lbm_init");
    ctx->vel_bc_script = script_form ? 1 : 0;
    return LBM_OK;
}

int lbm_set_grey_scale(lbm_ctx *ctx, const float *ns) {
    CTX_CHECK(ctx);
    if (ctx->inited) FAIL(ctx, LBM_ERR_STATE, "the solid fractions are fixed at lbm_init");
    if (ns && (ctx->cfg.sparse || ctx->dense_aa || ctx->cfg.halo_x))
        FAIL(ctx, LBM_ERR_INVALID, "the grey-scale lattice needs dense two-buffer storage on one GPU "
                                   "(not sparse, not in place, not an x-slab)");
    CU(ctx, cudaSetDevice(ctx->cfg.device));
    if (!ns) {
        lbm_pool::release(ctx->d_ns);
        ctx->d_ns = nullptr;
        return LBM_OK;
    }
    if (!ctx->d_ns) CU(ctx, big_alloc(ctx, (void **)&ctx->d_ns, ctx->N * sizeof(float)));
    CU(ctx, cudaMemcpy(ctx->d_ns, ns, ctx->N * sizeof(float), cudaMemcpyDefault));
    return LBM_OK;
}

int lbm_set_force_field(lbm_ctx *c, const float *force3) {
    CTX_CHECK(c);
    if (!c->inited) FAIL(c, LBM_ERR_STATE, "the force array is set after lbm_init (it is stored in node order)");
    CU(c, cudaSetDevice(c->cfg.device));
    CU(c, cudaStreamSynchronize(c->stream));
    // A fused launch closes step k (macro with the force of step k) and opens step k+1 (collision
    // with the new force).  Array -> array: keep the old planes as `ffm` for that one launch.
    // Uniform <-> array: close the pending step first and restart from the user-visible state.
    if (c->pipe_valid && (force3 == nullptr || c->d_ff == nullptr)) {
        int r = flush_pipeline(c);
        if (r) return r;
    }
    if (!force3) {                       // back to the uniform force of lbm_set_force
        lbm_pool::release(c->d_ff); lbm_pool::release(c->d_ffm);
        c->d_ff = c->d_ffm = nullptr;
        c->ffm_pending = false;
        return LBM_OK;
    }
    const size_t n = c->cfg.sparse ? c->nf : c->N;
    c->ff_stride = c->stride;
    const size_t bytes = 3 * c->ff_stride * sizeof(float);
    if (c->d_ff && c->pipe_valid && !c->ffm_pending) {
        float *t = c->d_ffm;             // the current array becomes the pending step's
        c->d_ffm = c->d_ff;
        c->d_ff = t;
        c->ffm_pending = true;
    }
    if (!c->d_ff) CU(c, big_alloc(c, (void **)&c->d_ff, bytes));
    // stage the caller's array on the device if it lives on the host
    cudaPointerAttributes at{};
    const bool dev = cudaPointerGetAttributes(&at, force3) == cudaSuccess && at.type == cudaMemoryTypeDevice;
    cudaGetLastError();
    const float *src = force3;
    float *tmp = nullptr;
    if (!dev) {
        CU(c, big_alloc(c, (void **)&tmp, c->N * 3 * sizeof(float)));
        cudaError_t e = cudaMemcpy(tmp, force3, c->N * 3 * sizeof(float), cudaMemcpyHostToDevice);
        if (e != cudaSuccess) { lbm_pool::release(tmp); CU(c, e); }
        src = tmp;
    }
    cudaError_t e = cudaMemset(c->d_ff, 0, bytes);
    if (e == cudaSuccess && n) {
        k_force_planes<<<nblocks(n, 256), 256>>>(src, c->cfg.sparse ? c->d_lin : nullptr, n, c->ff_stride, c->d_ff);
        e = cudaGetLastError();
    }
    if (e == cudaSuccess) e = cudaDeviceSynchronize();
    lbm_pool::release(tmp);
    CU(c, e);
    c->launches++;
    return LBM_OK;
}

int lbm_set_viscosity(lbm_ctx *ctx, double niu, int textbook_tau) {
    CTX_CHECK(ctx);
    default_relaxation(niu, textbook_tau, ctx->S);
    return LBM_OK;
}

int lbm_set_relaxation(lbm_ctx *ctx, const float S[19]) {
    CTX_CHECK(ctx);
    if (!S) FAIL(ctx, LBM_ERR_INVALID, "null S");
    for (int i = 0; i < 19; ++i) ctx->S[i] = S[i];
    return LBM_OK;
}

int lbm_set_inverse_matrix(lbm_ctx *ctx, const float invM[361]) {
    CTX_CHECK(ctx);
    if (!invM) FAIL(ctx, LBM_ERR_INVALID, "null matrix");
    memcpy(ctx->invM, invM, sizeof ctx->invM);
    if (ctx->inited && ctx->cfg.strict) {
        CU(ctx, cudaSetDevice(ctx->cfg.device));
        CU(ctx, lbm_strict::set_inverse_matrix(ctx->invM));
    }
    return LBM_OK;
}

int lbm_init(lbm_ctx *c) {
    CTX_CHECK(c);
    CU(c, cudaSetDevice(c->cfg.device));
    const int nx = c->cfg.nx, ny = c->cfg.ny, nz = c->cfg.nz;
    const size_t N = c->N;
    const size_t plane = (size_t)ny * nz;
    // release everything but the geometry (re-init allowed)
    {
        int8_t *keep = c->d_solid;
        c->d_solid = nullptr;
        free_device(c);
        c->d_solid = keep;
    }
    c->inited = false;
    if (c->d_ns) {
        // a link that leaves a solid node is never written by the grey-scale script and holds w[s];
        // its in-place velocity face reads such slots back (F[LR[s]]) and would need them stored
        for (int i = 0; i < 6; ++i)
            if (c->face[i].type == 2)
                FAIL(c, LBM_ERR_INVALID, "fixed-velocity faces are not available on a grey-scale lattice "
                                         "(periodic and fixed-pressure faces are)");
    }
    GeoParams g;
    g.nx = nx; g.ny = ny; g.nz = nz;
    g.halo_x = c->cfg.halo_x ? 1 : 0;
    // local plane index of the global x faces; with ghost planes the faces are the first /
    // last OWNED planes of the slabs flagged in cfg.x_face_mask (bit0: holds x0, bit1: holds x1)
    if (g.halo_x) {
        c->xface0 = (c->cfg.x_face_mask & 1) ? 1 : -1;
        c->xface1 = (c->cfg.x_face_mask & 2) ? nx - 2 : -1;
    } else {
        c->xface0 = 0;
        c->xface1 = nx - 1;
    }
    g.xface0 = c->xface0; g.xface1 = c->xface1;
    for (int i = 0; i < 6; ++i) g.bc_type[i] = c->face[i].type;
    g.two_phase = 0;
    g.vel_in_place = c->vel_bc_script;
    for (int i = 0; i < 6; ++i) g.bc_psi_type[i] = 0;

    CU(c, big_alloc(c, (void **)&c->d_scalar, 16));
    if (!c->cfg.sparse) {
        c->stride = (N + 31) / 32 * 32;
        CU(c, big_alloc(c, (void **)&c->d_flags, N * sizeof(uint32_t)));
        CU(c, big_alloc(c, (void **)&c->d_cls, N));
        k_build_flags<<<nblocks(N, 256), 256>>>(g, c->d_solid, c->d_flags, c->d_cls);
        CU(c, cudaGetLastError());
        c->launches++;
        // population layout: row-blocked [row][19][nzp] keeps the 19 populations of a z-row in
        // one contiguous 19*nzp*4-byte chunk (measured +4.5% DRAM throughput over SoA planes on
        // B200, scripts/microbench/streams.cu); SoA when the 32-bit element index would overflow
        c->nzp = (uint32_t)((nz + 31) / 32 * 32);
        const size_t rows = (size_t)nx * ny;
        c->layout = (rows * 19 * c->nzp + 2 * (size_t)(ny + 2) * 19 * c->nzp < ((size_t)1 << 32)) ? 1 : 0;
        if (const char *l = getenv("LBM3D_LAYOUT")) c->layout = atoi(l) ? 1 : 0;
        if (c->layout == 1) {
            c->prow = 19 * c->nzp;
            c->fsize = rows * c->prow;
        } else {
            c->prow = (uint32_t)nz;
            c->fsize = c->stride * 19;
        }
        // speculation pays when few nodes are solid (their speculative loads are wasted)
        {
            int64_t nfl = 0;
            auto it = thrust::make_transform_iterator((const int8_t *)c->d_solid, IsFluid());
            uint32_t *d_out = nullptr;
            CU(c, big_alloc(c, (void **)&d_out, sizeof(uint32_t)));
            size_t tmp_bytes = 0;
            void *tmp = nullptr;
            cub::DeviceReduce::Sum(nullptr, tmp_bytes, it, d_out, N);
            CU(c, big_alloc(c, (void **)&tmp, tmp_bytes));
            cudaError_t e = cub::DeviceReduce::Sum(tmp, tmp_bytes, it, d_out, N);
            uint32_t h = 0;
            cudaError_t e2 = cudaMemcpy(&h, d_out, sizeof h, cudaMemcpyDeviceToHost);
            lbm_pool::release(tmp); lbm_pool::release(d_out);
            CU(c, e); CU(c, e2);
            nfl = h;
            c->nf = (size_t)nfl;
            c->spec = (double)nfl >= 0.75 * (double)N ? 1 : 0;
            if (const char *sp = getenv("LBM3D_SPEC")) c->spec = atoi(sp) ? 1 : 0;
        }
        if (g.halo_x) {
            c->row_first = (uint32_t)ny;
            c->row_count = (uint32_t)(ny * (nx - 2));
            c->plane_first[0] = 0; c->plane_first[1] = (uint32_t)ny;
            c->plane_first[2] = (uint32_t)(ny * (nx - 2)); c->plane_first[3] = (uint32_t)(ny * (nx - 1));
            for (int i = 0; i < 4; ++i) c->plane_count[i] = (uint32_t)plane;
        } else {
            c->row_first = 0;
            c->row_count = (uint32_t)(nx * ny);
        }
    } else {
        // compacted fluid-node list + compressed pull table (replaces the pointer SNode tree :36-44)
        c->compressed = 1;
        if (const char *tb = getenv("LBM3D_SPARSE_TABLE")) c->compressed = strcmp(tb, "full") == 0 ? 0 : 1;
        // L2 prefetch of the table slice this many nodes ahead: off by default (measured on B200:
        // with the 20-byte table the step runs at the DRAM copy rate without it, 87.0 % vs 85.4 %
        // of the 152-byte roofline with one wave = 148 x 8 x 256 nodes ahead)
        c->prefetch_dist = 0;
        if (const char *pd = getenv("LBM3D_PREFETCH")) c->prefetch_dist = (uint32_t)atol(pd) / 256u * 256u;
        SparseTables t;
        std::string msg;
        cudaError_t e = build_sparse_tables(g, c->d_solid, c->compressed != 0, t, msg);
        c->launches += t.launches;
        if (e != cudaSuccess || !msg.empty()) {
            free_sparse_tables(t);
            if (!msg.empty()) FAIL(c, LBM_ERR_INVALID, "%s", msg.c_str());
            CU(c, e);
        }
        // the context owns the tables from here on
        c->nf = t.nf; c->stride = t.stride; c->n_exc = t.n_exc; c->exc_stride = t.exc_stride;
        c->d_rank = t.d_rank; c->d_lin = t.d_lin; c->d_flags = t.d_flags; c->d_nbr = t.d_nbr;
        c->d_rb8 = t.d_rb8; c->d_blk = t.d_blk; c->d_exc = t.d_exc;
        c->n_wide = t.n_wide;
        if (getenv("LBM3D_DEBUG"))
            fprintf(stderr, "[lbm3d] sparse table: %zu nodes, %zu blocks, %zu of them all-exception, %zu exception nodes\n",
                    t.nf, t.stride / 256, t.n_wide, t.n_exc);
        c->plane_rank = t.plane_rank;
        c->own_first = t.own_first; c->own_count = t.own_count;
        for (int i = 0; i < 4; ++i) { c->plane_first[i] = t.plane_first[i]; c->plane_count[i] = t.plane_count[i]; }
    }
    // populations (A-B), user-visible macros, pressure-BC velocities
    // guard band: the dense kernel pulls speculatively from idx -/+ (plane + row + 1)
    if (c->cfg.sparse) {
        c->pad = 0;
        c->fsize = c->stride * 19;
        c->prow = 0;
    } else {
        c->pad = (((size_t)ny + 1) * c->prow + 2 + 31) / 32 * 32;
    }
    const size_t fbytes = (c->fsize + 2 * c->pad) * sizeof(float);
    c->aa = ((c->cfg.sparse == 2 && c->compressed) || c->dense_aa) && !c->cfg.halo_x;
    c->parity = 0;
    for (int b = 0; b < (c->aa ? 1 : 2); ++b) {
        CU(c, big_alloc(c, (void **)&c->d_fbase[b], fbytes));
        CU(c, cudaMemset(c->d_fbase[b], 0, fbytes));
        c->d_f[b] = c->d_fbase[b] + c->pad;
    }
    if (c->aa) c->d_f[1] = c->d_f[0];
    CU(c, big_alloc(c, (void **)&c->d_rho, N * sizeof(float)));
    CU(c, big_alloc(c, (void **)&c->d_v, N * 3 * sizeof(float)));
    k_fill<<<nblocks(N, 256), 256>>>(c->d_rho, N, 1.0f);      // init() :165
    CU(c, cudaGetLastError());
    c->launches++;
    CU(c, cudaMemset(c->d_v, 0, N * 3 * sizeof(float)));       // :166
    const size_t fs[6] = {plane, plane, (size_t)nx * nz, (size_t)nx * nz, (size_t)nx * ny, (size_t)nx * ny};
    size_t tot = 0;
    for (int i = 0; i < 6; ++i) { c->vbc_off[i] = (uint32_t)tot; tot += fs[i]; }
    CU(c, big_alloc(c, (void **)&c->d_vbc, tot * 3 * sizeof(float)));
    CU(c, cudaMemset(c->d_vbc, 0, tot * 3 * sizeof(float)));
    if (c->cfg.strict) CU(c, lbm_strict::set_inverse_matrix(c->invM));
    CU(c, cudaDeviceSynchronize());
    c->cur = 0;
    c->pipe_valid = false;
    c->macro_valid = true;
    c->F_valid = true;
    c->inited = true;
    return LBM_OK;
}

static int ensure_pipeline(lbm_ctx *c, cudaStream_t st) {
    if (c->pipe_valid) return LBM_OK;
    // first collision of the user-visible state (:222-241 with the stored rho, v)
    StepArgs a;
    fill_args(c, a);
    set_buffers(c, a, nullptr, c->d_f[c->cur]);
    a.F = c->d_F;      // null = pristine init state (F = w, rho = 1, v = 0)
    int r = launch(c, MODE_COLLIDE, a, st);
    if (r) return r;
    c->pipe_valid = true;
    c->parity = 0;               // the collision writes the natural layout
    return LBM_OK;
}

int lbm_step(lbm_ctx *c, int nsteps, void *cuda_stream) {
    CTX_CHECK(c);
    if (!c->inited) FAIL(c, LBM_ERR_STATE, "lbm_init has not been called");
    if (nsteps < 0) FAIL(c, LBM_ERR_INVALID, "nsteps < 0");
    if (nsteps == 0) return LBM_OK;
    CU(c, cudaSetDevice(c->cfg.device));
    cudaStream_t st = (cudaStream_t)cuda_stream;
    c->stream = st;
    // Reference step n = collide(F_{n-1}) ; stream ; BC ; macro.  The pipeline holds the
    // post-collision state, so the first step after (re)initialisation is the collision
    // alone and every later launch is [stream+BC+macro of step k] + [collision of step k+1];
    // the closing stream+BC+macro runs on demand in sync_fields().
    if (!c->pipe_valid) {
        int r = ensure_pipeline(c, st);
        if (r) return r;
        nsteps -= 1;
    }
    StepArgs a;
    fill_args(c, a);
    for (int it = 0; it < nsteps; ++it) {
        set_buffers(c, a, c->d_f[c->cur], c->d_f[c->cur ^ 1]);
        if (c->aa) a.aa = c->parity ? AA_EVEN : AA_ODD;
        int r = launch(c, MODE_STEP, a, st);
        if (r) return r;
        if (c->aa) c->parity ^= 1;
        else c->cur ^= 1;
        if (c->ffm_pending) {            // only the first launch after a new force array closes an old step
            c->ffm_pending = false;
            fill_args(c, a);
        }
    }
    c->macro_valid = false;
    c->F_valid = false;
    return LBM_OK;
}

int64_t lbm_launch_count(const lbm_ctx *ctx) { return ctx ? ctx->launches : -1; }

int lbm_synchronize(lbm_ctx *ctx) {
    CTX_CHECK(ctx);
    CU(ctx, cudaSetDevice(ctx->cfg.device));
    CU(ctx, cudaStreamSynchronize(ctx->stream));
    return LBM_OK;
}

int lbm_get_rho(lbm_ctx *c, float *dst) {
    CTX_CHECK(c);
    if (!dst) FAIL(c, LBM_ERR_INVALID, "null destination");
    int r = sync_fields(c, false);
    if (r) return r;
    return copy_out(c, dst, c->d_rho, c->N * sizeof(float));
}

int lbm_get_v(lbm_ctx *c, float *dst) {
    CTX_CHECK(c);
    if (!dst) FAIL(c, LBM_ERR_INVALID, "null destination");
    int r = sync_fields(c, false);
    if (r) return r;
    return copy_out(c, dst, c->d_v, c->N * 3 * sizeof(float));
}

int lbm_get_F(lbm_ctx *c, float *dst) {
    CTX_CHECK(c);
    if (!dst) FAIL(c, LBM_ERR_INVALID, "null destination");
    int r = sync_fields(c, true);
    if (r) return r;
    return copy_out(c, dst, c->d_F, c->N * 19 * sizeof(float));
}

int lbm_get_solid(lbm_ctx *c, int8_t *dst) {
    CTX_CHECK(c);
    if (!dst) FAIL(c, LBM_ERR_INVALID, "null destination");
    CU(c, cudaSetDevice(c->cfg.device));
    CU(c, cudaMemcpy(dst, c->d_solid, c->N, cudaMemcpyDefault));
    return LBM_OK;
}

static int set_field(lbm_ctx *c, const float *src, int which) {
    if (!src) FAIL(c, LBM_ERR_INVALID, "null source");
    int r = sync_fields(c, true);     // the other fields must be current before one is replaced
    if (r) return r;
    CU(c, cudaStreamSynchronize(c->stream));
    float *dst = which == 0 ? c->d_rho : (which == 1 ? c->d_v : c->d_F);
    const size_t n = which == 0 ? c->N : (which == 1 ? c->N * 3 : c->N * 19);
    CU(c, cudaMemcpy(dst, src, n * sizeof(float), cudaMemcpyDefault));
    c->pipe_valid = false;            // next step restarts from the user-visible state
    c->ffm_pending = false;
    c->macro_valid = true;
    c->F_valid = true;
    return LBM_OK;
}
int lbm_set_rho(lbm_ctx *c, const float *src) { CTX_CHECK(c); return set_field(c, src, 0); }
int lbm_set_v(lbm_ctx *c, const float *src) { CTX_CHECK(c); return set_field(c, src, 1); }
int lbm_set_F(lbm_ctx *c, const float *src) { CTX_CHECK(c); return set_field(c, src, 2); }

int lbm_get_max_v(lbm_ctx *c, float *out) {
    CTX_CHECK(c);
    if (!out) FAIL(c, LBM_ERR_INVALID, "null destination");
    int r = sync_fields(c, false);
    if (r) return r;
    CU(c, max_v_reduce(c->d_v, c->N, c->d_scalar, c->stream, out));
    c->launches++;
    return LBM_OK;
}

int lbm_get_nodes(lbm_ctx *c, int64_t n, const int64_t *index, float *F_out, float *rho_out, float *v_out) {
    CTX_CHECK(c);
    if (n < 0 || (n > 0 && !index)) FAIL(c, LBM_ERR_INVALID, "bad node list");
    int r = sync_fields(c, F_out != nullptr);
    if (r) return r;
    if (n == 0) return LBM_OK;
    CU(c, cudaStreamSynchronize(c->stream));
    std::vector<int64_t> host(n);
    CU(c, cudaMemcpy(host.data(), index, n * sizeof(int64_t), cudaMemcpyDefault));
    for (int64_t i = 0; i < n; ++i)
        if (host[i] < 0 || (size_t)host[i] >= c->N) FAIL(c, LBM_ERR_INVALID, "node index %lld outside the lattice", (long long)host[i]);
    int64_t *d_index = nullptr;
    float *d_tmp = nullptr;
    CU(c, big_alloc(c, (void **)&d_index, n * sizeof(int64_t)));
    cudaError_t e = cudaMemcpy(d_index, host.data(), n * sizeof(int64_t), cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = big_alloc(c, (void **)&d_tmp, (size_t)n * 19 * sizeof(float));
    const float *src[3] = {c->d_F, c->d_rho, c->d_v};
    float *dst[3] = {F_out, rho_out, v_out};
    const int width[3] = {19, 1, 3};
    for (int k = 0; k < 3 && e == cudaSuccess; ++k) {
        if (!dst[k]) continue;
        k_take_rows<<<nblocks((size_t)n * width[k], 256), 256>>>(src[k], d_index, n, width[k], d_tmp);
        e = cudaGetLastError();
        if (e == cudaSuccess) e = cudaMemcpy(dst[k], d_tmp, (size_t)n * width[k] * sizeof(float), cudaMemcpyDefault);
        c->launches++;
    }
    lbm_pool::release(d_index);
    lbm_pool::release(d_tmp);
    CU(c, e);
    return LBM_OK;
}

int lbm_get_num_fluid(lbm_ctx *c, int64_t *n) {
    CTX_CHECK(c);
    if (!n) FAIL(c, LBM_ERR_INVALID, "null destination");
    if (!c->inited) FAIL(c, LBM_ERR_STATE, "lbm_init has not been called");
    if (c->cfg.sparse) { *n = (int64_t)c->nf; return LBM_OK; }
    // dense: count on demand
    CU(c, cudaSetDevice(c->cfg.device));
    auto it = thrust::make_transform_iterator((const int8_t *)c->d_solid, IsFluid());
    uint32_t *d_out = nullptr;
    CU(c, big_alloc(c, (void **)&d_out, sizeof(uint32_t)));
    size_t tmp_bytes = 0;
    void *tmp = nullptr;
    cub::DeviceReduce::Sum(nullptr, tmp_bytes, it, d_out, c->N);
    CU(c, big_alloc(c, (void **)&tmp, tmp_bytes));
    cudaError_t e = cub::DeviceReduce::Sum(tmp, tmp_bytes, it, d_out, c->N);
    uint32_t h = 0;
    cudaError_t e2 = cudaMemcpy(&h, d_out, sizeof h, cudaMemcpyDeviceToHost);
    lbm_pool::release(tmp); lbm_pool::release(d_out);
    CU(c, e); CU(c, e2);
    *n = h;
    return LBM_OK;
}

int lbm_get_fluid_index(lbm_ctx *c, int64_t *dst) {
    CTX_CHECK(c);
    if (!dst) FAIL(c, LBM_ERR_INVALID, "null destination");
    if (!c->inited || !c->cfg.sparse) FAIL(c, LBM_ERR_STATE, "needs an initialised sparse context");
    CU(c, cudaSetDevice(c->cfg.device));
    // widen on the host side of the copy: fetch u32, convert
    std::string buf(c->nf * sizeof(uint32_t), '\0');
    CU(c, cudaMemcpy(&buf[0], c->d_lin, c->nf * sizeof(uint32_t), cudaMemcpyDeviceToHost));
    const uint32_t *u = reinterpret_cast<const uint32_t *>(buf.data());
    cudaPointerAttributes at{};
    const bool dev = cudaPointerGetAttributes(&at, dst) == cudaSuccess && at.type == cudaMemoryTypeDevice;
    cudaGetLastError();
    if (!dev) {
        for (size_t i = 0; i < c->nf; ++i) dst[i] = (int64_t)u[i];
    } else {
        std::string w(c->nf * sizeof(int64_t), '\0');
        int64_t *p = reinterpret_cast<int64_t *>(&w[0]);
        for (size_t i = 0; i < c->nf; ++i) p[i] = (int64_t)u[i];
        CU(c, cudaMemcpy(dst, p, c->nf * sizeof(int64_t), cudaMemcpyHostToDevice));
    }
    return LBM_OK;
}

int lbm_get_neighbor_table(lbm_ctx *c, int32_t *dst) {
    CTX_CHECK(c);
    if (!dst) FAIL(c, LBM_ERR_INVALID, "null destination");
    if (!c->inited || !c->cfg.sparse) FAIL(c, LBM_ERR_STATE, "needs an initialised sparse context");
    CU(c, cudaSetDevice(c->cfg.device));
    int32_t *tmp = nullptr;
    CU(c, big_alloc(c, (void **)&tmp, (c->nf ? c->nf : 1) * 18 * sizeof(int32_t)));
    StepArgs a;
    fill_args(c, a);
    if (c->nf) k_decode_table<<<nblocks(c->nf, 256), 256>>>(a, (uint32_t)c->nf, tmp);
    cudaError_t e = cudaGetLastError();
    if (e == cudaSuccess) e = cudaMemcpy(dst, tmp, c->nf * 18 * sizeof(int32_t), cudaMemcpyDefault);
    lbm_pool::release(tmp);
    CU(c, e);
    c->launches++;
    return LBM_OK;
}

int lbm_get_link_flags(lbm_ctx *c, uint32_t *dst) {
    CTX_CHECK(c);
    if (!dst) FAIL(c, LBM_ERR_INVALID, "null destination");
    if (!c->inited) FAIL(c, LBM_ERR_STATE, "lbm_init has not been called");
    CU(c, cudaSetDevice(c->cfg.device));
    const size_t n = c->cfg.sparse ? c->nf : c->N;
    CU(c, cudaMemcpy(dst, c->d_flags, n * sizeof(uint32_t), cudaMemcpyDefault));
    return LBM_OK;
}

// ---- multi-GPU halo staging ----------------------------------------------------------------
int64_t lbm_halo_count(lbm_ctx *c, int plane) {
    if (!c || !c->inited || !c->cfg.halo_x || plane < 0 || plane > 3) return -1;
    return (int64_t)c->plane_count[plane];
}

int lbm_halo_pack(lbm_ctx *c, int side, int which, float *dst, void *cuda_stream) {
    CTX_CHECK(c);
    if (!c->inited || !c->cfg.halo_x) FAIL(c, LBM_ERR_STATE, "needs an initialised halo_x context");
    if (!c->pipe_valid) FAIL(c, LBM_ERR_STATE, "no post-collision state yet (call lbm_step_begin)");
    if (side < 0 || side > 1 || !dst) FAIL(c, LBM_ERR_INVALID, "bad side/destination");
    CU(c, cudaSetDevice(c->cfg.device));
    return halo_pack_impl(c, side, which, dst, (cudaStream_t)cuda_stream);
}

int lbm_halo_unpack(lbm_ctx *c, int side, int which, const float *src, void *cuda_stream) {
    CTX_CHECK(c);
    if (!c->inited || !c->cfg.halo_x) FAIL(c, LBM_ERR_STATE, "needs an initialised halo_x context");
    if (side < 0 || side > 1 || !src) FAIL(c, LBM_ERR_INVALID, "bad side/source");
    CU(c, cudaSetDevice(c->cfg.device));
    return halo_unpack_impl(c, side, which, src, (cudaStream_t)cuda_stream);
}

int lbm_p2p_export(lbm_ctx *c, void *blob256) {
    CTX_CHECK(c);
    if (!blob256) FAIL(c, LBM_ERR_INVALID, "null blob");
    if (!c->inited || !c->cfg.halo_x) FAIL(c, LBM_ERR_STATE, "needs an initialised halo_x context");
    if (c->cfg.sparse || c->aa) FAIL(c, LBM_ERR_STATE, "the peer-memory halo serves dense two-buffer slabs");
    CU(c, cudaSetDevice(c->cfg.device));
    if (!c->d_p2p) {
        CU(c, big_alloc(c, (void **)&c->d_p2p, 64));
        CU(c, cudaMemset(c->d_p2p, 0, 64));
    }
    P2pBlob b;
    memset(&b, 0, sizeof b);
    CU(c, cudaIpcGetMemHandle(&b.f[0], c->d_fbase[0]));
    CU(c, cudaIpcGetMemHandle(&b.f[1], c->d_fbase[1]));
    CU(c, cudaIpcGetMemHandle(&b.flags, c->d_p2p));
    b.nx = c->cfg.nx; b.ny = c->cfg.ny; b.nz = c->cfg.nz; b.layout = c->layout;
    b.prow = c->prow; b.pad = c->pad;
    b.pstride = c->layout == 1 ? (uint64_t)c->nzp : (uint64_t)c->stride;
    memset(blob256, 0, LBM_P2P_BLOB_BYTES);
    memcpy(blob256, &b, sizeof b);
    return LBM_OK;
}

int lbm_p2p_connect(lbm_ctx *c, const void *left_blob256, const void *right_blob256) {
    CTX_CHECK(c);
    if (!left_blob256 || !right_blob256) FAIL(c, LBM_ERR_INVALID, "null blob");
    if (!c->d_p2p) FAIL(c, LBM_ERR_STATE, "lbm_p2p_export has not been called");
    CU(c, cudaSetDevice(c->cfg.device));
    P2pBlob b[2];
    memcpy(&b[0], left_blob256, sizeof(P2pBlob));
    memcpy(&b[1], right_blob256, sizeof(P2pBlob));
    c->peer_shared = memcmp(&b[0], &b[1], sizeof(P2pBlob)) == 0;      // a world of two: one neighbour on both sides
    for (int side = 0; side < 2; ++side) {
        const P2pBlob &p = b[side];
        if (p.ny != c->cfg.ny || p.nz != c->cfg.nz || p.layout != c->layout || p.prow != c->prow || p.pad != c->pad)
            FAIL(c, LBM_ERR_INVALID, "neighbour slab has another cross-section or population layout");
        if (side == 1 && c->peer_shared) {
            for (int k = 0; k < 3; ++k) c->peer_map[1][k] = c->peer_map[0][k];
        } else {
            const cudaIpcMemHandle_t *h[3] = {&p.f[0], &p.f[1], &p.flags};
            for (int k = 0; k < 3; ++k)
                CU(c, cudaIpcOpenMemHandle(&c->peer_map[side][k], *h[k], cudaIpcMemLazyEnablePeerAccess));
        }
        c->peer_f[side][0] = (float *)c->peer_map[side][0] + p.pad;
        c->peer_f[side][1] = (float *)c->peer_map[side][1] + p.pad;
        c->peer_pstride[side] = (long long)p.pstride;
        // the left neighbour counts my launches in its word [1] (I am its right neighbour) and v.v.
        c->peer_flags[side] = (int *)c->peer_map[side][2] + (side == 0 ? 1 : 0);
        // my first owned plane (1) -> the left neighbour's right ghost plane (its nx - 1);
        // my last owned plane (nx - 2) -> the right neighbour's left ghost plane (0)
        const long long rows = side == 0 ? (long long)(p.nx - 1 - 1) * c->cfg.ny : -(long long)(c->cfg.nx - 2) * c->cfg.ny;
        c->peer_delta[side] = rows * (long long)c->prow;
    }
    c->p2p_ready = true;
    return LBM_OK;
}

int lbm_p2p_disconnect(lbm_ctx *c) {
    CTX_CHECK(c);
    cudaSetDevice(c->cfg.device);
    cudaDeviceSynchronize();
    for (int side = 0; side < (c->peer_shared ? 1 : 2); ++side)
        for (int k = 0; k < 3; ++k)
            if (c->peer_map[side][k]) cudaIpcCloseMemHandle(c->peer_map[side][k]);
    memset(c->peer_map, 0, sizeof c->peer_map);
    c->p2p_on = c->p2p_ready = false;
    return LBM_OK;
}

int lbm_p2p_enable(lbm_ctx *c, int on) {
    CTX_CHECK(c);
    if (on && !c->p2p_ready) FAIL(c, LBM_ERR_STATE, "lbm_p2p_connect has not succeeded");
    c->p2p_on = on != 0;
    return LBM_OK;
}

int lbm_comm_ready(int world, int rank) { return g_nccl.ok && have_shared_comm(world, rank) ? 1 : 0; }

int lbm_comm_unique_id(void *out128) {
    std::string err;
    if (!out128) return LBM_ERR_INVALID;
    if (!load_nccl(err)) { g_create_error = err; return LBM_ERR_CUDA; }
    NcclId id;
    if (g_nccl.GetUniqueId(&id) != 0) { g_create_error = "ncclGetUniqueId failed"; return LBM_ERR_CUDA; }
    memcpy(out128, &id, sizeof id);
    return LBM_OK;
}

int lbm_comm_init(lbm_ctx *c, const void *id128, int world, int rank) {
    CTX_CHECK(c);
    if (!c->inited || !c->cfg.halo_x) FAIL(c, LBM_ERR_STATE, "needs an initialised halo_x context");
    if (world < 1 || rank < 0 || rank >= world) FAIL(c, LBM_ERR_INVALID, "bad world/rank");
    CU(c, cudaSetDevice(c->cfg.device));
    if (world > 1) {
        std::string err;
        if (!load_nccl(err)) FAIL(c, LBM_ERR_CUDA, "%s", err.c_str());
        NcclId id;
        memset(&id, 0, sizeof id);
        if (id128) memcpy(&id, id128, sizeof id);
        else if (!have_shared_comm(world, rank)) FAIL(c, LBM_ERR_INVALID, "null NCCL id and no communicator yet");
        NC(c, shared_comm(world, rank, id, &c->comm));
    }
    c->comm_world = world;
    c->comm_rank = rank;
    if (!c->comm_stream) {
        // highest priority: while the interior kernel has tens of thousands of blocks queued,
        // the block scheduler must still pick the few blocks of the pack / NCCL / unpack
        // kernels first, otherwise the exchange only runs when the interior kernel drains
        int lo = 0, hi = 0;
        CU(c, cudaDeviceGetStreamPriorityRange(&lo, &hi));
        CU(c, cudaStreamCreateWithPriority(&c->comm_stream, cudaStreamNonBlocking, hi));
    }
    if (!c->ev_boundary) CU(c, cudaEventCreateWithFlags(&c->ev_boundary, cudaEventDisableTiming));
    if (!c->ev_comm) CU(c, cudaEventCreateWithFlags(&c->ev_comm, cudaEventDisableTiming));
    if (!c->ev_interior) CU(c, cudaEventCreateWithFlags(&c->ev_interior, cudaEventDisableTiming));
    const uint32_t sc[2] = {c->plane_count[1], c->plane_count[2]}, rc[2] = {c->plane_count[0], c->plane_count[3]};
    for (int i = 0; i < 2; ++i) {
        if (!c->d_send[i]) CU(c, big_alloc(c, (void **)&c->d_send[i], (size_t)5 * (sc[i] ? sc[i] : 1) * sizeof(float)));
        if (!c->d_recv[i]) CU(c, big_alloc(c, (void **)&c->d_recv[i], (size_t)5 * (rc[i] ? rc[i] : 1) * sizeof(float)));
    }
    return LBM_OK;
}

// nsteps reference steps of one x-slab with the halo exchange inside (the whole loop runs
// here so that the per-step host cost is a handful of launches, not a Python round trip)
int lbm_run_slab(lbm_ctx *c, int nsteps, int overlap, void *cuda_stream) {
    CTX_CHECK(c);
    if (!c->inited || !c->cfg.halo_x) FAIL(c, LBM_ERR_STATE, "needs an initialised halo_x context");
    if (!c->comm_stream) FAIL(c, LBM_ERR_STATE, "lbm_comm_init has not been called");
    if (nsteps <= 0) return LBM_OK;
    CU(c, cudaSetDevice(c->cfg.device));
    cudaStream_t st = (cudaStream_t)cuda_stream;
    c->stream = st;
    const int own = c->cfg.nx - 2;
    if (own < 3) overlap = 0;
    if (!c->pipe_valid) {
        int r = ensure_pipeline(c, st);
        if (r) return r;
        r = exchange_impl(c, 0, st);
        if (r) return r;
        nsteps -= 1;
    }
    StepArgs a;
    const char *tl_path = getenv("LBM3D_TIMELINE");
    const int tl_first = nsteps > 4 ? nsteps - 4 : 0;          // the last four steps of this call
    for (int it = 0; it < nsteps; ++it) {
        g_timeline.on = tl_path != nullptr && overlap && it >= tl_first;
        if (!overlap) {
            fill_args(c, a);
            set_buffers(c, a, c->d_f[c->cur], c->d_f[c->cur ^ 1]);
            int r = launch(c, MODE_STEP, a, st);
            if (r) return r;
            c->cur ^= 1;
            c->ffm_pending = false;
            r = exchange_impl(c, 0, st);
            if (r) return r;
            continue;
        }
        // Two streams, no per-step join.  Step n reads buffer cur, writes cur ^ 1:
        //   main stream  interior(n)  = planes 2..own-1; reads planes 1..own of cur, never a ghost
        //                plane: needs interior(n-1) (stream order) and boundary(n-1) (ev_boundary)
        //   side stream  boundary(n)  = planes 1 and own in ONE launch; needs interior(n-1)
        //                (ev_interior) and the ghost planes filled by exchange(n-1) (stream order);
        //                then exchange(n): pack (one launch) ; ncclSend/Recv ; unpack (one launch)
        // so consecutive interior kernels run back to back while the side stream (highest
        // priority) has a whole step of slack for the boundary planes and the exchange.
        if (it == 0) {
            CU(c, cudaEventRecord(c->ev_interior, st));          // everything enqueued on st so far
            CU(c, cudaEventRecord(c->ev_boundary, st));
        }
        CU(c, cudaStreamWaitEvent(st, c->ev_boundary, 0));       // boundary(n-1)
        g_timeline.mark("interior_begin", it, st);
        int r = lbm_step_planes(c, 2, own, st);                  // interior(n)
        if (r) return r;
        g_timeline.mark("interior_end", it, st);
        CU(c, cudaStreamWaitEvent(c->comm_stream, c->ev_interior, 0));   // interior(n-1)
        CU(c, cudaEventRecord(c->ev_interior, st));              // interior(n)
        g_timeline.mark("boundary_begin", it, c->comm_stream);
        r = launch_boundary_planes(c, c->comm_stream);           // boundary(n)
        if (r) return r;
        g_timeline.mark("boundary_end", it, c->comm_stream);
        CU(c, cudaEventRecord(c->ev_boundary, c->comm_stream));
        if (!p2p_active(c)) {                                    // else the boundary kernel was the exchange
            r = exchange_merged(c, 1, c->comm_stream, it);       // ghost planes of buffer cur ^ 1
            if (r) return r;
        }
        c->cur ^= 1;
        c->ffm_pending = false;
    }
    if (overlap) {                                               // st continues after both streams
        if (p2p_active(c)) {                                     // ... and after the neighbours' last stores
            k_p2p_drain<<<1, 1, 0, c->comm_stream>>>(c->d_p2p, c->p2p_launches);
            CU(c, cudaGetLastError());
        }
        CU(c, cudaEventRecord(c->ev_comm, c->comm_stream));
        CU(c, cudaStreamWaitEvent(st, c->ev_comm, 0));
    }
    if (tl_path) {
        g_timeline.on = true;
        g_timeline.write(tl_path, c->comm_world > 1 ? c->comm_rank : -1);
        g_timeline.on = false;
    }
    c->macro_valid = false;
    c->F_valid = false;
    return LBM_OK;
}

int lbm_step_begin(lbm_ctx *c, void *cuda_stream) {
    CTX_CHECK(c);
    if (!c->inited) FAIL(c, LBM_ERR_STATE, "lbm_init has not been called");
    CU(c, cudaSetDevice(c->cfg.device));
    c->stream = (cudaStream_t)cuda_stream;
    if (c->pipe_valid) return 1;      // nothing to do: a full step is pending
    int r = ensure_pipeline(c, c->stream);
    if (r) return r;
    c->macro_valid = false;
    c->F_valid = false;
    return LBM_OK;
}

int lbm_step_planes(lbm_ctx *c, int x_begin, int x_end, void *cuda_stream) {
    CTX_CHECK(c);
    if (!c->inited || !c->pipe_valid) FAIL(c, LBM_ERR_STATE, "pipeline not started");
    if (c->aa) FAIL(c, LBM_ERR_STATE, "plane-wise stepping needs two buffers (create the context with sparse = 0 or 1)");
    const int lo = c->cfg.halo_x ? 1 : 0, hi = c->cfg.halo_x ? c->cfg.nx - 1 : c->cfg.nx;
    if (x_begin < lo || x_end > hi || x_begin > x_end) FAIL(c, LBM_ERR_INVALID, "plane range outside the owned slab");
    CU(c, cudaSetDevice(c->cfg.device));
    StepArgs a;
    fill_args(c, a);
    set_buffers(c, a, c->d_f[c->cur], c->d_f[c->cur ^ 1]);
    if (!c->cfg.sparse) {
        a.row_first = (uint32_t)(c->cfg.ny * x_begin);
        a.row_count = (uint32_t)(c->cfg.ny * (x_end - x_begin));
    } else {
        a.first = c->plane_rank[x_begin];
        a.count = c->plane_rank[x_end] - a.first;
    }
    return launch(c, MODE_STEP, a, (cudaStream_t)cuda_stream);
}

int lbm_step_flip(lbm_ctx *c) {
    CTX_CHECK(c);
    if (!c->inited || !c->pipe_valid) FAIL(c, LBM_ERR_STATE, "pipeline not started");
    if (c->aa) FAIL(c, LBM_ERR_STATE, "plane-wise stepping needs two buffers (create the context with sparse = 1)");
    c->cur ^= 1;
    c->ffm_pending = false;
    c->macro_valid = false;
    c->F_valid = false;
    return LBM_OK;
}

int lbm_get_device_ptr(lbm_ctx *c, int which, void **ptr, size_t *bytes) {
    CTX_CHECK(c);
    if (!ptr || !bytes) FAIL(c, LBM_ERR_INVALID, "null output");
    if (!c->inited) FAIL(c, LBM_ERR_STATE, "lbm_init has not been called");
    switch (which) {
        case LBM_BUF_F_CUR: *ptr = c->d_f[c->cur]; *bytes = c->fsize * sizeof(float); break;
        case LBM_BUF_F_NEXT: *ptr = c->d_f[c->cur ^ 1]; *bytes = c->fsize * sizeof(float); break;
        case LBM_BUF_RHO: *ptr = c->d_rho; *bytes = c->N * sizeof(float); break;
        case LBM_BUF_V: *ptr = c->d_v; *bytes = c->N * 3 * sizeof(float); break;
        case LBM_BUF_FLAGS: *ptr = c->d_flags; *bytes = (c->cfg.sparse ? c->nf : c->N) * sizeof(uint32_t); break;
        default: FAIL(c, LBM_ERR_INVALID, "unknown buffer id");
    }
    return LBM_OK;
}

int lbm_get_layout(lbm_ctx *c, int64_t out[4]) {
    CTX_CHECK(c);
    if (!out) FAIL(c, LBM_ERR_INVALID, "null output");
    if (!c->inited) FAIL(c, LBM_ERR_STATE, "lbm_init has not been called");
    const bool blocked = !c->cfg.sparse && c->layout == 1;
    out[0] = c->cfg.sparse ? 2 : c->layout;                      // 0 SoA, 1 row-blocked, 2 compact list
    out[1] = blocked ? (int64_t)c->nzp : (int64_t)c->stride;     // elements between planes s and s+1
    out[2] = (int64_t)c->prow;                                   // elements between z-rows (dense)
    out[3] = (int64_t)c->fsize;                                  // elements per buffer
    return LBM_OK;
}

}  // extern "C"
